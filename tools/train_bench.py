"""Training-step measurement helper: BASELINE.json configs[2] (SSD300-VGG16, batch 32 per GPU) and configs[3]
(SSD300-MobileNetV2, batch 32 per GPU, 256 over 8 GPUs).  One full training step = IoU match + target encode ->
forward (BatchNorm on batch statistics for MobileNetV2) -> hard-negative-mining loss -> backward -> (gradient
all-reduce) -> Adam.  Used by bench.py (``training`` key) and runnable on its own / under torchrun.

    python tools/train_bench.py [--backbone vgg16|mobilenet_v2] [--steps K] [--warmup W] [--batch B] [--check-dp]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


BACKBONE = "vgg16"


def build(batch, seed=1234):
    import torch
    from tf_ssd_b200 import synth
    from tf_ssd_b200.models import ssd_mobilenet_v2, ssd_vgg16
    from tf_ssd_b200.models.train_engine import Trainer
    from tf_ssd_b200.utils import bbox_utils, train_utils
    hp = train_utils.get_hyper_params(BACKBONE)
    hp["total_labels"] = 21
    model = (ssd_mobilenet_v2 if BACKBONE == "mobilenet_v2" else ssd_vgg16).get_model(hp, seed=seed)
    trainer = Trainer(model)
    priors = bbox_utils.generate_prior_boxes(hp["feature_map_shapes"], hp["aspect_ratios"])
    return hp, model, trainer, priors


def make_batch(batch, hp, seed):
    import torch
    from tf_ssd_b200 import synth
    img = torch.from_numpy(synth.make_images(batch, hp["img_size"], seed=seed)).cuda()
    gt, lab = synth.make_ground_truth(batch, padded=16, seed=seed + 1)
    return img, torch.from_numpy(gt).cuda(), torch.from_numpy(lab).cuda()


def train_step(trainer, priors, hp, img, gt, lab):
    from tf_ssd_b200.utils import train_utils
    deltas, onehot = train_utils.calculate_actual_outputs(priors, gt, lab, hp)       # IoU match + encode (fused kernel)
    trainer.forward_backward(img, deltas, onehot)
    trainer.apply_gradients()


def measure(steps=10, warmup=3, batch=32):
    """Returns a dict with images/s (whole job) for the training step; call under torchrun for N > 1."""
    import torch
    from tf_ssd_b200 import dist_utils
    rank, local_rank, world = dist_utils.env_rank()
    torch.cuda.set_device(local_rank)
    dist_utils.init_from_env("nccl")
    hp, model, trainer, priors = build(batch)
    data = [make_batch(batch, hp, 100 + 10 * rank + i) for i in range(2)]
    for i in range(warmup):
        train_step(trainer, priors, hp, *data[i % 2])
    dist_utils.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        train_step(trainer, priors, hp, *data[i % 2])
    e1.record()
    dist_utils.barrier()
    torch.cuda.synchronize()
    ms = dist_utils.max_over_ranks(float(e0.elapsed_time(e1)), torch.device("cuda", local_rank))
    flops = 3 * 2.0 * model.macs_per_image * batch                        # fwd + dgrad + wgrad
    name = "SSD300-MobileNetV2" if BACKBONE == "mobilenet_v2" else "SSD300-VGG16"
    return {"workload": f"{name} batch={batch}/GPU training step (IoU match + encode, forward, hard-negative loss, backward, "
                        "grad all-reduce, Adam), fp16 activations / fp32 master weights",
            "value": world * batch * steps / (ms / 1e3), "unit": "images/s", "ms_per_step": ms / steps, "n_gpus": world,
            "steps": steps, "conv_tflops": world * flops * steps / (ms / 1e3) / 1e12,
            "grad_allreduce_bytes": int(sum(b.numel() for b in trainer.grads.buckets) * 4)}


def breakdown(batch=32):
    """Per-category CUDA-event timing of one training step (no profiler needed)."""
    import collections
    import torch
    from tf_ssd_b200 import _ffi
    from tf_ssd_b200.utils import train_utils
    torch.cuda.set_device(0)
    hp, model, trainer, priors = build(batch)
    img, gt, lab = make_batch(batch, hp, 5)
    for _ in range(2):
        train_step(trainer, priors, hp, img, gt, lab)
    torch.cuda.synchronize()
    st = trainer._prepare(batch)
    plan = st["plan"]
    evs = []

    def mark(tag):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        evs.append((tag, e))

    mark("start")
    d, oh = train_utils.calculate_actual_outputs(priors, gt, lab, hp); mark("match_encode")
    model._to_image_buffer(plan, img)
    for i, s_ in enumerate(plan.steps):
        plan.run(i, i + 1); mark("fwd:" + s_.kind)
    stream = _ffi.stream()
    lib = trainer.lib
    N, L = model.n_anchors, model.total_labels
    _ffi.check(lib.ssd_loss_fwd(_ffi.ptr(d), _ffi.ptr(plan.deltas), _ffi.ptr(oh), _ffi.ptr(plan.logits), batch, N, L, 3.0, 1.0, 1,
                                _ffi.ptr(st["loc"]), _ffi.ptr(st["conf"]), _ffi.ptr(st["ws"]), st["ws"].numel(), stream)); mark("loss_fwd")
    _ffi.check(lib.ssd_loss_bwd(_ffi.ptr(d), _ffi.ptr(plan.deltas), _ffi.ptr(oh), _ffi.ptr(plan.logits), batch, N, L, 1.0,
                                trainer.loss_scale / batch, _ffi.ptr(st["g_deltas"]), _ffi.ptr(st["g_logits"]), _ffi.ptr(st["ws"]),
                                st["ws"].numel(), stream)); mark("loss_bwd")
    trainer.grads.zero_(); mark("zero_grads")
    per_layer = []
    for fn, args, what in st["launches"]:
        _ffi.check(fn(*args, stream), what); mark("bwd:" + what.split(":")[-1])
        per_layer.append(what)
    trainer.apply_gradients(); mark("adam")
    torch.cuda.synchronize()
    agg = collections.OrderedDict()
    detail = []
    for (t0, e0), (t1, e1) in zip(evs[:-1], evs[1:]):
        ms = e0.elapsed_time(e1)
        agg[t1] = agg.get(t1, 0.0) + ms
        detail.append((t1, ms))
    total = sum(agg.values())
    print(json.dumps({"total_ms": total, "by_category_ms": {k: round(v, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])}}))
    bw = [(w, ms) for (w, (t, ms)) in zip(per_layer, [d_ for d_ in detail if d_[0].startswith("bwd:")])]
    top = sorted(bw, key=lambda t: -t[1])[:14]
    print(json.dumps({"top_backward_launches_ms": [(w, round(ms, 3)) for w, ms in top]}))


def check_dp():
    """2 ranks x B=2 must produce the same averaged gradients as 1 process on the concatenated batch of 4."""
    import torch
    from tf_ssd_b200 import dist_utils
    from tf_ssd_b200.utils import train_utils
    rank, local_rank, world = dist_utils.env_rank()
    torch.cuda.set_device(local_rank)
    dist_utils.init_from_env("nccl")
    hp, model, trainer, priors = build(2)
    img, gt, lab = make_batch(2 * world, hp, 7)
    b, e = dist_utils.shard_range(2 * world, rank, world)
    d, oh = train_utils.calculate_actual_outputs(priors, gt[b:e], lab[b:e], hp)
    trainer.forward_backward(img[b:e], d, oh)            # segments + overlapped SUM all-reduce per bucket
    for w in trainer._pending:
        w.wait()
    trainer._pending = []
    torch.cuda.synchronize()
    mine = torch.cat([bk.flatten() for bk in trainer.grads.buckets]) / world
    if rank == 0 and BACKBONE == "mobilenet_v2":
        # BatchNorm statistics stay per replica (Keras default): the reference is the mean of the
        # per-shard gradients computed one after the other in this process
        hp1, model1, trainer1, _ = build(2)
        ref = None
        for r in range(world):
            b1, e1 = dist_utils.shard_range(2 * world, r, world)
            d1, oh1 = train_utils.calculate_actual_outputs(priors, gt[b1:e1], lab[b1:e1], hp1)
            trainer1.forward_backward(img[b1:e1], d1, oh1, reduce=False)
            torch.cuda.synchronize()
            g1 = torch.cat([bk.flatten() for bk in trainer1.grads.buckets])
            ref = g1 if ref is None else ref + g1
        ref = ref / world
        rel = float((mine - ref).norm() / ref.norm())
        print(json.dumps({"dp_equivalence_rel_l2": rel, "world": world, "ok": rel < 2e-3, "backbone": BACKBONE}))
        assert rel < 2e-3, rel
    elif rank == 0:
        hp1, model1, trainer1, _ = build(2 * world)
        d1, oh1 = train_utils.calculate_actual_outputs(priors, gt, lab, hp1)
        trainer1.forward_backward(img, d1, oh1, reduce=False)
        torch.cuda.synchronize()
        ref = torch.cat([bk.flatten() for bk in trainer1.grads.buckets])
        rel = float((mine - ref).norm() / ref.norm())
        print(json.dumps({"dp_equivalence_rel_l2": rel, "world": world, "ok": rel < 2e-3}))
        assert rel < 2e-3, rel
    dist_utils.barrier()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--backbone", default="vgg16", choices=["vgg16", "mobilenet_v2"])
    ap.add_argument("--check-dp", action="store_true")
    ap.add_argument("--breakdown", action="store_true")
    a = ap.parse_args()
    BACKBONE = a.backbone
    if a.breakdown:
        breakdown(a.batch)
    elif a.check_dp:
        check_dp()
    else:
        out = measure(a.steps, a.warmup, a.batch)
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps(out), flush=True)
    import torch.distributed as _dist
    if _dist.is_initialized():
        _dist.barrier()
        _dist.destroy_process_group()
