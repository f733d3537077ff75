#!/usr/bin/env python
"""Benchmark of the tf-ssd hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One step = one pass of the inference hot path over one synthetic VOC-shaped
batch: SSD300-MobileNetV2, batch 32 per GPU, forward + softmax + decode + NMS
(BASELINE.json configs[1]).  Prints ONE JSON line (rank 0).

  value  images/s, whole job, inputs resident in HBM, CUDA events, max over ranks
  e2e    the same metric through the reference-facing API
         (get_model -> get_decoder_model -> predict) with HOST batches: pinned
         H2D of every step's images and D2H of its detections inside the timing
  roofline / cpu_baseline / box_kernels: see DESIGN.md

Input batches are uint8 NHWC (the image the reference's pipeline holds before
``tf.image.convert_image_dtype``, utils/data_utils.py:33-37); the conversion to
[0,1] float is part of the step in BOTH arms (fused into the first layer on the
GPU, ``u8 * float32(1/255)`` in the CPU restatement).

``--impl reference`` times the CPU restatement of the reference (oracle/, torch
CPU + NumPy, all host threads) on the SAME workload -- same weights recipe (seeded
random init, background bias calibrated to 200 NMS candidates per image), same
synthetic batches, same ``config``: the reference itself is TensorFlow 2.0 and
cannot be installed in this image (no wheel, no network).
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BACKBONE = "mobilenet_v2"
BATCH = 32
WORKLOAD = "SSD300-MobileNetV2 batch=32/GPU inference fwd+softmax+decode+NMS, 2268 anchors, 21 labels, fp16 convs / fp32 boxes"
METRIC = "SSD300 images/sec (fwd+decode+NMS)"
UNIT = "images/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "tflops_burst": float(p["bf16_tflops"]),
                "tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


def _ncu_traffic(kind, launches_per_step):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the dominant kernel group, from the
    committed ``ncu --set full`` capture of one step of this same workload (profiles/r2_traffic.json, written by
    tools/ncu_traffic.py): the group's DRAM bytes per step / its launches per step, i.e. per launch like ``achieved``.
    None when no capture is committed for that group."""
    e = _ncu_group(kind)
    return None if e is None else float(e["dram_bytes"]) / max(int(launches_per_step), 1)


def _ncu_group(kind):
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get(kind)


def _hyper_params():
    from tf_ssd_b200.utils import train_utils
    hp = train_utils.get_hyper_params(BACKBONE)
    hp["total_labels"] = 21
    return hp


CAND_TARGET = 200.0


def _config(world):
    """The workload description: identical in both arms (the reference arm runs a bounded sample of it)."""
    return {"workload": WORKLOAD, "global_batch": world * BATCH, "parallelism": f"independent inference shards x{world}",
            "input": "uint8 NHWC 300x300x3 host batches (pre-convert_image_dtype), conversion inside the step",
            "weights": f"random-init (seed 1234), BN folded, background bias calibrated to {CAND_TARGET:.0f} NMS candidates/image",
            "l2": "256 MiB memset between timed steps (outside the event pairs)", "cuda_graph": True}


def _make_images_u8(batch, size, seed):
    return np.random.default_rng(seed).integers(0, 256, (batch, size, size, 3), dtype=np.uint8)


def _calibrated_weights(model, hp, seed=1234, forward=None):
    """Random-init weights of the architecture (no checkpoints offline), with the
    background logit bias shifted so that a detector-like number of anchors
    (about 200 per image) passes the 0.5 score threshold and reaches NMS.
    ``forward(images_f32) -> logits`` replaces the GPU forward (the reference arm calibrates with the CPU oracle)."""
    img = _make_images_u8(4, hp["img_size"], seed).astype(np.float32) * np.float32(1.0 / 255.0)
    if forward is None:
        import torch
        _, z = model.forward_logits(img)
        torch.cuda.synchronize()
        z = z.cpu().numpy()
    else:
        z = forward(img)
    z = z.astype(np.float64)

    def candidates(shift):
        zz = z.copy()
        zz[..., 0] += shift
        zz -= zz.max(-1, keepdims=True)
        p = np.exp(zz)
        p /= p.sum(-1, keepdims=True)
        return float(((p[..., 1:].max(-1) > 0.5) & (p.argmax(-1) != 0)).sum() / z.shape[0])

    lo, hi = -50.0, 50.0
    for _ in range(40):
        mid = 0.5 * (lo + hi)
        if candidates(mid) > CAND_TARGET:
            lo = mid
        else:
            hi = mid
    shift = 0.5 * (lo + hi)
    w = {}
    for i in range(1, len(hp["feature_map_shapes"]) + 1):
        b = model.weights[f"{i}_conv_label_output/bias"].copy().reshape(-1, 21)
        b[:, 0] += np.float32(shift)
        w[f"{i}_conv_label_output/bias"] = b.reshape(-1)
    if forward is None:
        model.set_weights(w)
    else:
        model.weights.update(w)            # host-side only: the reference arm never touches the GPU
    return candidates(shift)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = int(get(self.h))
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop_evt.wait(0.02)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def _physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------ CPU arms --
def _cpu_reference_step(weights, hp, priors, images_u8):
    """The reference's inference path restated on the CPU (oracle/): convert_image_dtype, forward in
    fp32 (torch CPU conv2d), softmax, SSDDecoder (NumPy)."""
    from oracle import box_oracle as bo
    from oracle import net_oracle as no
    images = images_u8.astype(np.float32) * np.float32(1.0 / 255.0)         # utils/data_utils.py:36
    d, p = no.forward(BACKBONE, weights, hp, images, mode="fp32")
    return bo.ssd_decode(priors, hp["variances"], d, p)


def cpu_baseline(weights, hp, priors, budget_s=12.0):
    import torch
    from tf_ssd_b200 import synth
    cores = torch.get_num_threads()
    img = _make_images_u8(BATCH, hp["img_size"], seed=1000)     # rank 0's first batch of the GPU arm
    _cpu_reference_step(weights, hp, priors, img[:4])           # warm-up (thread pools, allocator)
    n, t0 = 0, time.perf_counter()
    while True:
        _cpu_reference_step(weights, hp, priors, img)
        n += BATCH
        el = time.perf_counter() - t0
        if el > budget_s or n >= 8 * BATCH:
            break
    return {"value": n / el, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} images (batches of {BATCH}) through oracle/net_oracle.py fp32 forward + oracle/box_oracle.py "
                      f"ssd_decode in {el:.1f} s; host has {os.cpu_count()} logical CPUs; restated reference, not TensorFlow"}


def run_reference(args):
    """--impl reference: the CPU restatement timed with all host threads (rank 0 only), same workload as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import box_oracle as bo
    from oracle import net_oracle as no
    from tf_ssd_b200.models.engine import SSDModel
    torch.set_num_threads(os.cpu_count() or 1)         # torchrun exports OMP_NUM_THREADS=1: use every host core
    hp = _hyper_params()
    model = SSDModel(BACKBONE, hp, seed=1234)          # host-side variable initialisation only (no GPU use)
    cand = _calibrated_weights(model, hp, forward=lambda x: no.forward(BACKBONE, model.weights, hp, x, mode="fp32",
                                                                       return_logits=True)[1])
    weights = model.weights
    priors = bo.prior_boxes(hp["feature_map_shapes"], hp["aspect_ratios"])
    # bounded sample per step so that steps+warmup stay within minutes
    img = _make_images_u8(BATCH, hp["img_size"], seed=1000)         # rank 0's first batch of the GPU arm
    t0 = time.perf_counter()
    _cpu_reference_step(weights, hp, priors, img[:4])
    per_img = (time.perf_counter() - t0) / 4
    total_steps = args.steps + args.warmup
    sample = int(max(1, min(BATCH, 150.0 / max(per_img * total_steps, 1e-9))))
    for _ in range(args.warmup):
        _cpu_reference_step(weights, hp, priors, img[:sample])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = _cpu_reference_step(weights, hp, priors, img[:sample])
    el = time.perf_counter() - t0
    value = args.steps * sample / el
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(args.gpus),
        "workload_check": {"nms_candidates_per_image": round(cand, 1),
                           "valid_detections_mean": float((out[2] > 0).sum(-1).mean()), "images_per_step": sample},
        "note": "CPU restatement of the reference (oracle/: torch-CPU fp32 forward + NumPy SSDDecoder) on rank 0's host "
                "cores; TensorFlow 2.0 is not installable in this image",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} images per step x {args.steps} steps, {os.cpu_count()} logical CPUs"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------- GPU arm --
def _profile_steps(dm, B, iters=5):
    """Per-launch CUDA-event timing of every kernel of one step (eager launches
    on the current stream; events recorded on that same stream)."""
    import torch
    st = dm._prepare(B, 0)
    plan = st["plan"]
    n = plan.n_launches
    acc = np.zeros(n + 1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=plan.device)
    all_evs = []
    # everything is queued without host synchronisation so that the GPU never waits for the host
    # between the short launches (an idle gap would be attributed to the following kernel)
    for it in range(iters + 1):
        flush.zero_()
        flush.zero_()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 2)]
        evs[0].record()
        for i in range(n):
            plan.run(i, i + 1, u8=True)
            evs[i + 1].record()
        st["enqueue_decode"]()
        evs[n + 1].record()
        all_evs.append(evs)
    torch.cuda.synchronize()
    for evs in all_evs[1:]:                             # first pass is warm-up
        acc += np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(n + 1)])
    return acc / iters                                  # ms per launch; last entry = decode+NMS (2 kernels)


def _box_kernel_rooflines(peaks, hp, iters=10, warmup=3):
    """BASELINE.json's second metric: anchor-IoU / matching / loss / decode kernels as HBM GB/s at the stress
    shape of SURVEY.md 8(d) (B=256, N=24564, G=42: working sets far larger than L2)."""
    import torch
    from tf_ssd_b200 import _ffi, synth
    from tf_ssd_b200.utils import bbox_utils, train_utils
    hp512 = train_utils.get_hyper_params("vgg16_512")
    hp512["total_labels"] = 21
    priors = bbox_utils.generate_prior_boxes(hp512["feature_map_shapes"], hp512["aspect_ratios"])
    B, N, G, L = 256, priors.shape[0], 42, 21
    lib = _ffi.lib()
    dev = priors.device
    out = {}

    def timeit(fn, nbytes):
        for _ in range(warmup):
            fn()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for a, b in evs:
            a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
        gbs = nbytes / ms / 1e6
        return {"ms": ms, "bytes": nbytes, "achieved_gbs": gbs, "frac": gbs / peaks["hbm_gbs"]}

    var = _ffi.f32_array(hp["variances"])
    deltas = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
    onehot = torch.empty((B, N, L), dtype=torch.float32, device=dev)
    for g_pad, max_boxes, tag in ((G, G, ""), (16, 8, "_G16")):     # stress (42 real boxes/image) and VOC-like (<= 8 of 16)
        gt, lab = synth.make_ground_truth(B, padded=g_pad, max_boxes=max_boxes, seed=5)
        gt_d, lab_d = _ffi.to_dev(gt), _ffi.to_dev(lab, dtype=torch.int32)
        iou = torch.empty((B, N, g_pad), dtype=torch.float32, device=dev)
        out["generate_iou_map" + tag] = timeit(
            lambda: _ffi.check(lib.ssd_iou_map(_ffi.ptr(priors), _ffi.ptr(gt_d), B, N, g_pad, 0, _ffi.ptr(iou), _ffi.stream())),
            4 * B * N * g_pad + 16 * N + 16 * B * g_pad)
        del iou
        out["match_encode" + tag] = timeit(
            lambda: _ffi.check(lib.ssd_match_encode(_ffi.ptr(priors), _ffi.ptr(gt_d), _ffi.ptr(lab_d), B, N, g_pad, L, 0.5, var,
                                                    _ffi.ptr(deltas), _ffi.ptr(onehot), None, None, _ffi.stream())),
            16 * N + 20 * B * g_pad + B * N * (16 + 4 * L))
    # batch sweep of the anchor-IoU kernel: where the launch stops being latency-bound
    sweep = {}
    for b_ in (8, 32, 64, 128, 256):
        gt, _ = synth.make_ground_truth(b_, padded=16, max_boxes=8, seed=5)
        gt_d = _ffi.to_dev(gt)
        iou = torch.empty((b_, N, 16), dtype=torch.float32, device=dev)
        r = timeit(lambda: _ffi.check(lib.ssd_iou_map(_ffi.ptr(priors), _ffi.ptr(gt_d), b_, N, 16, 0, _ffi.ptr(iou), _ffi.stream())),
                   4 * b_ * N * 16 + 16 * N + 16 * b_ * 16)
        sweep[str(b_)] = {"ms": round(r["ms"], 5), "frac": round(r["frac"], 4)}
        del iou
    out["generate_iou_map_G16_batch_sweep"] = sweep
    pd, logits = synth.make_head_outputs(B, N, L, seed=6, background_bias=10.0)   # ~2.5 % of the anchors reach NMS
    pd_d, z_d = _ffi.to_dev(pd), _ffi.to_dev(logits)
    ws = _ffi.workspace(lib.ssd_loss_workspace_bytes(B, N, L))
    loc = torch.empty(B, dtype=torch.float32, device=dev)
    conf = torch.empty(B, dtype=torch.float32, device=dev)
    out["ssd_loss_fwd"] = timeit(
        lambda: _ffi.check(lib.ssd_loss_fwd(_ffi.ptr(deltas), _ffi.ptr(pd_d), _ffi.ptr(onehot), _ffi.ptr(z_d), B, N, L, 3.0,
                                            1.0, 1, _ffi.ptr(loc), _ffi.ptr(conf), _ffi.ptr(ws), ws.numel(), _ffi.stream())),
        B * N * (16 + 16 + 4 * L + 4 * L) + 8 * B)
    T = 200
    ws2 = _ffi.workspace(lib.ssd_decode_nms_workspace_bytes(B, N, L, T, 0))
    ob = torch.empty((B, T, 4), dtype=torch.float32, device=dev)
    ol = torch.empty((B, T), dtype=torch.float32, device=dev)
    os_ = torch.empty((B, T), dtype=torch.float32, device=dev)
    ov = torch.empty((B,), dtype=torch.int32, device=dev)
    out["decode_nms"] = timeit(
        lambda: _ffi.check(lib.ssd_decode_nms(_ffi.ptr(priors), _ffi.ptr(pd_d), _ffi.ptr(z_d), B, N, L, var, 1, 0.5, 0.5, T,
                                              0, _ffi.ptr(ob), _ffi.ptr(ol), _ffi.ptr(os_), _ffi.ptr(ov), _ffi.ptr(ws2),
                                              ws2.numel(), _ffi.stream())),
        B * N * (16 + 4 * L) + 16 * N + 24 * B * T)
    # the same four kernels at the BASELINE shapes (launch-latency bound: 5-56 MB per launch = 1-9 us at the HBM peak) and
    # over the batch size at N = 24 564: where each launch stops being latency-bound
    del deltas, onehot, pd_d, z_d, ws, ws2
    torch.cuda.empty_cache()
    out["baseline_shapes"] = {}
    for tag, backbone, b_ in (("cfg2_mobilenet_v2_B32_N2268", "mobilenet_v2", 32), ("cfg3_vgg16_B32_N8732", "vgg16", 32),
                              ("cfg5_vgg16_512_B16_N24564", "vgg16_512", 16)):
        out["baseline_shapes"][tag] = _box_kernels_at(lib, timeit, var, backbone, b_, 16, L)
    sweep_all = {}
    for b_ in (8, 32, 128, 256):
        r = _box_kernels_at(lib, timeit, var, "vgg16_512", b_, 16, L)
        for k, v in r.items():
            sweep_all.setdefault(k, {})[str(b_)] = {"ms": v["ms"], "frac": v["frac"]}
    out["batch_sweep_N24564_G16"] = sweep_all
    return {"shape": {"B": B, "N": N, "G": G, "L": L}, "peak_gbs": peaks["hbm_gbs"], "kernels": out}


def _box_kernels_at(lib, timeit, var, backbone, B, G, L):
    """generate_iou_map / match_encode / ssd_loss_fwd / decode_nms at one (network, batch) shape: ms and HBM fraction."""
    import torch
    from tf_ssd_b200 import _ffi, synth
    from tf_ssd_b200.utils import bbox_utils, train_utils
    hp = train_utils.get_hyper_params(backbone)
    hp["total_labels"] = L
    priors = bbox_utils.generate_prior_boxes(hp["feature_map_shapes"], hp["aspect_ratios"])
    N, dev = priors.shape[0], priors.device
    gt, lab = synth.make_ground_truth(B, padded=G, max_boxes=8, seed=5)
    gt_d, lab_d = _ffi.to_dev(gt), _ffi.to_dev(lab, dtype=torch.int32)
    res = {}
    iou = torch.empty((B, N, G), dtype=torch.float32, device=dev)
    res["generate_iou_map"] = timeit(
        lambda: _ffi.check(lib.ssd_iou_map(_ffi.ptr(priors), _ffi.ptr(gt_d), B, N, G, 0, _ffi.ptr(iou), _ffi.stream())),
        4 * B * N * G + 16 * N + 16 * B * G)
    deltas = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
    onehot = torch.empty((B, N, L), dtype=torch.float32, device=dev)
    res["match_encode"] = timeit(
        lambda: _ffi.check(lib.ssd_match_encode(_ffi.ptr(priors), _ffi.ptr(gt_d), _ffi.ptr(lab_d), B, N, G, L, 0.5, var,
                                                _ffi.ptr(deltas), _ffi.ptr(onehot), None, None, _ffi.stream())),
        16 * N + 20 * B * G + B * N * (16 + 4 * L))
    pd, logits = synth.make_head_outputs(B, N, L, seed=6, background_bias=10.0)
    pd_d, z_d = _ffi.to_dev(pd), _ffi.to_dev(logits)
    ws = _ffi.workspace(lib.ssd_loss_workspace_bytes(B, N, L))
    loc = torch.empty(B, dtype=torch.float32, device=dev)
    conf = torch.empty(B, dtype=torch.float32, device=dev)
    res["ssd_loss_fwd"] = timeit(
        lambda: _ffi.check(lib.ssd_loss_fwd(_ffi.ptr(deltas), _ffi.ptr(pd_d), _ffi.ptr(onehot), _ffi.ptr(z_d), B, N, L, 3.0,
                                            1.0, 1, _ffi.ptr(loc), _ffi.ptr(conf), _ffi.ptr(ws), ws.numel(), _ffi.stream())),
        B * N * (16 + 16 + 4 * L + 4 * L) + 8 * B)
    T = 200
    ws2 = _ffi.workspace(lib.ssd_decode_nms_workspace_bytes(B, N, L, T, 0))
    ob = torch.empty((B, T, 4), dtype=torch.float32, device=dev)
    ol = torch.empty((B, T), dtype=torch.float32, device=dev)
    os_ = torch.empty((B, T), dtype=torch.float32, device=dev)
    ov = torch.empty((B,), dtype=torch.int32, device=dev)
    res["decode_nms"] = timeit(
        lambda: _ffi.check(lib.ssd_decode_nms(_ffi.ptr(priors), _ffi.ptr(pd_d), _ffi.ptr(z_d), B, N, L, var, 1, 0.5, 0.5, T,
                                              0, _ffi.ptr(ob), _ffi.ptr(ol), _ffi.ptr(os_), _ffi.ptr(ov), _ffi.ptr(ws2),
                                              ws2.numel(), _ffi.stream())),
        B * N * (16 + 4 * L) + 16 * N + 24 * B * T)
    return {k: {"ms": round(v["ms"], 5), "MB": round(v["bytes"] / 1e6, 2), "frac": round(v["frac"], 4)} for k, v in res.items()}


def _other_inference_configs(steps=10, warmup=3):
    """Device-resident inference (forward + softmax + decode + NMS, one CUDA graph, uint8 input) of the other BASELINE
    configurations: configs[2]'s network in inference (SSD300-VGG16, 32 images per GPU) and configs[4] (SSD512-VGG16,
    128 images over 8 GPUs = 16 per GPU, 24 564 anchors; the SSD512 graph is an extension).  Every
    rank runs its own shard (no collective); times are the slowest rank's."""
    import torch
    from tf_ssd_b200 import dist_utils
    from tf_ssd_b200.models import ssd_vgg16
    from tf_ssd_b200.models.decoder import get_decoder_model
    from tf_ssd_b200.utils import bbox_utils, train_utils
    rank, local_rank, world = dist_utils.env_rank()
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for backbone, batch, key in (("vgg16", 32, "ssd300_vgg16_b32_per_gpu_inference"),
                                 ("vgg16_512", 16, "ssd512_vgg16_b16_per_gpu_inference")):
        hp = train_utils.get_hyper_params(backbone)
        hp["total_labels"] = 21
        model = ssd_vgg16.get_model(hp, seed=1234)
        priors = bbox_utils.generate_prior_boxes(hp["feature_map_shapes"], hp["aspect_ratios"])
        dm = get_decoder_model(model, priors, hp)
        st = dm._prepare(batch, 0)
        st["plan"].image_u8.copy_(torch.from_numpy(_make_images_u8(batch, hp["img_size"], seed=300 + rank)))
        for _ in range(warmup):
            dm.run_resident(batch, 0, u8=True)
        dist_utils.barrier()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in evs:
            flush.zero_()
            a.record(); dm.run_resident(batch, 0, u8=True); b.record()
        dist_utils.barrier()
        torch.cuda.synchronize()
        ms = dist_utils.max_over_ranks(float(sum(a.elapsed_time(b) for a, b in evs)), st["plan"].device) / steps
        flops = 2.0 * model.macs_per_image * batch
        out[key] = {"workload": f"{backbone} batch={batch}/GPU inference fwd+softmax+decode+NMS, {model.n_anchors} anchors, "
                                "uint8 input resident, random-init weights (uncalibrated: every anchor scores ~1/21)",
                    "value": world * batch / ms * 1e3, "unit": UNIT, "ms_per_step": ms, "n_gpus": world, "steps": steps,
                    "conv_tflops_per_gpu": flops / (ms * 1e-3) / 1e12, "launches": dm.launches_per_batch(batch)}
        del dm, model, st
        torch.cuda.empty_cache()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    from tf_ssd_b200 import dist_utils
    rank, local_rank, world = dist_utils.env_rank()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU restatement)")
    torch.cuda.set_device(local_rank)
    # pinned staging buffers on the GPU's NUMA node (multi-GPU runs only: at N = 1 this process also runs the CPU baseline
    # on all host cores)
    numa_bound = dist_utils.bind_to_gpu_numa(local_rank) if world > 1 else False
    dist_utils.init_from_env("nccl")

    from tf_ssd_b200 import _ffi, synth
    from tf_ssd_b200.models import ssd_mobilenet_v2
    from tf_ssd_b200.models.decoder import get_decoder_model
    from tf_ssd_b200.utils import bbox_utils
    _ffi.check_device()
    peaks = _peaks()
    hp = _hyper_params()
    model = ssd_mobilenet_v2.get_model(hp, seed=1234)
    cand = _calibrated_weights(model, hp)
    priors = bbox_utils.generate_prior_boxes(hp["feature_map_shapes"], hp["aspect_ratios"])
    dm = get_decoder_model(model, priors, hp)
    B, K, W = BATCH, args.steps, max(args.warmup, 3)
    S = hp["img_size"]

    # distinct synthetic batches per rank (inference shards are independent: no data-path collective)
    host_batches = [torch.from_numpy(_make_images_u8(B, S, seed=1000 + 17 * rank + i)).pin_memory() for i in range(4)]
    st = dm._prepare(B, 0)
    plan = st["plan"]
    st["enqueue_decode"] = lambda: _decode_only(dm, st, B)
    plan.image_u8.copy_(host_batches[0], non_blocking=False)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=plan.device)      # > 126 MB L2

    def barrier():
        dist_utils.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") -------------------------------
    for _ in range(W):
        dm.run_resident(B, 0, u8=True)
    barrier()
    if args.profile_one_step:               # ncu --profile-from-start off: exactly one step's launches, eagerly
        torch.cuda.profiler.start()
        st["enqueue"](True)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    sampler = ClockSampler(_physical_gpu_index(local_rank))
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for a, b in evs:
        flush.zero_()                       # L2 flush between timed steps, outside the event pair
        a.record()
        dm.run_resident(B, 0, u8=True)
        b.record()
    barrier()
    clocks = sampler.finish()
    dev_ms = float(sum(a.elapsed_time(b) for a, b in evs))
    valid = st["valid"].cpu().numpy()

    # ---- end-to-end through the public API ("e2e") --------------------------
    def host_iter(n):
        for i in range(n):
            yield host_batches[i % len(host_batches)]
    dm.predict(host_iter(W), steps=W)
    barrier()
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    res = dm.predict(host_iter(K), steps=K)
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = max(float(e0.elapsed_time(e1)), 1e3 * (time.perf_counter() - t0))
    assert res[0].shape == (K * B, 200, 4)
    barrier()

    dev_ms = dist_utils.max_over_ranks(dev_ms, plan.device)        # slowest rank
    e2e_ms = dist_utils.max_over_ranks(e2e_ms, plan.device)

    # ---- training steps of BASELINE configs[2] / configs[3] (every rank takes part: gradient all-reduce) -----
    training = None
    if not args.skip_train:
        del flush
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import train_bench
        training = {}
        for bb, key in (("mobilenet_v2", "ssd300_mobilenet_v2_b32_per_gpu"), ("vgg16", "ssd300_vgg16_b32_per_gpu")):
            train_bench.BACKBONE = bb
            if world == 1:
                try:                    # a side measurement must never cost the headline line (single process: no peer can hang)
                    training[key] = train_bench.measure(steps=args.train_steps, warmup=3, batch=32)
                except Exception as exc:  # noqa: BLE001
                    training[key] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            else:
                training[key] = train_bench.measure(steps=args.train_steps, warmup=3, batch=32)
            torch.cuda.empty_cache()

    other = None
    if not args.skip_other:
        try:
            other = _other_inference_configs()
        except Exception as exc:  # noqa: BLE001
            if world > 1:
                raise
            other = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if rank == 0:
        total_images = world * B * K
        value = total_images / (dev_ms / 1e3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": _config(world),
            "workload_check": {"nms_candidates_per_image": round(cand, 1), "valid_detections_mean": float(valid.mean()),
                               "images_per_step": B, "peaks": peaks["source"], "numa_bound": numa_bound},
            "clocks": clocks,
            "e2e": {"value": total_images / (e2e_ms / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": int(B * S * S * 3), "d2h_bytes_per_step": int(B * 200 * 6 * 4 + B * 4),
                    "api": "get_decoder_model(...).predict(uint8 host batches): pinned H2D + graph replay + D2H, 2 slots in flight"},
            "gpu_launches": K * dm.launches_per_batch(B),
        }
        if training is not None:
            line["training"] = training
        if other is not None:
            line["other_configs"] = other
        if world == 1:
            per = _profile_steps(dm, B)
            steps = plan.steps
            groups = {}
            for s, ms in zip(steps, per[:-1]):
                g = groups.setdefault(s.kind, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
                g["ms"] += ms; g["flops"] += s.flops; g["bytes"] += s.bytes; g["launches"] += 1
            groups["decode_nms"] = {"ms": float(per[-1]), "flops": 0.0, "launches": 2,
                                    "bytes": float(B * model.n_anchors * (16 + 4 * 21) + 16 * model.n_anchors + 24 * B * 200)}
            total = sum(g["ms"] for g in groups.values())
            top = max(groups, key=lambda k: groups[k]["ms"])
            g = groups[top]
            t_flop = g["flops"] / (peaks["tflops_burst"] * 1e12)
            t_mem = g["bytes"] / (peaks["hbm_gbs"] * 1e9)
            if t_flop > t_mem:
                ach = g["flops"] / (g["ms"] * 1e-3) / 1e12
                roof = {"bound": "tensor", "achieved": ach, "peak": peaks["tflops_burst"], "unit": "TFLOP/s",
                        "frac": ach / peaks["tflops_burst"]}
            else:
                ach = g["bytes"] / (g["ms"] * 1e-3) / 1e9
                roof = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"]}
            roof.update({"traffic": _ncu_traffic(top, g["launches"]), "algorithmic_bytes_per_launch": g["bytes"] / max(g["launches"], 1), "kernel": {"conv": "ssd_conv2d launches: conv_tcgen05_kernel (+ conv_splitk_reduce_kernel for the multibox heads)",
                                                     "irblock": "ssd_irblock launches (whole inverted-residual block: 1x1 expand -> depthwise 3x3 -> 1x1 project): irblock_mma_kernel "
                                    "(blocks 1-6) + irblock_mma_grouped_kernel (blocks 7-12, 14, 15; conv_irblock_tcgen05_kernel when PDL is off)",
                         "dw": "depthwise3x3_kernel", "dwproj": "conv_dwproj_tcgen05_kernel (fused depthwise 3x3 -> 1x1 projection)", "decode_nms": "nms_candidates+nms_image",
                         "chain": "conv_chain_kernel (the small-map tail + its heads as one cluster launch)", "stem": "stem_conv3x3s2_mma_kernel",
                         "stemblock": "stem_dwproj_kernel (Conv1 3x3 s2 -> depthwise 3x3 -> 1x1 projection from the image, one launch)"}.get(top, top),
                         "launches_per_step": g["launches"], "ms_per_step": g["ms"], "share_of_step": g["ms"] / total,
                         "algorithmic_bytes_per_step": g["bytes"], "flops_per_step": g["flops"], "peak_source": peaks["source"],
                         "by_kind_ms": {k: round(v["ms"], 4) for k, v in groups.items()}})
            grp = _ncu_group(top)
            if grp is not None and grp.get("l1tex_throughput_pct") is not None:
                # what actually bounds the block kernels: the shared-memory (L1/TEX) pipe, from the committed ncu capture
                roof["on_chip"] = {"l1tex_throughput_pct_of_peak": round(float(grp["l1tex_throughput_pct"]), 1),
                                   "source": grp.get("source"), "note": "time-weighted l1tex__throughput of the group's launches"}
            if top == "irblock":
                # context for the fraction above: the fused kernel's algorithmic bytes exclude the 6x expanded tensors it
                # keeps on chip; the layer-by-layer path (expand -> depthwise -> project as three launches) would move these
                unfused = 0.0
                for s_ in steps:
                    if s_.kind == "irblock":
                        m_ = s_.meta
                        cexp = int(m_["exp_w"].shape[0])
                        hw_in = int(m_["x"].shape[1]) * int(m_["x"].shape[2])
                        unfused += s_.bytes + 2.0 * 2.0 * B * cexp * (hw_in + m_["Ho"] * m_["Wo"])
                roof["layer_by_layer_equivalent"] = {
                    "bytes_per_step": unfused, "gbs": unfused / (g["ms"] * 1e-3) / 1e9,
                    "frac": unfused / (g["ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                    "note": "bytes the unfused expand / depthwise / project launches of the same blocks would move; the fused "
                            "kernels are bound by the shared-memory pipe (expanded patch written once and read by the nine "
                            "depthwise taps, ldmatrix operands), not by HBM or the tensor pipe: see on_chip"}
            line["roofline"] = roof
            worst = sorted(zip(per[:-1], steps), key=lambda t: -t[0])[:8]
            line["top_launches"] = [{"name": s.name, "kind": s.kind, "ms": round(float(ms), 4),
                                     "tflops": round(s.flops / (ms * 1e-3) / 1e12, 2), "gbs": round(s.bytes / (ms * 1e-3) / 1e9, 1)}
                                    for ms, s in worst]
            if not args.skip_box:
                try:
                    line["box_kernels"] = _box_kernel_rooflines(peaks, hp)
                except Exception as exc:  # noqa: BLE001
                    line["box_kernels"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            if not args.skip_cpu:
                from oracle import box_oracle as bo
                line["cpu_baseline"] = cpu_baseline(model.weights, hp, bo.prior_boxes(hp["feature_map_shapes"], hp["aspect_ratios"]))
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _decode_only(dm, st, B):
    from tf_ssd_b200 import _ffi
    dec, lib, plan, m = dm.decoder, _ffi.lib(), st["plan"], dm.base_model
    _ffi.check(lib.ssd_decode_nms(_ffi.ptr(dec._priors()), _ffi.ptr(plan.deltas), _ffi.ptr(plan.logits), B, m.n_anchors,
                                  m.total_labels, _ffi.f32_array(dec.variances), 1, dec.score_threshold, dec.iou_threshold,
                                  dec.max_total_size, 0, _ffi.ptr(st["boxes"]), _ffi.ptr(st["labels"]), _ffi.ptr(st["scores"]),
                                  _ffi.ptr(st["valid"]), _ffi.ptr(st["ws"]), st["ws"].numel(), _ffi.stream()), "ssd_decode_nms")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg (profiling runs)")
    ap.add_argument("--skip-box", action="store_true", help="omit the box-kernel stress rooflines")
    ap.add_argument("--profile-one-step", action="store_true",
                    help="warm up, run ONE eager step between cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)")
    ap.add_argument("--skip-train", action="store_true", help="omit the training-step measurements (configs[2], configs[3])")
    ap.add_argument("--train-steps", type=int, default=10)
    ap.add_argument("--skip-other", action="store_true", help="omit the VGG16 / SSD512 inference side measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
