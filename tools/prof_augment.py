"""Device augmentation throughput: ``ssd_augment_batch`` on a 300x300 batch with plans from the product's
sampler.  Algorithmic bytes = one read + one write of the float32 batch (24 B / pixel); images that expand read the
input once more for the canvas mean, images that change contrast read + write the output once more.

    python tools/prof_augment.py [--batch 256] [--size 300]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from tf_ssd_b200 import _ffi, augmentation as aug, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--size", type=int, default=300)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    B, S = args.batch, args.size
    rng = np.random.default_rng(0)
    img = torch.rand((B, S, S, 3), device="cuda")
    gt, _ = synth.make_ground_truth(B, padded=8, seed=3)
    boxes = torch.from_numpy(gt).cuda()
    draws = aug.RandomDraws(11)
    plans = [aug.make_plan(S, S, gt[i], draws) for i in range(B)]
    d_plans = torch.from_numpy(aug.pack_plans(plans, S, S)).cuda()
    out = torch.empty_like(img)
    lib = _ffi.lib()
    ws = _ffi.workspace(int(lib.ssd_augment_workspace_bytes(B, S, S, S, S)))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def run():
        _ffi.check(lib.ssd_augment_batch(_ffi.ptr(img), _ffi.ptr(out), _ffi.ptr(boxes), B, S, S, S, S, int(boxes.shape[1]),
                                         _ffi.ptr(d_plans), _ffi.ptr(ws), ws.numel() * ws.element_size(), _ffi.stream()), "augment")
    for _ in range(3):
        run()
    times = []
    for _ in range(args.iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    ms = float(np.median(times))
    px = B * S * S
    n_expand = sum(1 for p in plans if p["patch"] and p["patch"]["expand"])
    n_contrast = sum(1 for p in plans if p["contrast"] is not None)
    minimal = 24.0 * px
    moved = minimal + 12.0 * S * S * n_expand + 24.0 * S * S * n_contrast
    peak = 6538.6
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    print(json.dumps({"kernel": "ssd_augment_batch", "batch": B, "size": S, "ms": ms, "images_per_s": B / ms * 1e3,
                      "algorithmic_gbs": minimal / ms / 1e6, "with_dependent_passes_gbs": moved / ms / 1e6,
                      "expand_images": n_expand, "contrast_images": n_contrast, "hbm_peak_gbs": peak,
                      "frac_of_peak": moved / ms / 1e6 / peak}))


if __name__ == "__main__":
    main()
