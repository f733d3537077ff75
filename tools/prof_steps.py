"""Per-launch CUDA-event timing of every step of an inference plan (L2 flushed before each pass).

    SSD_B200_FUSE_IR=0 python tools/prof_steps.py [--backbone mobilenet_v2] [--batch 32] [--iters 20]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backbone", default="mobilenet_v2")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    import torch
    from tf_ssd_b200.models import ssd_mobilenet_v2, ssd_vgg16
    from tf_ssd_b200.utils import train_utils
    hp = train_utils.get_hyper_params(a.backbone)
    hp["total_labels"] = 21
    mod = ssd_mobilenet_v2 if a.backbone == "mobilenet_v2" else ssd_vgg16
    m = mod.get_model(hp, seed=1234)
    plan = m.plan(a.batch)
    plan.image_u8.copy_(torch.from_numpy(np.random.default_rng(0).integers(0, 256, tuple(plan.image_u8.shape), dtype=np.uint8)))
    n = plan.n_launches
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    acc = np.zeros(n)
    for it in range(a.iters + 2):
        flush.zero_()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        evs[0].record()
        for i in range(n):
            plan.run(i, i + 1, u8=True, parallel=False)
            evs[i + 1].record()
        torch.cuda.synchronize()
        if it >= 2:
            acc += np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(n)])
    acc /= a.iters
    rows = [{"name": s.name, "kind": s.kind, "us": round(float(t) * 1e3, 1)} for s, t in zip(plan.steps, acc)]
    by_kind = {}
    for r in rows:
        by_kind[r["kind"]] = round(by_kind.get(r["kind"], 0.0) + r["us"], 1)
    print(json.dumps({"fuse_ir": os.environ.get("SSD_B200_FUSE_IR", "1"), "total_us": round(float(acc.sum()) * 1e3, 1),
                      "by_kind_us": by_kind, "steps": rows}))


if __name__ == "__main__":
    main()
