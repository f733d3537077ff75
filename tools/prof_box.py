"""Profiling driver: launches every box-side kernel once at the stress shape
(B=256, N=24564, G=42, L=21) after one warm-up, for `ncu --set full`.

    ncu --set full --clock-control none --import-source on -k regex:'iou_map|match_encode|loss_|nms_' \
        -o gpurun_out/box python tools/prof_box.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bench  # noqa: E402

if __name__ == "__main__":
    import torch
    torch.cuda.set_device(0)
    hp = bench._hyper_params()
    out = bench._box_kernel_rooflines(bench._peaks(), hp, iters=int(os.environ.get("ITERS", "1")),
                                      warmup=int(os.environ.get("WARMUP", "1")))
    print(json.dumps(out))
