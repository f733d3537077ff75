"""Device-resident inference throughput (forward + softmax + decode + NMS, one CUDA graph) of any backbone / batch:
BASELINE.json configs[4] (SSD512-VGG16, 16 images per GPU) and SSD300-VGG16 next to the headline MobileNetV2 config.

    python tools/infer_bench.py --backbone vgg16_512 --batch 16 [--steps 30]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backbone", default="vgg16_512", choices=["mobilenet_v2", "vgg16", "vgg16_512"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--profile-one-step", action="store_true", help="ncu --profile-from-start off: one eager step between profiler start/stop")
    a = ap.parse_args()
    import torch
    from tf_ssd_b200 import synth
    from tf_ssd_b200.models import ssd_mobilenet_v2, ssd_vgg16
    from tf_ssd_b200.models.decoder import get_decoder_model
    from tf_ssd_b200.utils import bbox_utils, train_utils
    torch.cuda.set_device(0)
    hp = train_utils.get_hyper_params(a.backbone)
    hp["total_labels"] = 21
    mod = ssd_mobilenet_v2 if a.backbone == "mobilenet_v2" else ssd_vgg16
    model = mod.get_model(hp, seed=1234)
    priors = bbox_utils.generate_prior_boxes(hp["feature_map_shapes"], hp["aspect_ratios"])
    dm = get_decoder_model(model, priors, hp)
    st = dm._prepare(a.batch, 0)
    st["plan"].image.copy_(torch.from_numpy(synth.make_images(a.batch, hp["img_size"], seed=3)))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(5):
        dm.run_resident(a.batch, 0)
    torch.cuda.synchronize()
    if a.profile_one_step:
        torch.cuda.profiler.start()
        st["plan"].run()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    for e0, e1 in evs:
        flush.zero_()
        e0.record(); dm.run_resident(a.batch, 0); e1.record()
    torch.cuda.synchronize()
    ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in evs]))
    flops = 2.0 * model.macs_per_image * a.batch
    print(json.dumps({"backbone": a.backbone, "batch": a.batch, "anchors": model.n_anchors, "ms_per_step": ms,
                      "images_per_s": a.batch / ms * 1e3, "conv_tflops": flops / (ms * 1e-3) / 1e12,
                      "launches": dm.launches_per_batch(a.batch)}))


if __name__ == "__main__":
    main()
