"""Experiment: S concurrent sub-batches (B/S images each, own stream + CUDA graph) vs one batch-B graph."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from tf_ssd_b200 import synth
from tf_ssd_b200.models import ssd_mobilenet_v2
from tf_ssd_b200.models.decoder import get_decoder_model
from tf_ssd_b200.utils import bbox_utils

torch.cuda.set_device(0)
hp = bench._hyper_params()
model = ssd_mobilenet_v2.get_model(hp, seed=1234)
bench._calibrated_weights(model, hp)
priors = bbox_utils.generate_prior_boxes(hp["feature_map_shapes"], hp["aspect_ratios"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for S in (1, 2, 4, 8):
    b = 32 // S
    dms = [get_decoder_model(model, priors, hp) for _ in range(S)]
    streams = [torch.cuda.Stream() for _ in range(S)]
    sts = [dm._prepare(b, 0) for dm in dms]
    for st in sts:
        st["plan"].image.copy_(torch.from_numpy(synth.make_images(b, 300, seed=3)))
    torch.cuda.synchronize()
    def step():
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(main)
        for st, s in zip(sts, streams):
            s.wait_event(ev)
            with torch.cuda.stream(s):
                dm._graph(st, False).replay()
            main.wait_stream(s)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print(f"S={S} sub-batch={b}: {ms:.3f} ms per 32 images -> {32 / ms * 1e3:.0f} img/s", flush=True)
