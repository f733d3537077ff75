"""Timelines from the debug stamps of ssd_debug_trace (CTA 0 of the instrumented kernels), relative to the first stamp.

    python tools/trace_kernel.py dwproj B H W C Cout stride       # fused depthwise -> projection kernel
    python tools/trace_kernel.py nms                              # per-image NMS pass on the bench workload (B=32)
    python tools/trace_kernel.py nms-stress                       # ... on the stress shape (B=256, N=24564)
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tf_ssd_b200 import _ffi  # noqa: E402

ROLES = 8


def dump(buf, names, tags, limit=200):
    t = buf.cpu().numpy().astype(np.uint64).reshape(ROLES, 512)
    ev = []
    for role in range(ROLES):
        for v in t[role]:
            if v:
                ev.append((int(v >> np.uint64(8)), role, int(v & np.uint64(0xff))))
    ev.sort()
    if not ev:
        print("no stamps recorded")
        return
    t0 = ev[0][0]
    for ts, role, tag in ev[:limit]:
        print(f"{(ts - t0) / 1000.0:9.2f} us  {names.get(role, role):5s} {tags.get(role, {}).get(tag, tag)}")
    print(f"... {len(ev)} stamps, last at {(ev[-1][0] - t0) / 1000.0:.2f} us")


def trace_dwproj(B, H, W, Cc, Cout, stride):
    from tf_ssd_b200._ffi_conv import DwProjDesc
    lib = _ffi.lib()
    ph, pw = ((1, 1), (1, 1)) if stride == 1 else ((1 - H % 2, 1), (1 - W % 2, 1))
    Ho, Wo = (H + sum(ph) - 3) // stride + 1, (W + sum(pw) - 3) // stride + 1
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    x, wd, wp = r(B, H, W, Cc).half(), (r(3, 3, Cc) * 0.3).half(), (r(Cout, Cc) * 0.1).half()
    bd, bp = r(Cc), r(Cout)
    out = torch.empty(B, Ho, Wo, Cout, device="cuda", dtype=torch.float16)
    d = DwProjDesc()
    d.inp, d.dw_weight, d.dw_bias, d.proj_weight, d.proj_bias, d.residual, d.out = (x.data_ptr(), wd.data_ptr(), bd.data_ptr(),
                                                                                    wp.data_ptr(), bp.data_ptr(), None, out.data_ptr())
    d.B, d.H, d.W, d.C, d.Ho, d.Wo, d.Cout = B, H, W, Cc, Ho, Wo, Cout
    d.stride, d.pad_top, d.pad_left, d.dw_act, d.act = stride, ph[0], pw[0], 2, 0
    for _ in range(3):
        _ffi.check(lib.ssd_dwproj(C.byref(d), _ffi.stream()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _ffi.check(lib.ssd_dwproj(C.byref(d), _ffi.stream()))
    e1.record()
    torch.cuda.synchronize()
    print("kernel us:", e0.elapsed_time(e1) * 100)
    buf = torch.zeros(ROLES * 512, dtype=torch.int64, device="cuda")
    lib.ssd_debug_trace(C.c_void_p(buf.data_ptr()))
    _ffi.check(lib.ssd_dwproj(C.byref(d), _ffi.stream()))
    torch.cuda.synchronize()
    lib.ssd_debug_trace(None)
    dump(buf, {0: "tma", 1: "mma", 2: "dw", 3: "epi"},
         {0: {1: "patch_issue"}, 1: {1: "tile_go", 2: "a_full"}, 2: {1: "patch_full", 2: "dw_done", 3: "a_free", 4: "a_written", 5: "stored", 6: "fenced"},
          3: {1: "epi_wait", 2: "d_full", 3: "epi_done"}})


def trace_nms(stress):
    import bench
    lib = _ffi.lib()
    buf = torch.zeros(ROLES * 512, dtype=torch.int64, device="cuda")
    tags = {0: {1: "start", 2: "histogram", 3: "scatter", 4: "ranked", 5: "boxes_fetched", 6: "suppressed", 7: "merged", 8: "written", 9: "fetched+bases", 10: "zeroed"}}
    if stress:
        lib.ssd_debug_trace(C.c_void_p(buf.data_ptr()))
        out = bench._box_kernel_rooflines(bench._peaks(), bench._hyper_params(), iters=1, warmup=1)
        torch.cuda.synchronize()
        lib.ssd_debug_trace(None)
        print({k: v for k, v in out["kernels"].items() if k == "decode_nms"})
    else:
        from tf_ssd_b200.models import ssd_mobilenet_v2
        from tf_ssd_b200.models.decoder import get_decoder_model
        from tf_ssd_b200.utils import bbox_utils
        hp = bench._hyper_params()
        model = ssd_mobilenet_v2.get_model(hp, seed=1234)
        bench._calibrated_weights(model, hp)
        priors = bbox_utils.generate_prior_boxes(hp["feature_map_shapes"], hp["aspect_ratios"])
        dm = get_decoder_model(model, priors, hp)
        B = bench.BATCH
        st = dm._prepare(B, 0)
        st["plan"].image_u8.copy_(torch.from_numpy(bench._make_images_u8(B, hp["img_size"], seed=1000)))
        for _ in range(3):
            dm.run_resident(B, 0, u8=True)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record()
        for _ in range(10):
            bench._decode_only(dm, st, B)
        e[1].record()
        torch.cuda.synchronize()
        print("decode_nms us per call:", e[0].elapsed_time(e[1]) * 100)
        lib.ssd_debug_trace(C.c_void_p(buf.data_ptr()))
        bench._decode_only(dm, st, B)
        torch.cuda.synchronize()
        lib.ssd_debug_trace(None)
        labels = st["labels"].cpu().numpy()
        valid = st["valid"].cpu().numpy()
        hist = np.bincount(labels[0, :valid[0]].astype(np.int64), minlength=21)
        print("image 0: valid", int(valid[0]), "detections per class", hist.tolist())
    dump(buf, {0: "nms"}, tags)


def trace_chain():
    import bench
    from tf_ssd_b200.models import ssd_mobilenet_v2
    lib = _ffi.lib()
    hp = bench._hyper_params()
    model = ssd_mobilenet_v2.get_model(hp, seed=1234)
    plan = model.plan(bench.BATCH)
    plan.image_u8.copy_(torch.from_numpy(bench._make_images_u8(bench.BATCH, hp["img_size"], seed=1000)))
    for _ in range(3):
        plan.run(u8=True)
    torch.cuda.synchronize()
    idx = [i for i, s in enumerate(plan.steps) if s.kind == "chain"][0]
    print([(n, ph) for (n, _), ph in zip(plan.steps[idx].meta["layers"], plan.steps[idx].meta["phases"])])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        plan.run(idx, idx + 1, u8=True, parallel=False)
    e1.record()
    torch.cuda.synchronize()
    print("chain us per call:", e0.elapsed_time(e1) * 100)
    buf = torch.zeros(ROLES * 512, dtype=torch.int64, device="cuda")
    lib.ssd_debug_trace(C.c_void_p(buf.data_ptr()))
    plan.run(idx, idx + 1, u8=True, parallel=False)
    torch.cuda.synchronize()
    lib.ssd_debug_trace(None)
    dump(buf, {0: "chain", 1: "warp0"}, {0: {1: "layer", 2: "synced", 3: "staged", 4: "computed", 5: "end"}, 1: {1: "unit", 2: "tap"}}, limit=80)


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "chain":
        trace_chain()
        sys.exit(0)
    if what == "dwproj":
        trace_dwproj(*map(int, sys.argv[2:8]))
    elif what == "nms":
        trace_nms(False)
    else:
        trace_nms(True)
