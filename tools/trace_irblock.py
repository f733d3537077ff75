"""Timeline of one CTA of ssd_irblock (debug): per-role globaltimer stamps, printed relative to the first one.

    python tools/trace_irblock.py B H W Cin Cexp Cout stride
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tf_ssd_b200 import _ffi  # noqa: E402
from tf_ssd_b200._ffi_conv import IrBlockDesc  # noqa: E402


def main():
    B, H, W, Cin, Cexp, Cout, stride = map(int, sys.argv[1:8])
    lib = _ffi.lib()
    ph, pw = ((1, 1), (1, 1)) if stride == 1 else ((1 - H % 2, 1), (1 - W % 2, 1))
    Ho, Wo = (H + sum(ph) - 3) // stride + 1, (W + sum(pw) - 3) // stride + 1
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    x, we, wd, wp = r(B, H, W, Cin).half(), (r(Cexp, Cin) * 0.2).half(), (r(3, 3, Cexp) * 0.3).half(), (r(Cout, Cexp) * 0.1).half()
    be, bd, bp = r(Cexp), r(Cexp), r(Cout)
    out = torch.empty(B, Ho, Wo, Cout, device="cuda", dtype=torch.float16)
    d = IrBlockDesc()
    d.inp, d.exp_weight, d.exp_bias, d.dw_weight, d.dw_bias = x.data_ptr(), we.data_ptr(), be.data_ptr(), wd.data_ptr(), bd.data_ptr()
    d.proj_weight, d.proj_bias, d.residual, d.out = wp.data_ptr(), bp.data_ptr(), None, out.data_ptr()
    d.B, d.H, d.W, d.Cin, d.Cexp, d.Ho, d.Wo, d.Cout = B, H, W, Cin, Cexp, Ho, Wo, Cout
    d.stride, d.pad_top, d.pad_left, d.exp_act, d.dw_act, d.act = stride, ph[0], pw[0], 2, 2, 0
    for _ in range(3):
        _ffi.check(lib.ssd_irblock(C.byref(d), _ffi.stream()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _ffi.check(lib.ssd_irblock(C.byref(d), _ffi.stream()))
    e1.record()
    torch.cuda.synchronize()
    print("kernel us:", e0.elapsed_time(e1) * 100)
    buf = torch.zeros(5 * 512, dtype=torch.int64, device="cuda")
    lib.ssd_irblock_trace(C.c_void_p(buf.data_ptr()))
    _ffi.check(lib.ssd_irblock(C.byref(d), _ffi.stream()))
    torch.cuda.synchronize()
    lib.ssd_irblock_trace(None)
    t = buf.cpu().numpy().astype(np.uint64).reshape(5, 512)
    ev = []
    names = {0: "tma", 1: "mma", 2: "dw", 3: "mid", 4: "epi"}
    tags = {0: {1: "patch_issued", 2: "wexp_issued", 3: "dwf_issued", 4: "wproj_issued"}, 1: {1: "expand_go", 2: "project_go", 3: "wexp_arrived", 4: "d1_free"},
            2: {1: "ep_full", 2: "dw_done", 3: "a2_written"}, 3: {1: "d1_full", 2: "ep_free", 3: "mid_done"}, 4: {4: "epi_wait", 5: "d2_full", 6: "epi_done"}}
    for role in range(5):
        for v in t[role]:
            if v:
                ev.append((int(v >> np.uint64(8)), role, int(v & np.uint64(0xff))))
    ev.sort()
    t0 = ev[0][0]
    for ts, role, tag in ev[:int(os.environ.get('TRACE_N', '160'))]:
        print(f"{(ts - t0) / 1000.0:9.2f} us  {names[role]:4s} {tags[role].get(tag, tag)}")


if __name__ == "__main__":
    main()
