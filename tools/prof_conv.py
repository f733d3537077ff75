"""Launch one ssd_conv2d shape a few times (for `ncu --set full`).

    python tools/prof_conv.py B H W Cin Cout k [iters]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tf_ssd_b200 import _ffi  # noqa: E402
from tf_ssd_b200._ffi_conv import ConvDesc  # noqa: E402

B, H, W, Cin, Cout, k = [int(v) for v in sys.argv[1:7]]
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 3
dev = torch.device("cuda")
x = torch.randn((B, H, W, Cin), device=dev).half()
w = (torch.randn((Cout, k, k, Cin), device=dev) / (k * k * Cin) ** 0.5).half()
bias = torch.randn(Cout, device=dev)
out = torch.empty((B, H, W, Cout), dtype=torch.float16, device=dev)
d = ConvDesc()
d.inp, d.weight, d.bias, d.out0 = x.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr()
d.B, d.H, d.W, d.Cin, d.Ho, d.Wo, d.Cout = B, H, W, Cin, H, W, Cout
d.KH = d.KW = k
d.stride = d.dilation = 1
d.pad_top = d.pad_left = k // 2
d.act, d.out_f32, d.split = 2, 0, Cout
d.img_stride0, d.pix_stride0 = H * W * Cout, Cout
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
evs = []
for _ in range(iters):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _ffi.check(_ffi.lib().ssd_conv2d(C.byref(d), _ffi.stream()), "ssd_conv2d")
    b.record()
    evs.append((a, b))
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for a, b in evs]
nbytes = (B * H * W * (Cin + Cout) + Cout * k * k * Cin) * 2
print("ms", ms, "GB/s", nbytes / min(ms) / 1e6, "TFLOP/s", 2 * B * H * W * k * k * Cin * Cout / min(ms) / 1e9)
