"""Parity of the forward engine (C-ABI conv / depthwise / pool / l2norm kernels and the
two SSD graphs) against the torch-CPU oracle.

Tolerances: the product computes with fp16 storage and fp32 accumulation
(BASELINE.json config 2: "fp16 convs / fp32 boxes").  Single launches are
checked at 2e-3 of the tensor's max magnitude (one fp16 output rounding) against
torch-CPU fp32 on the same fp16 operands; whole graphs are checked relative to
the oracle's own fp16-storage error (see ``test_forward_parity``)."""

import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import box_oracle as bo
from oracle import net_oracle as no

pytestmark = pytest.mark.gpu

TOL_SIM = 5e-3
TOL_FP32 = 3e-2


def _randomise(model, seed):
    """Non-trivial BatchNorm statistics, biases and L2-norm scales (the initialiser leaves them at identity)."""
    rng = np.random.default_rng(seed)
    w = {}
    for k, v in model.weights.items():
        if k.endswith("/bias") or k.endswith("/beta") or k.endswith("/moving_mean"):
            w[k] = rng.normal(0, 0.1, v.shape).astype(np.float32)
        elif k.endswith("/gamma") or k.endswith("/moving_variance"):
            w[k] = rng.uniform(0.5, 1.5, v.shape).astype(np.float32)
        elif k.endswith("/scale"):
            w[k] = rng.uniform(10, 30, v.shape).astype(np.float32)
    model.set_weights(w)


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-6))


def _conv_case(B, H, W, Cin, Cout, k, stride, dil, pads, act, residual, seed, split=None, out_f32=False):
    from tf_ssd_b200 import _ffi
    from tf_ssd_b200._ffi_conv import ConvDesc
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, H, W, Cin)).astype(np.float16)
    w = (rng.standard_normal((Cout, k, k, Cin)) / np.sqrt(k * k * Cin)).astype(np.float16)
    bias = rng.standard_normal(Cout).astype(np.float32)
    (pt, pb), (pl, pr) = pads
    Ho = (H + pt + pb - ((k - 1) * dil + 1)) // stride + 1
    Wo = (W + pl + pr - ((k - 1) * dil + 1)) // stride + 1
    res = rng.standard_normal((B, Ho, Wo, Cout)).astype(np.float16) if residual else None
    # reference: torch CPU fp32 conv on the fp16-rounded operands
    xt = torch.from_numpy(x.astype(np.float32)).permute(0, 3, 1, 2)
    wt = torch.from_numpy(w.astype(np.float32)).permute(0, 3, 1, 2)
    y = F.conv2d(F.pad(xt, (pl, pr, pt, pb)), wt, torch.from_numpy(bias), stride=stride, dilation=dil)
    if act == 1:
        y = torch.relu(y)
    elif act == 2:
        y = torch.clamp(y, 0, 6)
    y = y.permute(0, 2, 3, 1).numpy()
    if res is not None:
        y = y + res.astype(np.float32)

    dev = torch.device("cuda")
    xd, wd, bd = torch.from_numpy(x).to(dev), torch.from_numpy(w).to(dev), torch.from_numpy(bias).to(dev)
    rd = torch.from_numpy(res).to(dev) if res is not None else None
    odt = torch.float32 if out_f32 else torch.float16
    sp = Cout if split is None else split
    out0 = torch.zeros((B, Ho, Wo, sp), dtype=odt, device=dev)
    out1 = torch.zeros((B, Ho, Wo, max(Cout - sp, 1)), dtype=odt, device=dev)
    d = ConvDesc()
    d.inp, d.weight, d.bias = xd.data_ptr(), wd.data_ptr(), bd.data_ptr()
    d.residual = rd.data_ptr() if rd is not None else None
    d.out0, d.out1 = out0.data_ptr(), out1.data_ptr()
    d.B, d.H, d.W, d.Cin, d.Ho, d.Wo, d.Cout = B, H, W, Cin, Ho, Wo, Cout
    d.KH = d.KW = k
    d.stride, d.dilation, d.pad_top, d.pad_left, d.act, d.out_f32, d.split = stride, dil, pt, pl, act, int(out_f32), sp
    d.img_stride0, d.pix_stride0 = Ho * Wo * sp, sp
    d.img_stride1, d.pix_stride1 = Ho * Wo * (Cout - sp), Cout - sp
    _ffi.check(_ffi.lib().ssd_conv2d(C.byref(d), _ffi.stream()), "ssd_conv2d")
    torch.cuda.synchronize()
    got = out0.float().cpu().numpy()
    if sp < Cout:
        got = np.concatenate([got, out1.float().cpu().numpy()], axis=-1)
    return got, y


@pytest.mark.parametrize("case", [
    # B, H, W, Cin, Cout, k, stride, dil, pads, act, residual
    (2, 19, 19, 64, 384, 1, 1, 1, ((0, 0), (0, 0)), 2, False),      # MobileNetV2 expand
    (2, 19, 19, 384, 64, 1, 1, 1, ((0, 0), (0, 0)), 0, True),       # project + residual add
    (1, 150, 150, 32, 16, 1, 1, 1, ((0, 0), (0, 0)), 0, False),     # thin projection, many rows
    (3, 10, 10, 1280, 256, 1, 1, 1, ((0, 0), (0, 0)), 1, False),    # extra1_1
    (2, 10, 10, 256, 512, 3, 2, 1, ((0, 1), (0, 1)), 1, False),     # extra1_2, SAME stride 2 even input
    (2, 5, 5, 128, 256, 3, 2, 1, ((1, 1), (1, 1)), 1, False),       # extra2_2, odd input
    (1, 31, 29, 8, 32, 3, 2, 1, ((0, 1), (1, 1)), 2, False),        # stem-like, Cin padded to 8
    (1, 19, 19, 64, 96, 3, 1, 6, ((6, 6), (6, 6)), 1, False),       # conv6-like dilation 6
    (1, 5, 5, 128, 256, 3, 1, 1, ((0, 0), (0, 0)), 1, False),       # conv10_2 VALID
    (2, 38, 38, 128, 128, 3, 1, 1, ((1, 1), (1, 1)), 1, False),     # VGG body
    (1, 7, 9, 24, 40, 1, 1, 1, ((0, 0), (0, 0)), 0, False),         # ragged K (24) and N (40)
])
def test_conv2d_against_torch(case):
    got, ref = _conv_case(*case, seed=hash(case) % 1000)
    assert got.shape == ref.shape
    assert _rel(got, ref) < 2e-3                       # fp16 output rounding only


@pytest.mark.parametrize("case", [
    # B, H, W, Cin, Cout, k, stride, dil, pads, act, residual
    (2, 38, 38, 128, 256, 3, 1, 1, ((1, 1), (1, 1)), 1, False),     # VGG body: 23 M tiles (odd: phantom tile), one N tile
    (3, 19, 19, 512, 512, 3, 1, 1, ((1, 1), (1, 1)), 1, False),     # two N tiles of 256
    (2, 19, 19, 512, 1024, 3, 1, 6, ((6, 6), (6, 6)), 1, False),    # conv6: dilation 6
    (4, 19, 19, 1024, 128, 1, 1, 1, ((0, 0), (0, 0)), 2, True),     # 1x1 (2-D operand maps), BN = 128, residual
    (2, 20, 20, 256, 256, 3, 2, 1, ((0, 1), (0, 1)), 0, False),     # stride 2 through element strides
])
def test_conv2d_cta_pair_against_torch(case):
    """The cta_group::2 kernel (two CTAs, one M = 256 tcgen05.mma, half of the weight tile per CTA) on shapes small
    enough for a CPU check: forced through the test hook, since the automatic choice needs a tile per SM."""
    from tf_ssd_b200 import _ffi
    lib = _ffi.lib()
    B, H, W, Cin, Cout, k, stride, dil, pads, act, use_res = case
    _ffi.check(lib.ssd_debug_pair_mode(1))
    try:
        got, ref = _conv_case(B, H, W, Cin, Cout, k, stride, dil, pads, act, use_res, seed=3)
    finally:
        _ffi.check(lib.ssd_debug_pair_mode(-1))
    assert _rel(got, ref) < 3e-3, _rel(got, ref)
    _ffi.check(lib.ssd_debug_pair_mode(0))
    try:
        single, _ = _conv_case(B, H, W, Cin, Cout, k, stride, dil, pads, act, use_res, seed=3)
    finally:
        _ffi.check(lib.ssd_debug_pair_mode(-1))
    assert np.array_equal(got, single), "the pair kernel must reproduce the single-CTA kernel bit for bit (same K order)"


def test_conv2d_head_split_fp32():
    # head-style: two fp32 segments (A*L = 84 label channels, A*4 = 16 box channels)
    got, ref = _conv_case(2, 10, 10, 64, 100, 3, 1, 1, ((1, 1), (1, 1)), 0, False, seed=5, split=84, out_f32=True)
    assert _rel(got, ref) < 1e-4


def test_conv2d_rejects_bad_arguments():
    from tf_ssd_b200 import _ffi
    from tf_ssd_b200._ffi_conv import ConvDesc
    d = ConvDesc()
    assert _ffi.lib().ssd_conv2d(C.byref(d), None) == -1            # SSD_ERR_NULL
    assert b"NULL" in _ffi.lib().ssd_last_error()


@pytest.mark.parametrize("stride,H,W,C", [(1, 19, 19, 384), (2, 38, 38, 192), (2, 75, 75, 144), (1, 7, 5, 8), (2, 150, 150, 96)])
def test_depthwise_against_torch(stride, H, W, C):
    from tf_ssd_b200 import _ffi
    from tf_ssd_b200.models.engine import _out_size, _resolve_pads
    rng = np.random.default_rng(C + H)
    B = 2
    x = rng.standard_normal((B, H, W, C)).astype(np.float16)
    w = (rng.standard_normal((3, 3, C)) / 3).astype(np.float16)
    bias = rng.standard_normal(C).astype(np.float32)
    ph, pw = _resolve_pads(H, W, 3, stride, 1, "same" if stride == 1 else "correct")
    assert (ph, pw) == ((no.same_pad(H, 3, 1), no.same_pad(W, 3, 1)) if stride == 1 else (no.correct_pad(H), no.correct_pad(W)))
    Ho, Wo = _out_size(H, 3, stride, 1, ph), _out_size(W, 3, stride, 1, pw)
    xt = torch.from_numpy(x.astype(np.float32)).permute(0, 3, 1, 2)
    wt = torch.from_numpy(w.astype(np.float32)).permute(2, 0, 1).unsqueeze(1)
    y = F.conv2d(F.pad(xt, (pw[0], pw[1], ph[0], ph[1])), wt, torch.from_numpy(bias), stride=stride, groups=C)
    y = torch.clamp(y, 0, 6).permute(0, 2, 3, 1).numpy()
    dev = torch.device("cuda")
    xd, wd, bd = torch.from_numpy(x).to(dev), torch.from_numpy(w).to(dev), torch.from_numpy(bias).to(dev)
    out = torch.zeros((B, Ho, Wo, C), dtype=torch.float16, device=dev)
    _ffi.check(_ffi.lib().ssd_depthwise3x3(_ffi.ptr(xd), _ffi.ptr(wd), _ffi.ptr(bd), _ffi.ptr(out), B, H, W, C, Ho, Wo,
                                           stride, ph[0], pw[0], 2, _ffi.stream()), "ssd_depthwise3x3")
    assert _rel(out.float().cpu().numpy(), y) < 2e-3


@pytest.mark.parametrize("k,s,H", [(2, 2, 75), (2, 2, 38), (3, 1, 19)])
def test_maxpool_and_l2norm(k, s, H):
    from tf_ssd_b200 import _ffi
    rng = np.random.default_rng(H)
    B, C_ = 2, 64
    x = rng.standard_normal((B, H, H, C_)).astype(np.float16)
    pads = no.same_pad(H, k, s)
    Ho = -(-H // s)
    xt = torch.from_numpy(x.astype(np.float32)).permute(0, 3, 1, 2)
    y = F.max_pool2d(F.pad(xt, (pads[0], pads[1], pads[0], pads[1]), value=float("-inf")), k, s).permute(0, 2, 3, 1).numpy()
    dev = torch.device("cuda")
    xd = torch.from_numpy(x).to(dev)
    out = torch.zeros((B, Ho, Ho, C_), dtype=torch.float16, device=dev)
    _ffi.check(_ffi.lib().ssd_maxpool(_ffi.ptr(xd), _ffi.ptr(out), B, H, H, C_, Ho, Ho, k, s, pads[0], pads[0], _ffi.stream()))
    assert np.array_equal(out.float().cpu().numpy(), y)             # max is exact

    scale = rng.uniform(10, 30, C_).astype(np.float32)
    xf = x.astype(np.float32)
    ref = xf / np.sqrt(np.maximum((xf * xf).sum(-1, keepdims=True), 1e-12)) * scale
    sd = torch.from_numpy(scale).to(dev)
    out2 = torch.zeros_like(xd)
    _ffi.check(_ffi.lib().ssd_l2norm(_ffi.ptr(xd), _ffi.ptr(sd), _ffi.ptr(out2), B * H * H, C_, _ffi.stream()))
    assert _rel(out2.float().cpu().numpy(), ref) < 2e-3


@pytest.mark.parametrize("case", [
    # B, H, C, Cout, stride, residual
    (2, 19, 96, 24, 1, False), (2, 19, 144, 24, 1, True), (1, 20, 144, 32, 2, False), (3, 7, 32, 16, 1, False),
    (2, 10, 960, 160, 1, True), (1, 33, 8, 8, 2, False), (2, 5, 64, 256, 1, False), (1, 75, 24, 16, 1, True),
    (2, 38, 192, 64, 2, False),
])
def test_dwproj_fused_against_torch(case):
    """ssd_dwproj (depthwise 3x3 + bias + ReLU6 -> 1x1 + bias (+ residual)) on shapes beyond the MobileNetV2 ones:
    channel counts that are not multiples of 64, box widths that are / are not multiples of 4 (sliding-window and
    per-pixel depthwise paths), partial tiles, both strides."""
    import ctypes as C
    from tf_ssd_b200 import _ffi
    from tf_ssd_b200._ffi_conv import DwProjDesc
    B, H, Cc, Cout, stride, with_res = case
    rng = np.random.default_rng(sum(case[:5]))
    pads = no.same_pad(H, 3, 1) if stride == 1 else no.correct_pad(H)
    Ho = (H + pads[0] + pads[1] - 3) // stride + 1
    x = rng.standard_normal((B, H, H, Cc)).astype(np.float16)
    wd = (rng.standard_normal((3, 3, Cc)) / 3).astype(np.float16)
    bd = rng.uniform(-0.2, 0.5, Cc).astype(np.float32)
    wp = (rng.standard_normal((Cout, Cc)) / np.sqrt(Cc)).astype(np.float16)
    bp = rng.uniform(-0.2, 0.2, Cout).astype(np.float32)
    res = rng.standard_normal((B, Ho, Ho, Cout)).astype(np.float16) if with_res else None
    xt = torch.from_numpy(x.astype(np.float32)).permute(0, 3, 1, 2)
    h = F.conv2d(F.pad(xt, (pads[0], pads[1], pads[0], pads[1])), torch.from_numpy(wd.astype(np.float32)).permute(2, 0, 1).unsqueeze(1),
                 torch.from_numpy(bd), stride=stride, groups=Cc)
    h = torch.clamp(h, 0, 6).half().float()
    y = F.conv2d(h, torch.from_numpy(wp.astype(np.float32)).view(Cout, Cc, 1, 1), torch.from_numpy(bp)).permute(0, 2, 3, 1)
    if with_res:
        y = y + torch.from_numpy(res.astype(np.float32))
    dev = torch.device("cuda")
    t = lambda a: torch.from_numpy(a).to(dev)
    xd, wdd, bdd, wpd, bpd = t(x), t(wd), t(bd), t(wp), t(bp)
    rd = t(res) if with_res else None
    out = torch.full((B, Ho, Ho, Cout), 7.0, dtype=torch.float16, device=dev)
    d = DwProjDesc()
    d.inp, d.dw_weight, d.dw_bias, d.proj_weight, d.proj_bias = xd.data_ptr(), wdd.data_ptr(), bdd.data_ptr(), wpd.data_ptr(), bpd.data_ptr()
    d.residual = rd.data_ptr() if with_res else None
    d.out = out.data_ptr()
    d.B, d.H, d.W, d.C, d.Ho, d.Wo, d.Cout = B, H, H, Cc, Ho, Ho, Cout
    d.stride, d.pad_top, d.pad_left, d.dw_act, d.act = stride, pads[0], pads[0], 2, 0
    lib = _ffi.lib()
    assert lib.ssd_dwproj_supported(C.byref(d)) == 1
    _ffi.check(lib.ssd_dwproj(C.byref(d), _ffi.stream()), "ssd_dwproj")
    torch.cuda.synchronize()
    assert _rel(out.float().cpu().numpy(), y.numpy()) < 3e-3, _rel(out.float().cpu().numpy(), y.numpy())
    d.Cout = 320                                                    # more than one N tile: must be refused, not mis-computed
    assert lib.ssd_dwproj_supported(C.byref(d)) == 0 and lib.ssd_dwproj(C.byref(d), _ffi.stream()) < 0


def _model(backbone, seed=3):
    from tf_ssd_b200.models import ssd_mobilenet_v2, ssd_vgg16
    from tf_ssd_b200.utils import train_utils
    hp = train_utils.get_hyper_params(backbone)
    hp["total_labels"] = 21
    mod = ssd_mobilenet_v2 if backbone == "mobilenet_v2" else ssd_vgg16
    m = mod.get_model(hp, seed=seed)
    _randomise(m, seed + 1)
    return m, hp


def _rms(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-12))


def _act(y, act):
    return torch.relu(y) if act == 1 else torch.clamp(y, 0, 6) if act == 2 else y


def _irblock_reference(x, we, be, exp_act, wd, bd, stride, ph, pw, dw_act, wp, bp, act, res):
    """expand 1x1 -> depthwise 3x3 -> project 1x1 on NHWC float tensors, with the fp16 roundings of the fused kernel:
    the expanded activation and the depthwise output are rounded to fp16 (they are tensor-core operands)."""
    xn = x.permute(0, 3, 1, 2)
    h = _act(F.conv2d(xn, we.permute(0, 3, 1, 2), be), exp_act).half().float()
    (pt, pb), (pl, pr) = ph, pw
    h = F.conv2d(F.pad(h, (pl, pr, pt, pb)), wd.permute(2, 0, 1).unsqueeze(1), bd, stride=stride, groups=wd.shape[2])
    h = _act(h, dw_act).half().float()
    y = _act(F.conv2d(h, wp.permute(0, 3, 1, 2), bp), act).permute(0, 2, 3, 1)
    return y + res if res is not None else y


IRBLOCK_CASES = [
    # B, H, W, Cin, Cexp, Cout, stride, residual
    (2, 150, 150, 16, 96, 24, 2, False),      # block 1
    (2, 75, 75, 24, 144, 24, 1, True),        # block 2
    (2, 75, 75, 24, 144, 32, 2, False),       # block 3
    (2, 38, 38, 32, 192, 32, 1, True),        # blocks 4, 5
    (3, 38, 38, 32, 192, 64, 2, False),       # block 6
    (2, 19, 19, 64, 384, 64, 1, True),        # blocks 7-9
    (2, 19, 19, 96, 576, 96, 1, True),        # blocks 11, 12  (two input chunks)
    (3, 10, 10, 160, 960, 160, 1, True),      # blocks 14, 15  (three input chunks)
    (1, 7, 9, 8, 40, 16, 1, False),           # odd sizes, one slice, channels below one MMA K step
    (5, 5, 5, 72, 200, 48, 2, False),         # several images per tile
]


@pytest.mark.parametrize("mode", [1, 0])
@pytest.mark.parametrize("case", IRBLOCK_CASES + [
    (1, 33, 41, 24, 144, 24, 1, True),        # mma.sync variant: partial tiles in both directions, odd sizes
    (2, 31, 26, 16, 96, 24, 2, False), (1, 40, 17, 32, 192, 64, 2, False), (2, 21, 23, 32, 192, 32, 1, True),
    (1, 20, 34, 24, 144, 32, 2, False),
    (1, 23, 17, 64, 384, 96, 1, False),       # grouped mma.sync variant (small maps): partial tiles, block 10's shape
    (2, 7, 12, 160, 960, 160, 1, True), (1, 19, 19, 96, 576, 96, 1, False),
])
def test_irblock_fused_against_torch(case, mode):
    """ssd_irblock (whole inverted-residual block in one launch) against torch-CPU with the same fp16 roundings -- both
    implementations: the tcgen05 pipeline (mode 0) and, for the block shapes that have an instantiation, the mma.sync kernels
    (mode 1: one CTA per small tile for the large maps, channel-grouped for the small maps)."""
    from tf_ssd_b200 import _ffi
    from tf_ssd_b200._ffi_conv import IrBlockDesc
    import ctypes as C
    B, H, W, Cin, Cexp, Cout, stride, residual = case
    rng = np.random.default_rng(sum(case))
    if stride == 1:
        ph = pw = (1, 1)
    else:
        ph, pw = (1 - H % 2, 1), (1 - W % 2, 1)                          # keras_applications correct_pad
    Ho, Wo = (H + ph[0] + ph[1] - 3) // stride + 1, (W + pw[0] + pw[1] - 3) // stride + 1
    x = rng.standard_normal((B, H, W, Cin)).astype(np.float16)
    we = (rng.standard_normal((Cexp, 1, 1, Cin)) * np.sqrt(2.0 / Cin)).astype(np.float16)
    be = (0.3 * rng.standard_normal(Cexp)).astype(np.float32)
    wd = (rng.standard_normal((3, 3, Cexp)) * 0.4).astype(np.float16)
    bd = (0.3 * rng.standard_normal(Cexp)).astype(np.float32)
    wp = (rng.standard_normal((Cout, 1, 1, Cexp)) * np.sqrt(1.0 / Cexp)).astype(np.float16)
    bp = (0.3 * rng.standard_normal(Cout)).astype(np.float32)
    res = rng.standard_normal((B, Ho, Wo, Cout)).astype(np.float16) if residual else None
    t = lambda a: torch.from_numpy(a).cuda() if a is not None else None
    xt, wet, bet, wdt, bdt, wpt, bpt, rt = map(t, (x, we, be, wd, bd, wp, bp, res))
    out = torch.full((B, Ho, Wo, Cout), float("nan"), dtype=torch.float16, device="cuda")
    d = IrBlockDesc()
    d.inp, d.exp_weight, d.exp_bias = xt.data_ptr(), wet.data_ptr(), bet.data_ptr()
    d.dw_weight, d.dw_bias, d.proj_weight, d.proj_bias = wdt.data_ptr(), bdt.data_ptr(), wpt.data_ptr(), bpt.data_ptr()
    d.residual = rt.data_ptr() if rt is not None else None
    d.out = out.data_ptr()
    d.B, d.H, d.W, d.Cin, d.Cexp, d.Ho, d.Wo, d.Cout = B, H, W, Cin, Cexp, Ho, Wo, Cout
    d.stride, d.pad_top, d.pad_left, d.exp_act, d.dw_act, d.act = stride, ph[0], pw[0], 2, 2, 0
    lib = _ffi.lib()
    lib.ssd_debug_irblock_mode(mode)
    assert lib.ssd_irblock_supported(C.byref(d)) == 1
    runs = []
    for _ in range(6):          # repeated launches must agree bit for bit: the stages hand data over through mbarriers
        out.fill_(float("nan"))  # only (compute-sanitizer's racecheck does not model them), so a real race would show here
        _ffi.check(lib.ssd_irblock(C.byref(d), _ffi.stream()), "ssd_irblock")
        runs.append(out.clone())
    torch.cuda.synchronize()
    lib.ssd_debug_irblock_mode(-1)
    assert all(torch.equal(runs[0], r) for r in runs[1:])
    f = lambda a: torch.from_numpy(a).float() if a is not None else None
    y = _irblock_reference(f(x), f(we), f(be), 2, f(wd), f(bd), stride, ph, pw, 2, f(wp), f(bp), 0, f(res))
    got = out.float().cpu().numpy()
    assert np.isfinite(got).all()
    assert _rel(got, y.numpy()) < 3e-3, _rel(got, y.numpy())


@pytest.mark.parametrize("backbone,B", [("mobilenet_v2", 2), ("vgg16", 1), ("vgg16_512", 1)])
def test_every_layer_in_situ(backbone, B):
    """Each launch of the plan is checked against torch-CPU on the launch's ACTUAL device
    input buffer (read back after the forward), with the geometry the graph gave it.  This
    pins every layer's padding / stride / dilation / fusion independently of how rounding
    differences propagate through the (random-weight, error-amplifying) network."""
    from tf_ssd_b200 import synth
    m, hp = _model(backbone)
    img = synth.make_images(B, hp["img_size"], seed=9)
    m.forward_logits(img)
    torch.cuda.synchronize()
    plan = m.plan(B)
    f32 = lambda t: t.float().cpu()
    checked = 0
    def check_conv(name, mt):
        x = f32(mt["x"]).permute(0, 3, 1, 2)
        w = f32(mt["w"]).permute(0, 3, 1, 2)
        (pt, pb), (pl, pr) = mt["ph"], mt["pw"]
        y = F.conv2d(F.pad(x, (pl, pr, pt, pb)), w, f32(mt["bias"]), stride=mt["stride"], dilation=mt["dilation"])
        y = torch.relu(y) if mt["act"] == 1 else torch.clamp(y, 0, 6) if mt["act"] == 2 else y
        y = y.permute(0, 2, 3, 1)
        if mt["res"] is not None:
            y = y + f32(mt["res"])
        y = y.numpy()
        assert y.shape[1:3] == (mt["Ho"], mt["Wo"])
        if "head" in mt:
            off, cnt, A = mt["head"]
            lab = plan.logits[:, off:off + cnt].cpu().numpy().reshape(B, mt["Ho"], mt["Wo"], A * 21)
            box = plan.deltas[:, off:off + cnt].cpu().numpy().reshape(B, mt["Ho"], mt["Wo"], A * 4)
            got, tol = np.concatenate([lab, box], -1), 2e-4
        else:
            got, tol = f32(mt["out0"]).numpy(), 2e-3
        assert _rel(got, y) < tol, f"{name}: {_rel(got, y)}"

    for s in plan.steps:
        mt = s.meta
        if s.kind == "conv":
            check_conv(s.name, mt)
        elif s.kind == "chain":
            # the small-map tail + its heads as one launch (ssd_conv_chain): every member layer on its actual input
            assert len(mt["layers"]) >= 5 and sum(1 for _, m_ in mt["layers"] if "head" in m_) >= 3
            for name, sub in mt["layers"]:
                check_conv(name, sub)
        elif s.kind == "dw":
            x = f32(mt["x"]).permute(0, 3, 1, 2)
            w = f32(mt["w"]).permute(2, 0, 1).unsqueeze(1)
            (pt, pb), (pl, pr) = mt["ph"], mt["pw"]
            y = F.conv2d(F.pad(x, (pl, pr, pt, pb)), w, f32(mt["bias"]), stride=mt["stride"], groups=w.shape[0])
            y = torch.clamp(y, 0, 6).permute(0, 2, 3, 1).numpy()
            assert _rel(f32(mt["out"]).numpy(), y) < 2e-3, s.name
        elif s.kind == "dwproj":
            # fused depthwise 3x3 (+ bias + ReLU6, rounded to fp16 like the operand tile) -> 1x1 conv (+ bias, residual)
            x = f32(mt["x"]).permute(0, 3, 1, 2)
            wd = f32(mt["dw_w"]).permute(2, 0, 1).unsqueeze(1)
            (pt, pb), (pl, pr) = mt["dw_ph"], mt["dw_pw"]
            h = F.conv2d(F.pad(x, (pl, pr, pt, pb)), wd, f32(mt["dw_bias"]), stride=mt["dw_stride"], groups=wd.shape[0])
            h = torch.clamp(h, 0, 6).half().float()
            w = f32(mt["w"]).permute(0, 3, 1, 2)
            y = F.conv2d(h, w, f32(mt["bias"]))
            y = torch.relu(y) if mt["act"] == 1 else torch.clamp(y, 0, 6) if mt["act"] == 2 else y
            y = y.permute(0, 2, 3, 1)
            if mt["res"] is not None:
                y = y + f32(mt["res"])
            assert y.shape[1:3] == (mt["Ho"], mt["Wo"])
            assert _rel(f32(mt["out0"]).numpy(), y.numpy()) < 2e-3, f"{s.name}: {_rel(f32(mt['out0']).numpy(), y.numpy())}"
        elif s.kind == "irblock":
            y = _irblock_reference(f32(mt["x"]), f32(mt["exp_w"]), f32(mt["exp_bias"]), mt["exp_act"], f32(mt["dw_w"]),
                                   f32(mt["dw_bias"]), mt["dw_stride"], mt["dw_ph"], mt["dw_pw"], mt["dw_act"], f32(mt["w"]),
                                   f32(mt["bias"]), mt["act"], f32(mt["res"]) if mt["res"] is not None else None)
            assert y.shape[1:3] == (mt["Ho"], mt["Wo"])
            assert _rel(f32(mt["out0"]).numpy(), y.numpy()) < 3e-3, f"{s.name}: {_rel(f32(mt['out0']).numpy(), y.numpy())}"
        elif s.kind == "pool":
            x = f32(mt["x"]).permute(0, 3, 1, 2)
            (pt, pb), (pl, pr) = mt["ph"], mt["pw"]
            y = F.max_pool2d(F.pad(x, (pl, pr, pt, pb), value=float("-inf")), mt["k"], mt["stride"]).permute(0, 2, 3, 1)
            assert np.array_equal(f32(mt["out"]).numpy(), y.numpy()), s.name
        elif s.kind == "l2norm":
            x = f32(mt["x"]).numpy()
            y = x / np.sqrt(np.maximum((x * x).sum(-1, keepdims=True), 1e-12)) * f32(mt["scale"]).numpy()
            assert _rel(f32(mt["out"]).numpy(), y) < 2e-3, s.name
        elif s.kind == "stemblock":
            # Conv1 -> expanded_conv_depthwise -> expanded_conv_project as one launch (ssd_stem_dwproj)
            x = mt["x"].half().float().cpu().permute(0, 3, 1, 2)
            (pt, pb), (pl, pr) = mt["ph"], mt["pw"]
            h = F.conv2d(F.pad(x, (pl, pr, pt, pb)), f32(mt["stem_w"]).permute(0, 3, 1, 2), f32(mt["stem_bias"]), stride=2)
            h = _act(h, mt["stem_act"]).half().float()
            wd = f32(mt["dw_w"]).permute(2, 0, 1).unsqueeze(1)
            h = _act(F.conv2d(F.pad(h, (1, 1, 1, 1)), wd, f32(mt["dw_bias"]), groups=wd.shape[0]), mt["dw_act"]).half().float()
            y = _act(F.conv2d(h, f32(mt["w"]).permute(0, 3, 1, 2), f32(mt["bias"])), mt["act"]).permute(0, 2, 3, 1)
            assert y.shape[1:3] == (mt["Ho"], mt["Wo"])
            assert _rel(f32(mt["out0"]).numpy(), y.numpy()) < 3e-3, f"{s.name}: {_rel(f32(mt['out0']).numpy(), y.numpy())}"
        elif s.kind == "stem":
            x = mt["x"].half().float().cpu().permute(0, 3, 1, 2)          # the kernel rounds the image to fp16 first
            w = f32(mt["w"]).permute(0, 3, 1, 2)
            (pt, pb), (pl, pr) = mt["ph"], mt["pw"]
            y = F.conv2d(F.pad(x, (pl, pr, pt, pb)), w, f32(mt["bias"]), stride=mt["stride"])
            y = (torch.clamp(y, 0, 6) if mt["act"] == 2 else torch.relu(y) if mt["act"] == 1 else y).permute(0, 2, 3, 1).numpy()
            assert _rel(f32(mt["out"]).numpy(), y) < 2e-3, s.name
        else:
            assert s.kind == "cast"
            continue
        checked += 1
    assert checked == plan.n_launches - sum(1 for s in plan.steps if s.kind == "cast")
    assert any(s.kind in ("stem", "stemblock") for s in plan.steps)   # Conv1 (MobileNetV2) / conv1_1 (VGG16) read the image directly
    assert any(s.kind == "chain" for s in plan.steps)


@pytest.mark.parametrize("backbone,B", [("mobilenet_v2", 2), ("vgg16", 1)])
def test_forward_parity(backbone, B):
    """Whole-graph parity.  The random-weight networks amplify rounding differences, so the
    yardstick is the oracle's own fp16-storage error: our distance to the reference's fp32
    arithmetic must not exceed 1.5x the distance of the oracle's fp16sim mode to it."""
    from tf_ssd_b200 import synth
    m, hp = _model(backbone)
    img = synth.make_images(B, hp["img_size"], seed=9)
    d, z = m.forward_logits(img)
    torch.cuda.synchronize()
    d, z = d.cpu().numpy(), z.cpu().numpy()
    assert d.shape == (B, m.n_anchors, 4) and z.shape == (B, m.n_anchors, 21)
    (sd, sz), staps = no.forward(backbone, m.weights, hp, img, mode="fp16sim", return_logits=True, return_taps=True)
    (fd, fz), ftaps = no.forward(backbone, m.weights, hp, img, mode="fp32", return_logits=True, return_taps=True)
    taps = [t.t.float().cpu().numpy() for t in m.plan(B).taps]
    for i, (a, s_, f_) in enumerate(zip(taps, staps, ftaps)):
        assert a.shape == f_.shape, f"tap {i}"
        inherent = _rms(s_, f_)
        assert _rms(a, f_) < 1.5 * inherent + 1e-3, f"tap {i}: {_rms(a, f_)} vs inherent {inherent}"
        assert _rms(a, s_) < 1.5 * inherent + 1e-3, f"tap {i}: {_rms(a, s_)} vs inherent {inherent}"
    for a, s_, f_ in ((d, sd, fd), (z, sz, fz)):
        inherent = _rms(s_, f_)
        assert _rms(a, f_) < 1.5 * inherent + 1e-3 and _rms(a, s_) < 1.5 * inherent + 1e-3
    # public signature: model(x) -> (deltas, softmax probabilities)
    pd, pp = m(img)
    assert np.allclose(pp.cpu().numpy(), bo.softmax(z), rtol=1e-5, atol=1e-7)
    assert np.array_equal(pd.cpu().numpy(), d)


def test_decoder_model_end_to_end():
    """predictor.py:80-97 flow: get_model -> get_decoder_model -> predict; the fused graph
    (forward + softmax + decode + NMS) must equal the oracle's decoder run on the
    model's own head outputs, selection order included."""
    from tf_ssd_b200 import synth
    from tf_ssd_b200.models.decoder import get_decoder_model
    from tf_ssd_b200.utils import bbox_utils
    m, hp = _model("mobilenet_v2", seed=11)
    # make the random-weight head confident somewhere so NMS has work: boost the label biases
    w = {}
    rng = np.random.default_rng(0)
    for i in range(1, 7):
        b = m.weights[f"{i}_conv_label_output/bias"].copy().reshape(-1, 21)
        b += rng.normal(0, 2.5, b.shape).astype(np.float32)
        w[f"{i}_conv_label_output/bias"] = b.reshape(-1)
        # keep the regression outputs in a trained detector's range (|delta| of a few units): exp() of the
        # raw random-weight outputs overflows and inf - inf = NaN has no defined clip in either implementation
        w[f"{i}_conv_boxes_output/kernel"] = m.weights[f"{i}_conv_boxes_output/kernel"] * np.float32(0.02)
    m.set_weights(w)
    priors = bbox_utils.generate_prior_boxes(hp["feature_map_shapes"], hp["aspect_ratios"])
    dm = get_decoder_model(m, priors, hp)
    B = 4
    batches = [synth.make_images(B, 300, seed=s) for s in (1, 2)]
    boxes, labels, scores = dm.predict(batches, steps=2)
    assert boxes.shape == (2 * B, 200, 4) and labels.shape == (2 * B, 200) and scores.shape == (2 * B, 200)
    from tf_ssd_b200.models.decoder import SSDDecoder
    from tf_ssd_b200.models.header import softmax
    for i, img in enumerate(batches):
        d, z = m.forward_logits(img)
        sl = slice(i * B, (i + 1) * B)
        # (1) the fused graph == the stand-alone decoder layer on the model's own outputs, bit for bit
        gb, gl, gs = SSDDecoder(priors, hp["variances"])([d, softmax(z)])
        assert np.array_equal(boxes[sl], gb.cpu().numpy()) and np.array_equal(labels[sl], gl.cpu().numpy())
        assert np.array_equal(scores[sl], gs.cpu().numpy())
        # (2) against the oracle's decoder on the SAME probabilities (the device softmax, itself held to the oracle's
        # softmax in test_forward_parity): selection order, labels and scores bit for bit
        rb, rl, rs = bo.ssd_decode(priors.cpu().numpy(), hp["variances"], d.cpu().numpy(), softmax(z).cpu().numpy())
        assert (rs > 0).sum() > 0 and np.isfinite(rb).all()
        assert np.array_equal(labels[sl], rl) and np.array_equal(scores[sl], rs)
        assert np.isclose(boxes[sl], rb, rtol=1e-5, atol=2e-5).all(-1).mean() > 0.999
        # (3) against the oracle end to end (its own softmax).  exp() differs in the last bit between NumPy and the device,
        # so two detections whose scores are one ulp apart may swap places: scores must agree to 1e-6, labels everywhere
        # except at such swaps, and all but a handful of boxes.
        rb, rl, rs = bo.ssd_decode(priors.cpu().numpy(), hp["variances"], d.cpu().numpy(), bo.softmax(z.cpu().numpy()))
        assert np.allclose(scores[sl], rs, rtol=1e-6, atol=0)
        for bi, ki in np.argwhere(labels[sl] != rl):
            gaps = [abs(rs[bi, ki] - rs[bi, k2]) for k2 in (ki - 1, ki + 1) if 0 <= k2 < rs.shape[1]]
            assert min(gaps) <= 2e-6 * rs[bi, ki], (bi, ki, rs[bi, max(ki - 1, 0):ki + 2])
        assert (labels[sl] != rl).mean() < 0.02
        close = np.isclose(boxes[sl], rb, rtol=1e-5, atol=2e-5).all(-1)
        assert close.mean() > 0.98, close.mean()


def test_programmatic_dependent_launch_does_not_change_results():
    """The plan with programmatic dependent launch on every kernel -> kernel edge (the default) and with plain stream
    order (ssd_set_pdl(0): what runs under Nsight Compute / compute-sanitizer) gives bit-identical head outputs and
    detections, eagerly and through the captured graph (same block-kernel implementations in both runs: across
    implementations the outputs agree within the parity tolerance, not bit for bit)."""
    from tf_ssd_b200 import _ffi
    from tf_ssd_b200.models.decoder import get_decoder_model
    from tf_ssd_b200.utils import bbox_utils
    lib = _ffi.lib()
    u8 = np.random.default_rng(11).integers(0, 256, (3, 300, 300, 3), dtype=np.uint8)
    outs = []
    try:
        for mode in (1, 0):
            lib.ssd_set_pdl(mode)
            lib.ssd_debug_irblock_mode(1)                      # same block kernels in both runs
            m, hp = _model("mobilenet_v2")
            d, z = [t.clone() for t in m.forward_logits(u8)]
            priors = bbox_utils.generate_prior_boxes(hp["feature_map_shapes"], hp["aspect_ratios"])
            dets = get_decoder_model(m, priors, hp).predict([u8, u8], steps=2)
            torch.cuda.synchronize()
            outs.append((d, z, dets))
    finally:
        lib.ssd_set_pdl(-1)
        lib.ssd_debug_irblock_mode(-1)
    (d1, z1, p1), (d0, z0, p0) = outs
    assert torch.equal(d1, d0) and torch.equal(z1, z0)
    assert all(np.array_equal(a, b) for a, b in zip(p1, p0))
    assert np.array_equal(p1[0][:3], p1[0][3:])               # the two identical batches of the graph path agree too


@pytest.mark.parametrize("backbone", ["mobilenet_v2", "vgg16"])
def test_uint8_input_is_bit_identical_to_converted_float32(backbone):
    """A uint8 NHWC batch (the image before ``tf.image.convert_image_dtype``, utils/data_utils.py:33-37) must give
    bit-identical head outputs to the float32 batch ``float32(u8) * float32(1/255)`` the reference model receives."""
    m, hp = _model(backbone)
    B, S = 2, hp["img_size"]
    u8 = np.random.default_rng(5).integers(0, 256, (B, S, S, 3), dtype=np.uint8)
    f32 = u8.astype(np.float32) * np.float32(1.0 / 255.0)
    d0, z0 = [t.clone() for t in m.forward_logits(f32)]
    d1, z1 = [t.clone() for t in m.forward_logits(u8)]
    torch.cuda.synchronize()
    assert torch.equal(d0, d1) and torch.equal(z0, z1)
    d2, _ = m.forward_logits(torch.from_numpy(u8).cuda())            # device-resident uint8 tensor
    assert torch.equal(d0, d2)


def test_decoder_model_uint8_batches_and_parallel_heads():
    """predict() on uint8 host batches == predict() on the converted float32 batches (bit for bit), and the captured
    graph with the heads as parallel branches == the same launches in plain stream order."""
    from tf_ssd_b200.models import engine
    from tf_ssd_b200.models.decoder import get_decoder_model
    from tf_ssd_b200.utils import bbox_utils
    m, hp = _model("mobilenet_v2", seed=11)
    priors = bbox_utils.generate_prior_boxes(hp["feature_map_shapes"], hp["aspect_ratios"])
    dm = get_decoder_model(m, priors, hp)
    B = 3
    rng = np.random.default_rng(6)
    u8 = [rng.integers(0, 256, (B, 300, 300, 3), dtype=np.uint8) for _ in range(3)]
    f32 = [x.astype(np.float32) * np.float32(1.0 / 255.0) for x in u8]
    r8 = dm.predict(u8, steps=3)
    rf = dm.predict(f32, steps=3)
    for a, b in zip(r8, rf):
        assert np.array_equal(a, b)
    plan = m.plan(B)
    # heads 1-3 are side branches hoisted behind their taps; heads 4-6 live inside the fused tail (ssd_conv_chain)
    assert sum(1 for s in plan.steps if s.branch) == 3 and plan.steps[-1].kind == "chain"
    first_head = next(i for i, s in enumerate(plan.steps) if s.branch)
    assert first_head < len(plan.steps) - 8
    m._to_image_buffer(plan, u8[0])
    plan.run(u8=True, parallel=True)
    torch.cuda.synchronize()
    d_par, z_par = plan.deltas.clone(), plan.logits.clone()
    plan.deltas.zero_(); plan.logits.zero_()
    plan.run(u8=True, parallel=False)
    torch.cuda.synchronize()
    assert torch.equal(d_par, plan.deltas) and torch.equal(z_par, plan.logits)
    # eager (no CUDA graph) decoder model agrees with the captured one
    dm2 = engine.DecoderModel(m, dm.decoder, use_cuda_graph=False)
    for a, b in zip(dm2.predict(u8[:1], steps=1), [r[:B] for r in r8]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("H,W,pad", [(300, 300, (0, 0)), (31, 45, (1, 1)), (64, 322, (0, 1)), (9, 7, (1, 0))])
def test_stem_kernel_shapes(H, W, pad):
    """ssd_stem_conv3x3s2 / _u8 on odd sizes, both paddings, rows longer than one CTA chunk (Wo > 160)."""
    from tf_ssd_b200 import _ffi
    lib = _ffi.lib()
    rng = np.random.default_rng(H * W)
    B = 2
    pt, pl = pad
    Ho, Wo = (H + pt - 2) // 2 + 1, (W + pl - 2) // 2 + 1          # at most one padded row / column after
    u8 = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    x32 = u8.astype(np.float32) * np.float32(1.0 / 255.0)
    w = (rng.standard_normal((32, 3, 3, 3)) * 0.3).astype(np.float16)
    bias = rng.standard_normal(32).astype(np.float32)
    xt, ut = torch.from_numpy(x32).cuda(), torch.from_numpy(u8).cuda()
    wt, bt = torch.from_numpy(w).cuda(), torch.from_numpy(bias).cuda()
    outs = []
    for fn, src in ((lib.ssd_stem_conv3x3s2, xt), (lib.ssd_stem_conv3x3s2_u8, ut)):
        out = torch.full((B, Ho, Wo, 32), float("nan"), dtype=torch.float16, device="cuda")
        _ffi.check(fn(_ffi.ptr(src), _ffi.ptr(wt), _ffi.ptr(bt), _ffi.ptr(out), B, H, W, 32, Ho, Wo, pt, pl, 2, _ffi.stream()))
        outs.append(out)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    xr = torch.from_numpy(x32).half().float().permute(0, 3, 1, 2)
    pb, pr = max(0, (Ho - 1) * 2 + 3 - H - pt), max(0, (Wo - 1) * 2 + 3 - W - pl)
    y = F.conv2d(F.pad(xr, (pl, pr, pt, pb)), torch.from_numpy(w).float().permute(0, 3, 1, 2), torch.from_numpy(bias), stride=2)
    y = torch.clamp(y, 0, 6).permute(0, 2, 3, 1).numpy()
    assert y.shape == (B, Ho, Wo, 32)
    assert _rel(outs[0].float().cpu().numpy(), y) < 2e-3


@pytest.mark.parametrize("case", [
    # B, H, W, pad (top, left), Cout, u8
    (2, 300, 300, (0, 0), 16, True), (2, 300, 300, (0, 0), 16, False), (1, 67, 45, (1, 1), 8, False),
    (3, 64, 122, (0, 1), 32, True), (2, 9, 7, (1, 0), 24, True), (1, 130, 25, (0, 0), 16, False),
])
def test_stem_dwproj_fused_against_torch(case):
    """ssd_stem_dwproj (Conv1 3x3 s2 + ReLU6 -> depthwise 3x3 + ReLU6 -> 1x1 projection, one launch) against torch-CPU
    with the same fp16 roundings: SSD300's shape, partial tiles in both directions, both paddings, row lengths that are
    not multiples of four elements (scalar staging path), all output widths, float32 and uint8 images (bit-identical)."""
    import ctypes as C
    from tf_ssd_b200 import _ffi
    from tf_ssd_b200._ffi_conv import StemDwProjDesc
    B, H, W, (pt, pl), Cout, u8_in = case
    lib = _ffi.lib()
    rng = np.random.default_rng(H * W + Cout)
    Hs, Ws = (H + pt - 2) // 2 + 1, (W + pl - 2) // 2 + 1
    u8 = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    x32 = u8.astype(np.float32) * np.float32(1.0 / 255.0)
    ws = (rng.standard_normal((32, 3, 3, 3)) * 0.5).astype(np.float16)
    bs = rng.standard_normal(32).astype(np.float32)
    wd = (rng.standard_normal((3, 3, 32)) * 0.4).astype(np.float16)
    bd = (0.3 * rng.standard_normal(32)).astype(np.float32)
    wp = (rng.standard_normal((Cout, 32)) * 0.25).astype(np.float16)
    bp = (0.3 * rng.standard_normal(Cout)).astype(np.float32)
    t = lambda a: torch.from_numpy(a).cuda()
    xt, ut, wst, bst, wdt, bdt, wpt, bpt = map(t, (x32, u8, ws, bs, wd, bd, wp, bp))
    outs = []
    for is_u8 in (False, True):
        out = torch.full((B, Hs, Ws, Cout), float("nan"), dtype=torch.float16, device="cuda")
        d = StemDwProjDesc()
        d.image, d.image_u8 = (ut if is_u8 else xt).data_ptr(), int(is_u8)
        d.stem_weight, d.stem_bias, d.dw_weight, d.dw_bias = wst.data_ptr(), bst.data_ptr(), wdt.data_ptr(), bdt.data_ptr()
        d.proj_weight, d.proj_bias, d.out = wpt.data_ptr(), bpt.data_ptr(), out.data_ptr()
        d.B, d.H, d.W, d.Hs, d.Ws, d.Cmid, d.Cout = B, H, W, Hs, Ws, 32, Cout
        d.pad_top, d.pad_left, d.stem_act, d.dw_act, d.act = pt, pl, 2, 2, 0
        assert lib.ssd_stem_dwproj_supported(C.byref(d)) == 1
        _ffi.check(lib.ssd_stem_dwproj(C.byref(d), _ffi.stream()), "ssd_stem_dwproj")
        outs.append(out)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])                       # convert_image_dtype fused into the load: bit-identical
    xr = torch.from_numpy(x32).half().float().permute(0, 3, 1, 2)
    pb, pr = max(0, (Hs - 1) * 2 + 3 - H - pt), max(0, (Ws - 1) * 2 + 3 - W - pl)
    h = F.conv2d(F.pad(xr, (pl, pr, pt, pb)), torch.from_numpy(ws).float().permute(0, 3, 1, 2), torch.from_numpy(bs), stride=2)
    h = torch.clamp(h, 0, 6).half().float()
    h = F.conv2d(F.pad(h, (1, 1, 1, 1)), torch.from_numpy(wd).float().permute(2, 0, 1).unsqueeze(1), torch.from_numpy(bd), groups=32)
    h = torch.clamp(h, 0, 6).half().float()
    y = F.conv2d(h, torch.from_numpy(wp).float().view(Cout, 32, 1, 1), torch.from_numpy(bp)).permute(0, 2, 3, 1).numpy()
    got = outs[1 if u8_in else 0].float().cpu().numpy()
    assert y.shape == got.shape and np.isfinite(got).all()
    assert _rel(got, y) < 3e-3, _rel(got, y)
    d.Cmid = 64                                                 # another stem width: refused, not mis-computed
    assert lib.ssd_stem_dwproj_supported(C.byref(d)) == 0 and lib.ssd_stem_dwproj(C.byref(d), _ffi.stream()) < 0


@pytest.mark.parametrize("H,W", [(300, 300), (33, 47), (5, 200), (1, 1)])
def test_stem_kernel_stride1_vgg_first_layer(H, W):
    """ssd_stem_conv3x3 / _u8 with stride 1, Cout 64, SAME padding (VGG16 conv1_1, models/ssd_vgg16.py:80)."""
    from tf_ssd_b200 import _ffi
    lib = _ffi.lib()
    rng = np.random.default_rng(H * 1000 + W)
    B = 2
    u8 = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    x32 = u8.astype(np.float32) * np.float32(1.0 / 255.0)
    w = (rng.standard_normal((64, 3, 3, 3)) * 0.3).astype(np.float16)
    bias = rng.standard_normal(64).astype(np.float32)
    xt, ut = torch.from_numpy(x32).cuda(), torch.from_numpy(u8).cuda()
    wt, bt = torch.from_numpy(w).cuda(), torch.from_numpy(bias).cuda()
    outs = []
    for fn, src in ((lib.ssd_stem_conv3x3, xt), (lib.ssd_stem_conv3x3_u8, ut)):
        out = torch.full((B, H, W, 64), float("nan"), dtype=torch.float16, device="cuda")
        _ffi.check(fn(_ffi.ptr(src), _ffi.ptr(wt), _ffi.ptr(bt), _ffi.ptr(out), B, H, W, 64, H, W, 1, 1, 1, 1, _ffi.stream()))
        outs.append(out)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    xr = torch.from_numpy(x32).half().float().permute(0, 3, 1, 2)
    y = F.conv2d(xr, torch.from_numpy(w).float().permute(0, 3, 1, 2), torch.from_numpy(bias), padding=1)
    y = torch.relu(y).permute(0, 2, 3, 1).numpy()
    assert _rel(outs[0].float().cpu().numpy(), y) < 2e-3
    # unsupported combinations are refused
    assert lib.ssd_stem_conv3x3(_ffi.ptr(xt), _ffi.ptr(wt), _ffi.ptr(bt), _ffi.ptr(outs[0]), B, H, W, 48, H, W, 1, 1, 1, 1, _ffi.stream()) == -4


@pytest.mark.parametrize("B", [1, 3, 5])
def test_conv_chain_against_torch(B):
    """ssd_conv_chain on a VGG-like tail with ragged channel counts: 1x1 -> 3x3 stride 2 SAME -> 1x1 -> 3x3 VALID, plus a
    split fp32 "head" (3x3 SAME, 150 = 6*21 + 6*4 channels) on the first 3x3's map in the same phase as the next 1x1.
    B = 3 and 5 leave a partially filled cluster (images per cluster does not divide the batch)."""
    from tf_ssd_b200 import _ffi
    from tf_ssd_b200._ffi_conv import ConvDesc
    rng = np.random.default_rng(40 + B)
    dev = torch.device("cuda")
    lib = _ffi.lib()

    def layer(x_t, H, W, cin, cout, k, stride, pads, act, head=None):
        (pt, pb), (pl, pr) = pads
        Ho, Wo = (H + pt + pb - k) // stride + 1, (W + pl + pr - k) // stride + 1
        w = (rng.standard_normal((cout, k, k, cin)) * np.sqrt(2.0 / (k * k * cin))).astype(np.float16)
        b = (0.1 * rng.standard_normal(cout)).astype(np.float32)
        wd, bd = torch.from_numpy(w).to(dev), torch.from_numpy(b).to(dev)
        d = ConvDesc()
        d.inp, d.weight, d.bias, d.residual = x_t.data_ptr(), wd.data_ptr(), bd.data_ptr(), None
        d.B, d.H, d.W, d.Cin, d.Ho, d.Wo, d.Cout = B, H, W, cin, Ho, Wo, cout
        d.KH = d.KW = k
        d.stride, d.dilation, d.pad_top, d.pad_left, d.act = stride, 1, pt, pl, act
        if head is None:
            out = torch.zeros((B, Ho, Wo, cout), dtype=torch.float16, device=dev)
            d.out0, d.out1, d.out_f32, d.split = out.data_ptr(), None, 0, cout
            d.img_stride0, d.pix_stride0, d.img_stride1, d.pix_stride1 = Ho * Wo * cout, cout, 0, 0
            outs = (out,)
        else:
            sp = head
            o0 = torch.zeros((B, Ho, Wo, sp), dtype=torch.float32, device=dev)
            o1 = torch.zeros((B, Ho, Wo, cout - sp), dtype=torch.float32, device=dev)
            d.out0, d.out1, d.out_f32, d.split = o0.data_ptr(), o1.data_ptr(), 1, sp
            d.img_stride0, d.pix_stride0, d.img_stride1, d.pix_stride1 = Ho * Wo * sp, sp, Ho * Wo * (cout - sp), cout - sp
            outs = (o0, o1)
        return dict(d=d, w=w, b=b, pads=pads, k=k, stride=stride, act=act, x=x_t, outs=outs, keep=(wd, bd), Ho=Ho, Wo=Wo)

    x0 = torch.from_numpy(rng.standard_normal((B, 5, 5, 256)).astype(np.float16)).to(dev)
    l0 = layer(x0, 5, 5, 256, 128, 1, 1, ((0, 0), (0, 0)), 1)
    l1 = layer(l0["outs"][0], 5, 5, 128, 256, 3, 2, ((1, 1), (1, 1)), 1)                   # -> 3x3x256
    l2 = layer(l1["outs"][0], 3, 3, 256, 128, 1, 1, ((0, 0), (0, 0)), 1)
    lh = layer(l1["outs"][0], 3, 3, 256, 150, 3, 1, ((1, 1), (1, 1)), 0, head=126)          # head on l1's map, phase of l2
    l3 = layer(l2["outs"][0], 3, 3, 128, 40, 3, 1, ((0, 0), (0, 0)), 2)                    # VALID -> 1x1x40 (ragged tiles)
    layers, phases = [l0, l1, l2, lh, l3], [0, 1, 2, 2, 3]
    descs = (ConvDesc * len(layers))(*[l["d"] for l in layers])
    ph = (C.c_int32 * len(layers))(*phases)
    assert lib.ssd_conv_chain_supported(descs, ph, len(layers)) == 1
    _ffi.check(lib.ssd_conv_chain(descs, ph, len(layers), _ffi.stream()), "ssd_conv_chain")
    torch.cuda.synchronize()
    for i, l in enumerate(layers):
        x = l["x"].float().cpu().permute(0, 3, 1, 2)
        (pt, pb), (pl, pr) = l["pads"]
        y = F.conv2d(F.pad(x, (pl, pr, pt, pb)), torch.from_numpy(l["w"].astype(np.float32)).permute(0, 3, 1, 2),
                     torch.from_numpy(l["b"]), stride=l["stride"])
        y = torch.relu(y) if l["act"] == 1 else torch.clamp(y, 0, 6) if l["act"] == 2 else y
        y = y.permute(0, 2, 3, 1).numpy()
        got = np.concatenate([o.float().cpu().numpy() for o in l["outs"]], -1)
        assert got.shape == y.shape
        assert _rel(got, y) < (2e-4 if len(l["outs"]) == 2 else 2e-3), (i, _rel(got, y))
    # unsupported chains are refused, not mis-computed
    bad = (C.c_int32 * len(layers))(0, 1, 2, 1, 3)
    assert lib.ssd_conv_chain_supported(descs, bad, len(layers)) == 0
    assert lib.ssd_conv_chain(descs, bad, len(layers), _ffi.stream()) == -4
