"""Forward engine behind ``get_model`` (models/ssd_mobilenet_v2.py:15-47,
models/ssd_vgg16.py:66-121, models/header.py:54-90 of the reference).

The reference builds Keras graphs; here each graph is written once against a
small "net" interface and interpreted twice:

* ``_ParamTracer``  walks it to enumerate the trainable variables (Keras layer
  names, Keras layouts) and to initialise them;
* ``_PlanBuilder``  walks it for a concrete batch size and emits a flat list
  of C-ABI kernel launches (``ssd_conv2d``, ``ssd_depthwise3x3``, ...) over
  preallocated NHWC fp16 device buffers.  A plan allocates nothing and never
  synchronises when it runs, so it is captured once into a CUDA graph and
  replayed.

torch is used for device buffers, streams and graph capture only -- no
``torch.nn`` and no torch arithmetic on the activation path.
"""

from __future__ import annotations

import contextlib
import ctypes as C
import gc
import math
import os
from dataclasses import dataclass
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from tf_ssd_b200 import _ffi
from tf_ssd_b200._ffi_conv import ACT_NONE, ACT_RELU, ACT_RELU6, ConvDesc, DwProjDesc, IrBlockDesc, StemDwProjDesc

BN_EPS = 1e-3        # keras_applications.mobilenet_v2: BatchNormalization(epsilon=1e-3, momentum=0.999)

# (expansion, out channels, stride) of inverted-residual blocks 0..16, alpha = 1.0
MNV2_BLOCKS: List[Tuple[int, int, int]] = [
    (1, 16, 1),
    (6, 24, 2), (6, 24, 1),
    (6, 32, 2), (6, 32, 1), (6, 32, 1),
    (6, 64, 2), (6, 64, 1), (6, 64, 1), (6, 64, 1),
    (6, 96, 1), (6, 96, 1), (6, 96, 1),
    (6, 160, 2), (6, 160, 1), (6, 160, 1),
    (6, 320, 1),
]


# ------------------------------------------------------------------ geometry --
def same_pad(size: int, k: int, s: int, d: int = 1) -> Tuple[int, int]:
    """TensorFlow ``padding="same"``: total padding split with the odd pixel AFTER."""
    out = -(-size // s)
    total = max((out - 1) * s + (k - 1) * d + 1 - size, 0)
    return total // 2, total - total // 2


def correct_pad(size: int, k: int = 3) -> Tuple[int, int]:
    """keras_applications ``correct_pad`` (ZeroPadding2D before the stride-2 VALID convs)."""
    return k // 2 - (1 - size % 2), k // 2


def _out_size(size: int, k: int, s: int, d: int, pads: Tuple[int, int]) -> int:
    return (size + pads[0] + pads[1] - ((k - 1) * d + 1)) // s + 1


def _resolve_pads(H: int, W: int, k: int, s: int, d: int, pad: str):
    if pad == "same":
        return same_pad(H, k, s, d), same_pad(W, k, s, d)
    if pad == "correct":
        return correct_pad(H, k), correct_pad(W, k)
    if pad == "valid":
        return (0, 0), (0, 0)
    raise ValueError(f"unknown padding {pad!r}")


# ---------------------------------------------------------------- the graphs --
@dataclass
class Act:
    """An NHWC activation: device buffer (None while tracing) + logical shape."""
    t: Optional[torch.Tensor]
    H: int
    W: int
    C: int
    is_image: bool = False          # the fp16 copy of the input image (3 real channels padded to 8)


def mobilenet_v2_graph(n: Any, x: Act, hp: Dict[str, Any]) -> List[Act]:
    """models/ssd_mobilenet_v2.py:24-46 over keras_applications MobileNetV2
    (alpha 1.0, include_top=False): returns the six head taps."""
    x = n.conv(x, "Conv1", 32, k=3, stride=2, pad="correct", act=ACT_RELU6, bn="bn_Conv1", use_bias=False)
    taps: List[Act] = []
    for bid, (t, c, s) in enumerate(MNV2_BLOCKS):
        prefix = f"block_{bid}_" if bid else "expanded_conv_"
        inp = x
        if bid:
            # block_13_expand_relu is a head tap (:27): it must exist in memory, so that block is not fused whole
            x = n.conv(x, prefix + "expand", t * inp.C, act=ACT_RELU6, bn=prefix + "expand_BN", use_bias=False,
                       tap=(bid == 13))
            if bid == 13:
                taps.append(x)
        x = n.dw(x, prefix + "depthwise", stride=s, act=ACT_RELU6, bn=prefix + "depthwise_BN")
        res = inp if (s == 1 and inp.C == c) else None          # block_i_add
        x = n.conv(x, prefix + "project", c, act=ACT_NONE, bn=prefix + "project_BN", use_bias=False, residual=res)
    x = n.conv(x, "Conv_1", 1280, act=ACT_RELU6, bn="Conv_1_bn", use_bias=False)
    taps.append(x)                                              # out_relu   :28
    for i, (c1, c2) in enumerate([(256, 512), (128, 256), (128, 256), (128, 256)], start=1):   # :31-41
        x = n.conv(x, f"extra{i}_1", c1, pad="valid", act=ACT_RELU)
        x = n.conv(x, f"extra{i}_2", c2, k=3, stride=2, pad="same", act=ACT_RELU)
        taps.append(x)
    return taps


def vgg16_graph(n: Any, x: Act, hp: Dict[str, Any]) -> List[Act]:
    """models/ssd_vgg16.py:78-119.  With seven feature maps the SSD512 extension
    (not in the reference) is built instead."""
    ssd512 = len(hp["feature_map_shapes"]) == 7
    kw = dict(k=3, act=ACT_RELU, init="glorot_normal", l2=True)
    conv4_3 = None
    for bname, reps, cout in [("conv1", 2, 64), ("conv2", 2, 128), ("conv3", 3, 256), ("conv4", 3, 512),
                              ("conv5", 3, 512)]:
        for r in range(1, reps + 1):
            x = n.conv(x, f"{bname}_{r}", cout, **kw)
        if bname == "conv4":
            conv4_3 = x
        x = n.maxpool(x, 2, 2) if bname != "conv5" else n.maxpool(x, 3, 1)      # :82-101
    x = n.conv(x, "conv6", 1024, dilation=6, **kw)                              # :103
    kw1 = dict(kw, k=1)
    conv7 = n.conv(x, "conv7", 1024, **kw1)                                     # :104
    x = n.conv(conv7, "conv8_1", 256, pad="valid", **kw1)
    conv8_2 = n.conv(x, "conv8_2", 512, stride=2, **kw)
    x = n.conv(conv8_2, "conv9_1", 128, pad="valid", **kw1)
    conv9_2 = n.conv(x, "conv9_2", 256, stride=2, **kw)
    x = n.conv(conv9_2, "conv10_1", 128, pad="valid", **kw1)
    if ssd512:
        conv10_2 = n.conv(x, "conv10_2", 256, stride=2, **kw)
        x = n.conv(conv10_2, "conv11_1", 128, pad="valid", **kw1)
        conv11_2 = n.conv(x, "conv11_2", 256, stride=2, **kw)
        x = n.conv(conv11_2, "conv12_1", 128, pad="valid", **kw1)
        conv12_2 = n.conv(x, "conv12_2", 256, stride=2, **kw)
    else:
        conv10_2 = n.conv(x, "conv10_2", 256, pad="valid", **kw)                # :111
        x = n.conv(conv10_2, "conv11_1", 128, pad="valid", **kw1)
        conv11_2 = n.conv(x, "conv11_2", 256, pad="valid", **kw)                # :113
    norm = n.l2norm(conv4_3, "l2_normalization")                                # :116
    return [norm, conv7, conv8_2, conv9_2, conv10_2, conv11_2] + ([conv12_2] if ssd512 else [])


GRAPHS: Dict[str, Callable[[Any, Act, Dict[str, Any]], List[Act]]] = {
    "mobilenet_v2": mobilenet_v2_graph,
    "vgg16": vgg16_graph,
}


# ------------------------------------------------------- parameter enumeration --
class _ParamTracer:
    """Walks a graph with shapes only and creates the Keras-named variables."""

    def __init__(self, rng: np.random.Generator):
        self.rng = rng
        self.weights: Dict[str, np.ndarray] = {}
        self.l2_kernels: List[str] = []          # kernels carrying l2(5e-4) (ssd_vgg16.py:76)
        self.macs = 0                            # multiply-accumulates per image

    def _kernel(self, shape, fan_in, fan_out, init):
        if init == "glorot_normal":              # Keras: truncated normal, std = sqrt(2/(fan_in+fan_out))/.8796
            std = math.sqrt(2.0 / (fan_in + fan_out))
            return np.clip(self.rng.standard_normal(shape), -2, 2).astype(np.float32) * np.float32(std / 0.87962566)
        if init == "he_normal":
            return (self.rng.standard_normal(shape) * math.sqrt(2.0 / fan_in)).astype(np.float32)
        lim = math.sqrt(6.0 / (fan_in + fan_out))            # glorot_uniform, the Keras default
        return self.rng.uniform(-lim, lim, shape).astype(np.float32)

    def _bn(self, name, c):
        self.weights[name + "/gamma"] = np.ones(c, np.float32)
        self.weights[name + "/beta"] = np.zeros(c, np.float32)
        self.weights[name + "/moving_mean"] = np.zeros(c, np.float32)
        self.weights[name + "/moving_variance"] = np.ones(c, np.float32)

    def conv(self, x, name, cout, k=1, stride=1, pad="same", dilation=1, act=ACT_NONE, bn=None, use_bias=True,
             residual=None, init="he_normal", l2=False, tap=False):
        ph, pw = _resolve_pads(x.H, x.W, k, stride, dilation, pad)
        Ho, Wo = _out_size(x.H, k, stride, dilation, ph), _out_size(x.W, k, stride, dilation, pw)
        self.weights[name + "/kernel"] = self._kernel((k, k, x.C, cout), k * k * x.C, k * k * cout, init)
        if use_bias:
            self.weights[name + "/bias"] = np.zeros(cout, np.float32)
        if bn:
            self._bn(bn, cout)
        if l2:
            self.l2_kernels.append(name + "/kernel")
        self.macs += Ho * Wo * k * k * x.C * cout
        return Act(None, Ho, Wo, cout)

    def dw(self, x, name, stride=1, act=ACT_RELU6, bn=None):
        ph, pw = _resolve_pads(x.H, x.W, 3, stride, 1, "same" if stride == 1 else "correct")
        Ho, Wo = _out_size(x.H, 3, stride, 1, ph), _out_size(x.W, 3, stride, 1, pw)
        self.weights[name + "/depthwise_kernel"] = self._kernel((3, 3, x.C, 1), 9, 9, "he_normal")
        if bn:
            self._bn(bn, x.C)
        self.macs += Ho * Wo * 9 * x.C
        return Act(None, Ho, Wo, x.C)

    def maxpool(self, x, k, s):
        ph, pw = same_pad(x.H, k, s), same_pad(x.W, k, s)
        return Act(None, _out_size(x.H, k, s, 1, ph), _out_size(x.W, k, s, 1, pw), x.C)

    def l2norm(self, x, name, scale_factor=20.0):
        self.weights[name + "/scale"] = np.full(x.C, scale_factor, np.float32)      # ssd_vgg16.py:38-44
        return x

    def head(self, taps, hp):
        L = int(hp["total_labels"])
        for i, t in enumerate(taps):
            A = len(hp["aspect_ratios"][i]) + 1
            for nm, c in ((f"{i + 1}_conv_label_output", A * L), (f"{i + 1}_conv_boxes_output", A * 4)):
                self.weights[nm + "/kernel"] = self._kernel((3, 3, t.C, c), 9 * t.C, 9 * c, "glorot_uniform")
                self.weights[nm + "/bias"] = np.zeros(c, np.float32)
                self.macs += t.H * t.W * 9 * t.C * c


@contextlib.contextmanager
def _no_gc_during_capture():
    """Stream capture (global mode) is invalidated by any "unsafe" CUDA call of the process -- including the
    cudaGraphExecDestroy / cudaFree that Python's cyclic garbage collector issues when it happens to reclaim an older
    DecoderModel / trainer (their graphs and buffers sit in reference cycles with the enqueue closures).  Collect now,
    and keep the collector off until the capture has ended."""
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was_enabled:
            gc.enable()


# ------------------------------------------------------------------ the plan --
@dataclass
class Step:
    name: str
    kind: str               # conv | dw | pool | l2norm | cast | softmax
    fn: Any
    args: tuple             # everything but the trailing stream argument
    flops: float            # 2 * MACs for the whole batch
    bytes: float            # algorithmic bytes: inputs + weights + outputs, once each
    keep: tuple = ()        # objects that must outlive the plan (descs, tensors)
    meta: Optional[Dict[str, Any]] = None   # tensors + geometry of the launch (tests check every layer in situ)
    branch: int = 0         # 0: the main chain; k > 0: side branch k (a multibox head: consumes its tap only)


class Plan:
    """A flat, allocation-free launch list for one (backbone, batch) pair."""

    def __init__(self, B: int, img_size: int, device: torch.device):
        self.B, self.img_size, self.device = B, img_size, device
        self.steps: List[Step] = []
        self.image = torch.zeros((B, img_size, img_size, 3), dtype=torch.float32, device=device)
        # uint8 NHWC input (what the reference's pipeline holds before convert_image_dtype, utils/data_utils.py:33-37):
        # ``first_u8`` replaces ``steps[0]`` (the only step that reads the image) when a uint8 batch is fed
        self.image_u8 = torch.zeros((B, img_size, img_size, 3), dtype=torch.uint8, device=device)
        self.first_u8: Optional[Step] = None
        self.logits: Optional[torch.Tensor] = None
        self.deltas: Optional[torch.Tensor] = None
        self.taps: List[Act] = []

    # Multibox heads consume their tap only (models/header.py:68-85), so each head is a side branch of the launch
    # graph: it is forked (event wait on a side stream) right after the step that produces its tap and joined at the
    # end.  Inside a CUDA-graph capture this yields parallel graph branches; eagerly it is plain multi-stream overlap.
    PARALLEL_HEADS = os.environ.get("SSD_B200_PARALLEL_HEADS", "1") not in ("0", "")
    _DEBUG_SKIP = frozenset(filter(None, os.environ.get("SSD_B200_DEBUG_SKIP_STEPS", "").split(",")))   # timing experiments only

    def run(self, first: int = 0, last: Optional[int] = None, u8: bool = False, parallel: Optional[bool] = None) -> None:
        parallel = self.PARALLEL_HEADS if parallel is None else parallel
        main = torch.cuda.current_stream()
        st = vp_main = _ffi.stream()
        used: Dict[int, torch.cuda.Stream] = {}
        for i, s in enumerate(self.steps[first:last], start=first):
            if s.name in self._DEBUG_SKIP:
                continue
            if u8 and i == 0:
                s = self.first_u8
            if parallel and s.branch:
                side = used.get(s.branch)
                if side is None:
                    side = self._side_stream(s.branch)
                    side.wait_stream(main)                      # fork: everything enqueued on main so far
                    used[s.branch] = side
                st = C.c_void_p(side.cuda_stream)
            else:
                st = vp_main
            rc = s.fn(*s.args, st)
            if rc != 0:
                _ffi.check(rc, f"{s.fn.__name__} [{s.name}]")
        for side in used.values():
            main.wait_stream(side)                              # join

    def _side_stream(self, k: int) -> torch.cuda.Stream:
        pool = getattr(self, "_sides", None)
        if pool is None:
            pool = self._sides = {}
        if k not in pool:
            pool[k] = torch.cuda.Stream(device=self.device)
        return pool[k]

    def hoist_heads(self) -> None:
        """Moves every head step directly behind the step that produces its tap (and marks it as a side branch) so that
        the fork happens as early as the data dependency allows.  A tap-only pre-processing step (VGG16's
        L2Normalization of conv4_3) travels with its head."""
        heads = [s for s in self.steps if s.meta and "head" in s.meta]
        rest = [s for s in self.steps if not (s.meta and "head" in s.meta)]

        def produces(step, t):
            m = step.meta or {}
            return any(m.get(k) is t for k in ("out", "out0"))

        def producer(t):
            return max((i for i, s in enumerate(rest) if produces(s, t)), default=len(rest) - 1)

        for bi, h in enumerate(heads, start=1):
            h.branch = bi
            idx = producer(h.meta["x"])
            prod = rest[idx]
            if prod.kind == "l2norm":                           # consumed by this head only: runs on the head's branch,
                prod.branch = bi                                # right behind the producer of ITS input
                rest.pop(idx)
                idx = producer(prod.meta["x"]) + 1
                rest.insert(idx, prod)
            rest.insert(idx + 1, h)
        self.steps = rest

    # the small-map tail (extras behind the 5x5 / 10x10 map + the heads of the maps it produces) as ONE launch: "1" / "0"
    FUSE_TAIL = os.environ.get("SSD_B200_FUSE_TAIL", "1")
    FUSE_TAIL_MAX_GFLOP = float(os.environ.get("SSD_B200_FUSE_TAIL_MAX_GFLOP", "0.8"))

    def fuse_tail(self, lib) -> None:
        """Replaces the longest suffix of the main chain that ``ssd_conv_chain`` can run -- plain convolutions on small
        maps, each fed by its predecessor (models/ssd_mobilenet_v2.py:33-41 extra2_1..extra4_2, models/ssd_vgg16.py:108-113
        conv9_1..conv11_2) -- and the multibox heads of the feature maps that suffix produces (models/header.py:68-85) by
        one launch.  Layers are grouped in phases by dependency depth: a head runs in the same phase as the extra layer
        that reads the same map."""
        if self.FUSE_TAIL in ("0", ""):
            return
        main = [s for s in self.steps if not s.branch]

        def chain_for(start: int):
            suffix = main[start:]
            if len(suffix) < 2 or any(s.kind != "conv" for s in suffix):
                return None
            produced = {id(s.meta["out0"]): s for s in suffix if isinstance(s.meta.get("out0"), torch.Tensor)}
            for a, b in zip(suffix[:-1], suffix[1:]):
                if b.meta["x"] is not a.meta["out0"]:
                    return None
            heads = [s for s in self.steps if s.branch and s.kind == "conv" and "head" in (s.meta or {})
                     and id(s.meta["x"]) in produced]
            if any(s.branch and s not in heads and id((s.meta or {}).get("x")) in produced for s in self.steps):
                return None                                         # some other side step consumes a map of the suffix
            phase = {}
            for s in suffix + heads:
                prod = produced.get(id(s.meta["x"]))
                phase[id(s)] = phase[id(prod)] + 1 if prod is not None else 0
            layers = sorted(suffix + heads, key=lambda s: phase[id(s)])     # stable: extras before the heads of a phase
            return layers, [phase[id(s)] for s in layers]

        for start in range(len(main)):
            got = chain_for(start)
            if got is None:
                continue
            layers, phases = got
            descs = (ConvDesc * len(layers))(*[s.args[0]._obj for s in layers])
            ph = (C.c_int32 * len(layers))(*phases)
            if lib.ssd_conv_chain_waves(descs, ph, len(layers)) != 1:      # unsupported, or the clusters need a second wave
                continue
            if sum(s.flops for s in layers) > self.FUSE_TAIL_MAX_GFLOP * 1e9:
                continue        # the cluster kernel is fragment-load bound (~20 TFLOP/s): beyond this, separate tcgen05 launches win
            members = {id(s) for s in layers}
            first = min(i for i, s in enumerate(self.steps) if id(s) in members)
            keep = tuple(k for s in layers for k in s.keep) + (descs, ph)
            step = Step("tail_chain", "chain", lib.ssd_conv_chain, (descs, ph, len(layers)), sum(s.flops for s in layers),
                        sum(s.bytes for s in layers), keep, dict(layers=[(s.name, s.meta) for s in layers], phases=phases))
            self.steps = [s for i, s in enumerate(self.steps[:first]) if id(s) not in members] + [step] + \
                         [s for s in self.steps[first:] if id(s) not in members]
            return

    @property
    def n_launches(self) -> int:
        return len(self.steps)


class _PlanBuilder:
    # depthwise + 1x1 projection as one launch (ssd_dwproj): "all", "s2" (stride-2 depthwise layers only: where the
    # fused kernel beats the two launches today) or "0"
    FUSE_DW = os.environ.get("SSD_B200_FUSE_DW", "all")
    FUSE_MIN_TILES = int(os.environ.get("SSD_B200_FUSE_MIN_TILES", "0"))     # fuse only layers with at least this many 128-pixel tiles

    def __init__(self, model: "SSDModel", B: int):
        self.m = model
        self.B = B
        self.dev = _ffi.require_cuda()
        self.lib = _ffi.lib()
        self.plan = Plan(B, model.img_size, self.dev)
        self.no_stem = False            # training plans keep the first layer on the generic path (its filter gradient reads the fp16 image)

    def _buf(self, H, W, C) -> torch.Tensor:
        return torch.empty((self.B, H, W, C), dtype=torch.float16, device=self.dev)

    def input(self) -> Act:
        """The fp32 NHWC image buffer itself (t is None): the first layer decides how to
        consume it -- the MobileNetV2 stem reads it directly, anything else gets an fp16 copy."""
        S = self.m.img_size
        return Act(None, S, S, 3)

    def _image_as_f16(self) -> Act:
        S = self.m.img_size
        x = self._buf(S, S, 8)
        npx = self.B * S * S
        self.plan.steps.append(Step("input_cast", "cast", self.lib.ssd_image_to_f16c8,
                                    (_ffi.ptr(self.plan.image), _ffi.ptr(x), npx), 0.0, npx * (12 + 16), (x,)))
        self.plan.first_u8 = Step("input_cast_u8", "cast", self.lib.ssd_image_u8_to_f16c8,
                                  (_ffi.ptr(self.plan.image_u8), _ffi.ptr(x), npx), 0.0, npx * (3 + 16), (x,))
        return Act(x, S, S, 8, is_image=True)

    def _stem(self, x, name, cout, stride, ph, pw, act, bn):
        """The first layer straight from the image (float32 or uint8): MobileNetV2's Conv1 (3x3 stride 2 -> 32) and
        VGG16's conv1_1 (3x3 stride 1 -> 64)."""
        Ho, Wo = _out_size(x.H, 3, stride, 1, ph), _out_size(x.W, 3, stride, 1, pw)
        w, b = self.m._packed_conv(name, bn, 3)
        out = self._buf(Ho, Wo, cout)
        args = (_ffi.ptr(self.plan.image), _ffi.ptr(w), _ffi.ptr(b), _ffi.ptr(out), self.B, x.H, x.W, cout, Ho, Wo,
                stride, ph[0], pw[0], act)
        nbytes = self.B * (x.H * x.W * 3 * 4 + Ho * Wo * cout * 2) + 27 * cout * 2
        meta = dict(w=w, bias=b, out=out, ph=ph, pw=pw, act=act, stride=stride)
        self.plan.steps.append(Step(name, "stem", self.lib.ssd_stem_conv3x3, args, 2.0 * self.B * Ho * Wo * 27 * cout,
                                    nbytes, (w, b, out), dict(meta, x=self.plan.image)))
        self.plan.first_u8 = Step(name + "_u8", "stem", self.lib.ssd_stem_conv3x3_u8,
                                  (_ffi.ptr(self.plan.image_u8),) + args[1:], 2.0 * self.B * Ho * Wo * 27 * cout,
                                  nbytes - self.B * x.H * x.W * 9, (w, b, out), dict(meta, x=self.plan.image_u8))
        return Act(out, Ho, Wo, cout)

    # first layer of a training plan through the first-layer kernel ("1") or the generic tensor-map path ("0")
    TRAIN_STEM = os.environ.get("SSD_B200_TRAIN_STEM", "1") not in ("0", "")

    def _emit_conv(self, name, x: Act, w: torch.Tensor, bias, cout, k, stride, dilation, ph, pw, act,
                   residual: Optional[Act], out0, out1=None, out_f32=0, split=None, strides=None, real_cin=None):
        Ho, Wo = _out_size(x.H, k, stride, dilation, ph), _out_size(x.W, k, stride, dilation, pw)
        d = ConvDesc()
        d.inp, d.weight, d.bias = x.t.data_ptr(), w.data_ptr(), (bias.data_ptr() if bias is not None else None)
        d.residual = residual.t.data_ptr() if residual is not None else None
        d.out0, d.out1 = out0.data_ptr() if isinstance(out0, torch.Tensor) else out0, out1
        d.B, d.H, d.W, d.Cin = self.B, x.H, x.W, x.C
        d.Ho, d.Wo, d.Cout = Ho, Wo, cout
        d.KH = d.KW = k
        d.stride, d.dilation, d.pad_top, d.pad_left = stride, dilation, ph[0], pw[0]
        d.act, d.out_f32 = act, out_f32
        d.split = cout if split is None else split
        if strides is None:
            strides = (Ho * Wo * cout, cout, 0, 0)
        d.img_stride0, d.pix_stride0, d.img_stride1, d.pix_stride1 = strides
        cin = real_cin or x.C
        macs = self.B * Ho * Wo * k * k * cin * cout
        nbytes = self.B * x.H * x.W * cin * 2 + k * k * cin * cout * 2 + \
            self.B * Ho * Wo * cout * (4 if out_f32 else 2) + (self.B * Ho * Wo * cout * 2 if residual is not None else 0)
        meta = dict(x=x.t, w=w, bias=bias, res=residual.t if residual is not None else None, out0=out0, out1=out1,
                    k=k, stride=stride, dilation=dilation, ph=ph, pw=pw, act=act, Ho=Ho, Wo=Wo, cout=cout,
                    split=d.split, out_f32=out_f32, strides=strides)
        if (self.TRAIN_STEM and x.is_image and k == 3 and dilation == 1 and residual is None and not out_f32 and d.split == cout
                and (stride, cout) in ((2, 32), (1, 64)) and ph[0] <= 1 and pw[0] <= 1 and isinstance(out0, torch.Tensor)):
            # first layer of a TRAINING plan: same tensors as the tensor-map path (fp16 image padded to 8 channels, OHWI
            # weights padded to 8 input channels -- what its filter gradient reads), computed by the first-layer kernel
            # instead of nine 64-channel k-blocks for 3 real channels
            args = (_ffi.ptr(x.t), _ffi.ptr(w), _ffi.ptr(bias) if bias is not None else None, _ffi.ptr(out0), self.B, x.H, x.W,
                    cout, Ho, Wo, stride, ph[0], pw[0], act)
            self.plan.steps.append(Step(name, "conv", self.lib.ssd_stem_conv3x3_f16c8, args, 2.0 * macs, nbytes,
                                        (d, w, bias, x.t, out0), meta))
            return Ho, Wo
        self.plan.steps.append(Step(name, "conv", self.lib.ssd_conv2d, (C.byref(d),), 2.0 * macs, nbytes,
                                    (d, w, bias, x.t, out0), meta))
        return Ho, Wo

    def conv(self, x, name, cout, k=1, stride=1, pad="same", dilation=1, act=ACT_NONE, bn=None, use_bias=True,
             residual=None, init=None, l2=False, tap=False):
        ph, pw = _resolve_pads(x.H, x.W, k, stride, dilation, pad)
        if x.t is None:                                  # first layer, fed by the fp32 image
            if (k == 3 and dilation == 1 and residual is None and (stride, cout) in ((2, 32), (1, 64)) and not self.no_stem
                    and ph[0] <= 1 and pw[0] <= 1):
                return self._stem(x, name, cout, stride, ph, pw, act, bn)
            x = self._image_as_f16()
        w, b = self.m._packed_conv(name, bn, x.C)
        Ho, Wo = _out_size(x.H, k, stride, dilation, ph), _out_size(x.W, k, stride, dilation, pw)
        out = self._buf(Ho, Wo, cout)
        real_cin = self.m.weights[name + "/kernel"].shape[2]
        last = self.plan.steps[-1] if self.plan.steps else None
        if (self.FUSE_DW not in ("0", "") and last is not None and last.kind == "dw" and last.meta["out"] is x.t
                and (self.FUSE_DW in ("1", "all") or last.meta["stride"] == 2) and k == 1 and stride == 1
                and dilation == 1 and cout <= 256 and cout % 8 == 0 and type(self) is _PlanBuilder):
            # depthwise 3x3 -> 1x1 projection as ONE launch (ssd_dwproj): the depthwise output stays on chip
            dm = last.meta
            if self._try_stemblock(name, x, w, b, cout, act, residual, out, Ho, Wo):
                return Act(out, Ho, Wo, cout)
            if self._try_irblock(name, x, w, b, cout, act, residual, out, Ho, Wo):
                return Act(out, Ho, Wo, cout)
            d = DwProjDesc()
            d.inp, d.dw_weight, d.dw_bias = dm["x"].data_ptr(), dm["w"].data_ptr(), dm["bias"].data_ptr()
            d.proj_weight, d.proj_bias = w.data_ptr(), b.data_ptr()
            d.residual = residual.t.data_ptr() if residual is not None else None
            d.out = out.data_ptr()
            d.B, d.H, d.W, d.C = self.B, dm["x"].shape[1], dm["x"].shape[2], x.C
            d.Ho, d.Wo, d.Cout = Ho, Wo, cout
            d.stride, d.pad_top, d.pad_left, d.dw_act, d.act = dm["stride"], dm["ph"][0], dm["pw"][0], dm["act"], act
            if self.B * Ho * Wo < 128 * self.FUSE_MIN_TILES or not self.lib.ssd_dwproj_supported(C.byref(d)):
                self._emit_conv(name, x, w, b, cout, k, stride, dilation, ph, pw, act, residual, out, real_cin=real_cin)
                return Act(out, Ho, Wo, cout)
            self.plan.steps.pop()
            nbytes = self.B * dm["x"].shape[1] * dm["x"].shape[2] * x.C * 2 + 9 * x.C * 2 + x.C * cout * 2 + \
                self.B * Ho * Wo * cout * 2 * (2 if residual is not None else 1)
            flops = 2.0 * self.B * Ho * Wo * x.C * (9 + cout)
            meta = dict(x=dm["x"], dw_w=dm["w"], dw_bias=dm["bias"], dw_stride=dm["stride"], dw_ph=dm["ph"], dw_pw=dm["pw"],
                        dw_act=dm["act"], w=w, bias=b, res=residual.t if residual is not None else None, out0=out, act=act,
                        Ho=Ho, Wo=Wo, cout=cout, dw_name=last.name)
            self.plan.steps.append(Step(name, "dwproj", self.lib.ssd_dwproj, (C.byref(d),), flops, nbytes,
                                        (d, w, b, dm["x"], dm["w"], dm["bias"], out), meta))
            return Act(out, Ho, Wo, cout)
        self._emit_conv(name, x, w, b, cout, k, stride, dilation, ph, pw, act, residual, out, real_cin=real_cin)
        if tap:
            self.plan.steps[-1].meta["tap"] = True
        return Act(out, Ho, Wo, cout)

    # first layer 3x3 s2 -> depthwise 3x3 -> project 1x1 (MobileNetV2 Conv1 + block 0) as ONE launch (ssd_stem_dwproj): "1" / "0"
    FUSE_STEM = os.environ.get("SSD_B200_FUSE_STEM", "1")

    def _try_stemblock(self, name, x, w, b, cout, act, residual, out, Ho, Wo) -> bool:
        """Called while emitting a 1x1 projection with a depthwise step last in the plan: when that depthwise layer
        (stride 1) reads the first-layer kernel's output and nothing else does, the three layers become one launch that
        reads the image and writes the projection (the 32-channel stem output never reaches HBM)."""
        steps = self.plan.steps
        if self.FUSE_STEM in ("0", "") or len(steps) != 2 or type(self) is not _PlanBuilder or residual is not None:
            return False
        dws, stem = steps[-1], steps[-2]
        dm, sm = dws.meta, stem.meta
        if not (stem.kind == "stem" and sm["stride"] == 2 and sm["out"] is dm["x"] and dm["stride"] == 1
                and tuple(dm["ph"]) == (1, 1) and tuple(dm["pw"]) == (1, 1)):
            return False
        Hs, Ws = dm["x"].shape[1], dm["x"].shape[2]
        descs = []
        for is_u8, img in ((0, self.plan.image), (1, self.plan.image_u8)):
            d = StemDwProjDesc()
            d.image, d.image_u8 = img.data_ptr(), is_u8
            d.stem_weight, d.stem_bias = sm["w"].data_ptr(), sm["bias"].data_ptr()
            d.dw_weight, d.dw_bias = dm["w"].data_ptr(), dm["bias"].data_ptr()
            d.proj_weight, d.proj_bias, d.out = w.data_ptr(), b.data_ptr(), out.data_ptr()
            d.B, d.H, d.W, d.Hs, d.Ws = self.B, img.shape[1], img.shape[2], Hs, Ws
            d.Cmid, d.Cout = x.C, cout
            d.pad_top, d.pad_left = sm["ph"][0], sm["pw"][0]
            d.stem_act, d.dw_act, d.act = sm["act"], dm["act"], act
            if not self.lib.ssd_stem_dwproj_supported(C.byref(d)):
                return False
            descs.append(d)
        steps.pop(); steps.pop()
        B, H, W = self.B, self.plan.image.shape[1], self.plan.image.shape[2]
        flops = 2.0 * B * Hs * Ws * x.C * (27 + 9 + cout)
        wbytes = (27 * x.C + 9 * x.C + x.C * cout) * 2
        meta = dict(stem_w=sm["w"], stem_bias=sm["bias"], stem_act=sm["act"], ph=sm["ph"], pw=sm["pw"], stride=2,
                    dw_w=dm["w"], dw_bias=dm["bias"], dw_act=dm["act"], w=w, bias=b, res=None, out0=out, act=act,
                    Ho=Ho, Wo=Wo, cout=cout, stem_name=stem.name, dw_name=dws.name)
        keep = (w, b, sm["w"], sm["bias"], dm["w"], dm["bias"], out)
        steps.append(Step(name, "stemblock", self.lib.ssd_stem_dwproj, (C.byref(descs[0]),), flops,
                          B * (H * W * 3 * 4 + Hs * Ws * cout * 2) + wbytes, keep + (descs[0],), dict(meta, x=self.plan.image)))
        self.plan.first_u8 = Step(name + "_u8", "stemblock", self.lib.ssd_stem_dwproj, (C.byref(descs[1]),), flops,
                                  B * (H * W * 3 + Hs * Ws * cout * 2) + wbytes, keep + (descs[1],),
                                  dict(meta, x=self.plan.image_u8))
        return True

    # expand 1x1 -> depthwise 3x3 -> project 1x1 as ONE launch (ssd_irblock): "1" / "0"
    FUSE_IR = os.environ.get("SSD_B200_FUSE_IR", "1")

    def _try_irblock(self, name, x, w, b, cout, act, residual, out, Ho, Wo) -> bool:
        """Called while emitting a block's 1x1 projection with a depthwise step last in the plan: when the step
        before that is the block's 1x1 expansion (consumed by the depthwise layer only), the three become one launch."""
        steps = self.plan.steps
        if self.FUSE_IR in ("0", "") or len(steps) < 2 or type(self) is not _PlanBuilder:
            return False
        dws, exp = steps[-1], steps[-2]
        dm, em = dws.meta, exp.meta
        if not (exp.kind == "conv" and em["k"] == 1 and em["stride"] == 1 and em["res"] is None and not em.get("tap")
                and not em["out_f32"] and em["out0"] is dm["x"] and "head" not in em):
            return False
        xin = em["x"]
        d = IrBlockDesc()
        d.inp, d.exp_weight, d.exp_bias = xin.data_ptr(), em["w"].data_ptr(), em["bias"].data_ptr()
        d.dw_weight, d.dw_bias = dm["w"].data_ptr(), dm["bias"].data_ptr()
        d.proj_weight, d.proj_bias = w.data_ptr(), b.data_ptr()
        d.residual = residual.t.data_ptr() if residual is not None else None
        d.out = out.data_ptr()
        d.B, d.H, d.W, d.Cin, d.Cexp = self.B, xin.shape[1], xin.shape[2], xin.shape[3], x.C
        d.Ho, d.Wo, d.Cout = Ho, Wo, cout
        d.stride, d.pad_top, d.pad_left = dm["stride"], dm["ph"][0], dm["pw"][0]
        d.exp_act, d.dw_act, d.act = em["act"], dm["act"], act
        if not self.lib.ssd_irblock_supported(C.byref(d)):
            return False
        steps.pop(); steps.pop()
        B, H, W, cin, cexp = self.B, d.H, d.W, d.Cin, d.Cexp
        nbytes = B * H * W * cin * 2 + (cin * cexp + 9 * cexp + cexp * cout) * 2 + \
            B * Ho * Wo * cout * 2 * (2 if residual is not None else 1)
        flops = 2.0 * B * (H * W * cin * cexp + Ho * Wo * cexp * (9 + cout))
        meta = dict(x=xin, exp_w=em["w"], exp_bias=em["bias"], exp_act=em["act"], dw_w=dm["w"], dw_bias=dm["bias"],
                    dw_stride=dm["stride"], dw_ph=dm["ph"], dw_pw=dm["pw"], dw_act=dm["act"], w=w, bias=b,
                    res=residual.t if residual is not None else None, out0=out, act=act, Ho=Ho, Wo=Wo, cout=cout,
                    exp_name=exp.name, dw_name=dws.name)
        steps.append(Step(name, "irblock", self.lib.ssd_irblock, (C.byref(d),), flops, nbytes,
                          (d, w, b, xin, em["w"], em["bias"], dm["w"], dm["bias"], out), meta))
        return True

    def dw(self, x, name, stride=1, act=ACT_RELU6, bn=None):
        ph, pw = _resolve_pads(x.H, x.W, 3, stride, 1, "same" if stride == 1 else "correct")
        Ho, Wo = _out_size(x.H, 3, stride, 1, ph), _out_size(x.W, 3, stride, 1, pw)
        w, b = self.m._packed_dw(name, bn)
        out = self._buf(Ho, Wo, x.C)
        args = (_ffi.ptr(x.t), _ffi.ptr(w), _ffi.ptr(b), _ffi.ptr(out), self.B, x.H, x.W, x.C, Ho, Wo, stride,
                ph[0], pw[0], act)
        nbytes = self.B * (x.H * x.W + Ho * Wo) * x.C * 2 + 9 * x.C * 2
        self.plan.steps.append(Step(name, "dw", self.lib.ssd_depthwise3x3, args, 2.0 * self.B * Ho * Wo * 9 * x.C,
                                    nbytes, (w, b, x.t, out),
                                    dict(x=x.t, w=w, bias=b, out=out, stride=stride, ph=ph, pw=pw, act=act)))
        return Act(out, Ho, Wo, x.C)

    def maxpool(self, x, k, s):
        ph, pw = same_pad(x.H, k, s), same_pad(x.W, k, s)
        Ho, Wo = _out_size(x.H, k, s, 1, ph), _out_size(x.W, k, s, 1, pw)
        out = self._buf(Ho, Wo, x.C)
        args = (_ffi.ptr(x.t), _ffi.ptr(out), self.B, x.H, x.W, x.C, Ho, Wo, k, s, ph[0], pw[0])
        self.plan.steps.append(Step(f"pool{k}x{k}s{s}_{x.H}", "pool", self.lib.ssd_maxpool, args, 0.0,
                                    self.B * (x.H * x.W + Ho * Wo) * x.C * 2, (x.t, out),
                                    dict(x=x.t, out=out, k=k, stride=s, ph=ph, pw=pw)))
        return Act(out, Ho, Wo, x.C)

    def l2norm(self, x, name, scale_factor=20.0):
        scale = self.m._dev_f32(name + "/scale")
        out = self._buf(x.H, x.W, x.C)
        rows = self.B * x.H * x.W
        self.plan.steps.append(Step(name, "l2norm", self.lib.ssd_l2norm,
                                    (_ffi.ptr(x.t), _ffi.ptr(scale), _ffi.ptr(out), rows, x.C), 0.0,
                                    rows * x.C * 4, (scale, x.t, out), dict(x=x.t, scale=scale, out=out)))
        return Act(out, x.H, x.W, x.C)

    def head(self, taps: Sequence[Act], hp):
        """models/header.py:54-90: per tap ONE 3x3 SAME convolution producing the
        label and box channels together, written straight into the concatenated
        ``[B,N,L]`` logits and ``[B,N,4]`` deltas (HeadWrapper :46-51 is free)."""
        L = int(hp["total_labels"])
        counts = [t.H * t.W * (len(hp["aspect_ratios"][i]) + 1) for i, t in enumerate(taps)]
        N = sum(counts)
        logits = torch.empty((self.B, N, L), dtype=torch.float32, device=self.dev)
        deltas = torch.empty((self.B, N, 4), dtype=torch.float32, device=self.dev)
        off = 0
        for i, t in enumerate(taps):
            A = len(hp["aspect_ratios"][i]) + 1
            w, b = self.m._packed_head(i + 1, t.C)
            ph, pw = same_pad(t.H, 3, 1), same_pad(t.W, 3, 1)
            out0 = logits.data_ptr() + off * L * 4
            out1 = deltas.data_ptr() + off * 4 * 4
            self._emit_conv(f"{i + 1}_conv_head", t, w, b, A * (L + 4), 3, 1, 1, ph, pw, ACT_NONE, None, out0, out1,
                            out_f32=1, split=A * L, strides=(N * L, A * L, N * 4, A * 4))
            self.plan.steps[-1].meta["head"] = (off, counts[i], A)
            off += counts[i]
        self.plan.logits, self.plan.deltas, self.plan.taps = logits, deltas, list(taps)
        return deltas, logits


class _TrainPlanBuilder(_PlanBuilder):
    """Training-mode interpretation of a graph (Keras ``fit``: ``training=True``): BatchNormalization is NOT
    folded -- a BN'd layer becomes  convolution (raw fp16 kernel, no bias, no activation) -> ``ssd_bn_train_fwd``
    (batch statistics, affine, activation, shortcut add).  Layers without BN are emitted exactly as in the
    inference plan.  The first layer always reads the fp16 copy of the image (its filter gradient needs it)."""

    BN_MOMENTUM = 0.999      # keras_applications.mobilenet_v2: BatchNormalization(momentum=0.999)

    def __init__(self, model: "SSDModel", B: int):
        super().__init__(model, B)
        self.bn_ws: Optional[torch.Tensor] = None

    def _bn_workspace(self, C: int) -> torch.Tensor:
        need = int(self.lib.ssd_bn_workspace_bytes(int(C)))
        if self.bn_ws is None or self.bn_ws.numel() < need:
            self.bn_ws = _ffi.workspace(need)          # one shared scratch (launches are stream-ordered)
        return self.bn_ws

    def _emit_bn(self, bn: str, x_pre: Act, act: int, residual: Optional[Act]) -> Act:
        m = self.m
        C_ = x_pre.C
        gamma, beta = m._train_f32(bn + "/gamma"), m._train_f32(bn + "/beta")
        mm, mv = m._train_f32(bn + "/moving_mean"), m._train_f32(bn + "/moving_variance")
        save = torch.zeros(2 * C_, dtype=torch.float32, device=self.dev)
        out = self._buf(x_pre.H, x_pre.W, C_)
        ws = self._bn_workspace(1280)
        M = self.B * x_pre.H * x_pre.W
        args = (_ffi.ptr(x_pre.t), _ffi.ptr(gamma), _ffi.ptr(beta), _ffi.ptr(mm), _ffi.ptr(mv), M, C_, BN_EPS,
                self.BN_MOMENTUM, act, _ffi.ptr(residual.t) if residual is not None else None, _ffi.ptr(out),
                _ffi.ptr(save), _ffi.ptr(ws), ws.numel())
        nbytes = M * C_ * 2 * (3 + (1 if residual is not None else 0))
        self.plan.steps.append(Step(bn, "bn", self.lib.ssd_bn_train_fwd, args, 0.0, nbytes,
                                    (gamma, beta, mm, mv, save, out, ws, x_pre.t),
                                    dict(x=x_pre.t, out=out, gamma=gamma, beta=beta, save=save, act=act, ws=ws,
                                         res=residual.t if residual is not None else None, M=M, C=C_)))
        return Act(out, x_pre.H, x_pre.W, C_)

    def conv(self, x, name, cout, k=1, stride=1, pad="same", dilation=1, act=ACT_NONE, bn=None, use_bias=True,
             residual=None, init=None, l2=False, tap=False):
        if x.t is None:
            x = self._image_as_f16()
        if bn is None:
            return super().conv(x, name, cout, k=k, stride=stride, pad=pad, dilation=dilation, act=act, bn=None,
                                use_bias=use_bias, residual=residual)
        ph, pw = _resolve_pads(x.H, x.W, k, stride, dilation, pad)
        w = self.m._train_conv_kernel(name, x.C)
        Ho, Wo = _out_size(x.H, k, stride, dilation, ph), _out_size(x.W, k, stride, dilation, pw)
        pre = self._buf(Ho, Wo, cout)
        real_cin = self.m.weights[name + "/kernel"].shape[2]
        self._emit_conv(name, x, w, None, cout, k, stride, dilation, ph, pw, ACT_NONE, None, pre, real_cin=real_cin)
        return self._emit_bn(bn, Act(pre, Ho, Wo, cout), act, residual)

    def dw(self, x, name, stride=1, act=ACT_RELU6, bn=None):
        ph, pw = _resolve_pads(x.H, x.W, 3, stride, 1, "same" if stride == 1 else "correct")
        Ho, Wo = _out_size(x.H, 3, stride, 1, ph), _out_size(x.W, 3, stride, 1, pw)
        w = self.m._train_dw_kernel(name)
        pre = self._buf(Ho, Wo, x.C)
        args = (_ffi.ptr(x.t), _ffi.ptr(w), None, _ffi.ptr(pre), self.B, x.H, x.W, x.C, Ho, Wo, stride, ph[0], pw[0],
                ACT_NONE)
        nbytes = self.B * (x.H * x.W + Ho * Wo) * x.C * 2 + 9 * x.C * 2
        self.plan.steps.append(Step(name, "dw", self.lib.ssd_depthwise3x3, args, 2.0 * self.B * Ho * Wo * 9 * x.C,
                                    nbytes, (w, x.t, pre),
                                    dict(x=x.t, w=w, bias=None, out=pre, stride=stride, ph=ph, pw=pw, act=ACT_NONE,
                                         Ho=Ho, Wo=Wo)))
        return self._emit_bn(bn, Act(pre, Ho, Wo, x.C), act, None)


# ----------------------------------------------------------------- the model --
class SSDModel(object):
    """What ``get_model(hyper_params)`` returns: the object ``trainer.py`` /
    ``predictor.py`` call (``model(x)``, ``load_weights``, ``predict``).

    ``model(images)`` -> ``(pred_deltas [B,N,4], pred_labels [B,N,L])`` with
    ``pred_labels`` the softmax probabilities (header.py:88-90), both float32
    CUDA tensors.  Weights live in ``self.weights`` as float32 NumPy arrays in
    Keras layouts under Keras variable names, so a converted ``.h5`` drops in.
    Inference arithmetic: BatchNorm folded into the preceding kernel, fp16
    storage, fp32 accumulation, fp32 head outputs."""

    def __init__(self, backbone: str, hyper_params: Dict[str, Any], seed: int = 0):
        if "total_labels" not in hyper_params:
            raise KeyError("hyper_params['total_labels'] must be set by the caller (trainer.py:63-64)")
        self.backbone = "vgg16" if backbone.startswith("vgg16") else backbone
        self.hyper_params = hyper_params
        self.img_size = int(hyper_params["img_size"])
        self.total_labels = int(hyper_params["total_labels"])
        tr = _ParamTracer(np.random.default_rng(seed))
        taps = GRAPHS[self.backbone](tr, Act(None, self.img_size, self.img_size, 3), hyper_params)
        fm = [t.H for t in taps]
        if fm != list(hyper_params["feature_map_shapes"]):
            raise ValueError(f"graph yields feature maps {fm}, hyper_params say {hyper_params['feature_map_shapes']}")
        tr.head(taps, hyper_params)
        self.weights: Dict[str, np.ndarray] = tr.weights
        self.l2_kernels = tr.l2_kernels
        self.macs_per_image = tr.macs
        self.n_anchors = sum(t.H * t.W * (len(hyper_params["aspect_ratios"][i]) + 1) for i, t in enumerate(taps))
        self._packed: Dict[str, Tuple[torch.Tensor, Optional[torch.Tensor]]] = {}
        self._plans: Dict[Tuple[int, int], Plan] = {}
        self._train_vars: Dict[str, torch.Tensor] = {}       # un-folded device variables of the training plans
        self._train_plans: Dict[int, Plan] = {}
        self.has_batchnorm = any(k.endswith("/gamma") for k in self.weights)

    # -- weights ---------------------------------------------------------------
    def set_weights(self, weights: Dict[str, np.ndarray]) -> None:
        for k, v in weights.items():                     # validate everything before touching any state
            if k not in self.weights:
                raise KeyError(f"unknown variable {k!r}")
            if tuple(v.shape) != tuple(self.weights[k].shape):
                raise ValueError(f"{k}: shape {v.shape} != {self.weights[k].shape}")
        for k, v in weights.items():
            self.weights[k] = np.ascontiguousarray(v, dtype=np.float32)
        # ``load_weights`` after ``compile`` (trainer.py:91-99 loads AFTER compiling): the trainer's device-side
        # variables are rebuilt from the new host values (optimizer moments restart at zero)
        trainer = getattr(self, "trainer", None)
        self.trainer = None
        self._invalidate()
        if trainer is not None:
            from tf_ssd_b200.models.train_engine import Trainer
            self.trainer = Trainer(self, **trainer.init_kwargs)

    def get_weights(self) -> Dict[str, np.ndarray]:
        return dict(self.weights)

    def save_weights(self, path: str) -> None:
        np.savez(path, **self.weights)

    def load_weights(self, path: str) -> None:
        """``model.load_weights`` (trainer.py:99, predictor.py:83).  ``.npz`` keyed
        by Keras variable names; ``h5py`` is not available in this image."""
        with np.load(path) as z:
            self.set_weights({k: z[k] for k in z.files})

    def _invalidate(self) -> None:
        self._packed.clear()
        self._plans.clear()
        self._version = getattr(self, "_version", 0) + 1     # DecoderModel drops its captured graphs when this moves
        if getattr(self, "trainer", None) is None:       # a live trainer owns the training-mode variables
            self._train_vars.clear()
            self._train_plans.clear()

    def _bn_fold(self, bn: Optional[str], cout: int, bias: Optional[np.ndarray]):
        b = bias if bias is not None else np.zeros(cout, np.float32)
        if bn is None:
            return np.ones(cout, np.float32), b
        w = self.weights
        scale = w[bn + "/gamma"] / np.sqrt(w[bn + "/moving_variance"] + np.float32(BN_EPS))
        return scale.astype(np.float32), (b * scale + (w[bn + "/beta"] - w[bn + "/moving_mean"] * scale)).astype(np.float32)

    def _upload(self, key, kernel_ohwi: np.ndarray, bias: np.ndarray):
        dev = _ffi.require_cuda()
        wt = torch.from_numpy(np.ascontiguousarray(kernel_ohwi)).to(dev).to(torch.float16).contiguous()
        bt = torch.from_numpy(np.ascontiguousarray(bias, dtype=np.float32)).to(dev)
        self._packed[key] = (wt, bt)
        return wt, bt

    def _packed_conv(self, name: str, bn: Optional[str], cin_buf: int):
        """HWIO float32 -> BN-folded OHWI fp16 (Cin zero-padded to the buffer's channel count)."""
        if name in self._packed:
            return self._packed[name]
        k = self.weights[name + "/kernel"]                          # [kh,kw,Cin,Cout]
        scale, bias = self._bn_fold(bn, k.shape[3], self.weights.get(name + "/bias"))
        k = k * scale.reshape(1, 1, 1, -1)
        if cin_buf != k.shape[2]:
            k = np.concatenate([k, np.zeros(k.shape[:2] + (cin_buf - k.shape[2], k.shape[3]), np.float32)], axis=2)
        return self._upload(name, k.transpose(3, 0, 1, 2), bias)

    def _packed_dw(self, name: str, bn: Optional[str]):
        if name in self._packed:
            return self._packed[name]
        k = self.weights[name + "/depthwise_kernel"][:, :, :, 0]    # [3,3,C]
        scale, bias = self._bn_fold(bn, k.shape[2], None)
        return self._upload(name, k * scale.reshape(1, 1, -1), bias)

    def _packed_head(self, index: int, cin_buf: int):
        key = f"{index}_conv_head"
        if key in self._packed:
            return self._packed[key]
        kl, kb = self.weights[f"{index}_conv_label_output/kernel"], self.weights[f"{index}_conv_boxes_output/kernel"]
        k = np.concatenate([kl, kb], axis=3)
        bias = np.concatenate([self.weights[f"{index}_conv_label_output/bias"],
                               self.weights[f"{index}_conv_boxes_output/bias"]])
        return self._upload(key, k.transpose(3, 0, 1, 2), bias)

    # -- training-mode variables (BatchNorm not folded; see _TrainPlanBuilder) ----
    def _train_f32(self, key: str) -> torch.Tensor:
        if key not in self._train_vars:
            self._train_vars[key] = torch.from_numpy(np.ascontiguousarray(self.weights[key], np.float32)).to(_ffi.require_cuda())
        return self._train_vars[key]

    def _train_conv_kernel(self, name: str, cin_buf: int) -> torch.Tensor:
        key = name + "/kernel"
        if key not in self._train_vars:
            k = self.weights[key]                                     # HWIO
            if cin_buf != k.shape[2]:
                k = np.concatenate([k, np.zeros(k.shape[:2] + (cin_buf - k.shape[2], k.shape[3]), np.float32)], axis=2)
            t = torch.from_numpy(np.ascontiguousarray(k.transpose(3, 0, 1, 2))).to(_ffi.require_cuda())
            self._train_vars[key] = t.to(torch.float16).contiguous()  # OHWI fp16
        return self._train_vars[key]

    def _train_dw_kernel(self, name: str) -> torch.Tensor:
        key = name + "/depthwise_kernel"
        if key not in self._train_vars:
            k = self.weights[key][:, :, :, 0]                         # [3,3,C]
            self._train_vars[key] = torch.from_numpy(np.ascontiguousarray(k)).to(_ffi.require_cuda()).to(torch.float16).contiguous()
        return self._train_vars[key]

    def train_plan(self, B: int) -> Plan:
        """Launch plan of the TRAINING-mode forward (``model(x, training=True)`` inside Keras ``fit``)."""
        if B not in self._train_plans:
            _ffi.check_device()
            # graphs without BatchNorm train on the inference launch list, but layer by layer: the backward pass needs every
            # layer as its own step (no fused tail)
            pb = _TrainPlanBuilder(self, B) if self.has_batchnorm else _PlanBuilder(self, B)
            pb.no_stem = True
            taps = GRAPHS[self.backbone](pb, pb.input(), self.hyper_params)
            pb.head(taps, self.hyper_params)
            if not self.has_batchnorm:
                pb.plan.hoist_heads()
            self._train_plans[B] = pb.plan
        return self._train_plans[B]

    def _dev_f32(self, key: str) -> torch.Tensor:
        if key not in self._packed:
            self._packed[key] = (torch.from_numpy(self.weights[key]).to(_ffi.require_cuda()), None)
        return self._packed[key][0]

    # -- plans -----------------------------------------------------------------
    def plan(self, B: int, slot: int = 0) -> Plan:
        """Launch plan for batch size ``B``.  ``slot`` selects an independent set of
        activation buffers (double-buffered pipelines keep two in flight)."""
        if (B, slot) not in self._plans:
            _ffi.check_device()
            pb = _PlanBuilder(self, B)
            taps = GRAPHS[self.backbone](pb, pb.input(), self.hyper_params)
            pb.head(taps, self.hyper_params)
            pb.plan.hoist_heads()
            pb.plan.fuse_tail(pb.lib)
            self._plans[(B, slot)] = pb.plan
        return self._plans[(B, slot)]

    @staticmethod
    def _is_u8(images: Any) -> bool:
        return (images.dtype == torch.uint8) if isinstance(images, torch.Tensor) else (np.asarray(images).dtype == np.uint8)

    def _to_image_buffer(self, plan: Plan, images: Any) -> bool:
        """Copies a batch into the plan's input buffer; returns True when it was a uint8 batch (NHWC uint8 = the image
        before ``convert_image_dtype``, utils/data_utils.py:33-37: the first layer then converts on the fly)."""
        u8 = self._is_u8(images)
        if isinstance(images, torch.Tensor):
            x = images
        else:
            x = torch.from_numpy(np.ascontiguousarray(images, np.uint8 if u8 else np.float32))
        if tuple(x.shape) != tuple(plan.image.shape):
            raise ValueError(f"expected images {tuple(plan.image.shape)} (NHWC float32 or uint8), got {tuple(x.shape)}")
        if u8:
            plan.image_u8.copy_(x, non_blocking=True)
        else:
            plan.image.copy_(x.to(torch.float32), non_blocking=True)
        return u8

    def forward_logits(self, images: Any) -> Tuple[torch.Tensor, torch.Tensor]:
        """``(pred_deltas, logits)`` -- the pre-softmax head outputs (what Keras feeds
        the cross-entropy inside ``fit`` and what the fused decoder consumes)."""
        B = int(images.shape[0])
        plan = self.plan(B)
        plan.run(u8=self._to_image_buffer(plan, images))
        return plan.deltas, plan.logits

    def __call__(self, images: Any, training: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        deltas, logits = self.forward_logits(images)
        B, N, L = logits.shape
        probs = torch.empty_like(logits)
        _ffi.check(_ffi.lib().ssd_softmax(_ffi.ptr(logits), B * N, L, _ffi.ptr(probs), _ffi.stream()), "ssd_softmax")
        return deltas.clone(), probs

    # -- training (trainer.py:91-127) ------------------------------------------
    def compile(self, optimizer: Any = None, loss: Any = None, **kwargs: Any) -> None:
        """``model.compile(optimizer=Adam(learning_rate=1e-3), loss=[loc_loss_fn, conf_loss_fn])``
        (trainer.py:91-94).  ``loss`` are the bound methods of a ``CustomLoss``; its
        ``neg_pos_ratio`` / ``loc_loss_alpha`` parameterise the fused loss kernels."""
        from tf_ssd_b200.models.train_engine import Trainer
        lr = float(getattr(optimizer, "learning_rate", 1e-3) if optimizer is not None else 1e-3)
        owner = getattr(loss[0], "__self__", None) if loss else None
        ratio = float(getattr(owner, "neg_pos_ratio", 3.0))
        alpha = float(getattr(owner, "loc_loss_alpha", 1.0))
        extra = {k: kwargs[k] for k in ("loss_scale", "dynamic_loss_scale", "scale_check_every", "scale_growth_interval",
                                        "use_cuda_graph") if k in kwargs}      # mixed-precision knobs (no Keras counterpart)
        self.trainer = Trainer(self, learning_rate=lr, neg_pos_ratio=ratio, loc_loss_alpha=alpha,
                               beta_1=float(getattr(optimizer, "beta_1", 0.9)), beta_2=float(getattr(optimizer, "beta_2", 0.999)),
                               epsilon=float(getattr(optimizer, "epsilon", 1e-7)), **extra)

    def train_on_batch(self, x: Any, y: Tuple[Any, Any], learning_rate: Optional[float] = None) -> Dict[str, float]:
        if getattr(self, "trainer", None) is None:
            raise RuntimeError("call model.compile(optimizer=..., loss=[...]) first")
        return self.trainer.train_on_batch(x, y, learning_rate)

    def fit(self, data: Iterable[Any], steps_per_epoch: Optional[int] = None, validation_data: Optional[Iterable[Any]] = None,
            validation_steps: Optional[int] = None, epochs: int = 1, callbacks: Optional[Sequence[Any]] = None,
            verbose: int = 0) -> Dict[str, List[float]]:
        """``model.fit(generator, steps_per_epoch=, validation_data=, validation_steps=, epochs=, callbacks=)``
        (trainer.py:120-127).  ``data`` yields ``(img, (actual_deltas, actual_labels))`` like
        ``train_utils.generator``.  Callbacks may implement ``on_epoch_begin(epoch, logs)`` returning a
        learning rate (the reference's ``LearningRateScheduler(scheduler)``) and ``on_epoch_end(epoch, logs)``."""
        if getattr(self, "trainer", None) is None:
            raise RuntimeError("call model.compile(optimizer=..., loss=[...]) first")
        history: Dict[str, List[float]] = {"loss": [], "val_loss": []}
        for cb in callbacks or ():
            if hasattr(cb, "set_model"):
                cb.set_model(self)
        it = iter(data)
        vit = iter(validation_data) if validation_data is not None else None
        for epoch in range(epochs):
            lr = None
            for cb in callbacks or ():
                hook = getattr(cb, "on_epoch_begin", None)
                r = hook(epoch, {}) if hook else None
                if isinstance(r, float):
                    lr = r
            tot, n = None, 0                             # the running loss stays on the device: one read per epoch
            for _ in range(steps_per_epoch or 1):
                img, targets = next(it)
                step_loss = self.trainer.train_step(img, targets, lr)
                tot = step_loss if tot is None else tot + step_loss
                n += 1
            logs = {"loss": float(tot) / max(n, 1) if tot is not None else 0.0}
            if self.trainer.dynamic_loss_scale:
                self.trainer.update_loss_scale()
            if verbose:
                print(f"Epoch {epoch + 1}/{epochs} - loss: {logs['loss']:.4f}", flush=True)
            if vit is not None:
                vt, vn = 0.0, 0
                for _ in range(validation_steps or 1):
                    img, targets = next(vit)
                    out = self.trainer.evaluate_batch(img, targets)
                    vt += out["loss"]; vn += 1
                logs["val_loss"] = vt / max(vn, 1)
                history["val_loss"].append(logs["val_loss"])
            history["loss"].append(logs["loss"])
            for cb in callbacks or ():
                hook = getattr(cb, "on_epoch_end", None)
                if hook:
                    hook(epoch, logs)
        self.trainer.sync_weights_to_host()
        return history

    def predict(self, data: Iterable[Any], steps: Optional[int] = None, verbose: int = 0):
        outs_d, outs_p = [], []
        for i, batch in enumerate(data):
            if steps is not None and i >= steps:
                break
            img = batch[0] if isinstance(batch, (tuple, list)) else batch
            d, p = self(img)
            outs_d.append(d.cpu().numpy())
            outs_p.append(p.cpu().numpy())
        return np.concatenate(outs_d, 0), np.concatenate(outs_p, 0)


class DecoderModel(object):
    """``get_decoder_model`` result (models/decoder.py:96-108): images ->
    ``(boxes [B,200,4], labels [B,200], scores [B,200])``.

    The forward plan and the fused softmax+decode+NMS are captured into one
    CUDA graph per (batch size, slot, input dtype); the multibox heads are parallel
    branches of that graph.  Batches may be float32 in [0,1] (what the reference model
    receives, utils/data_utils.py:36) or uint8 (the image before
    ``convert_image_dtype``: a quarter of the host->device bytes; the conversion is
    fused into the first layer and the results are bit-identical).  ``predict`` keeps two slots in flight:
    the host->device image copy of batch i+1 and the device->host copy of
    batch i-1's detections overlap batch i's graph replay."""

    N_SLOTS = 2

    def __init__(self, base_model: SSDModel, decoder: Any, use_cuda_graph: bool = True):
        self.base_model = base_model
        self.decoder = decoder
        self.use_cuda_graph = use_cuda_graph
        self._state: Dict[Tuple[int, int], Dict[str, Any]] = {}
        self._copy_stream: Optional[Tuple[torch.cuda.Stream, torch.cuda.Stream]] = None

    def _prepare(self, B: int, slot: int = 0) -> Dict[str, Any]:
        version = getattr(self.base_model, "_version", 0)
        if version != getattr(self, "_seen_version", version):
            self._state.clear()                       # weights were re-packed (set_weights / BatchNorm training)
        self._seen_version = version
        if (B, slot) in self._state:
            return self._state[(B, slot)]
        plan = self.base_model.plan(B, slot)
        dev, T = plan.device, self.decoder.max_total_size
        st: Dict[str, Any] = {
            "plan": plan,
            # one packed result buffer so a single D2H brings everything back:
            # boxes [B,T,4] | labels [B,T] | scores [B,T] | int32 valid counts [B] (views of ``packed``)
            "packed": torch.empty((B * T * 6 + B,), dtype=torch.float32, device=dev),
            "graph": {},                      # input dtype (False: float32, True: uint8) -> captured CUDA graph
            "ws": _ffi.workspace(_ffi.lib().ssd_decode_nms_workspace_bytes(B, self.base_model.n_anchors,
                                                                            self.base_model.total_labels, T, 0)),
        }
        st.update(self._unpack(st["packed"], B, T))
        dec, lib = self.decoder, _ffi.lib()
        var = _ffi.f32_array(dec.variances)
        priors = dec._priors()
        N, L = self.base_model.n_anchors, self.base_model.total_labels

        def enqueue(u8: bool = False):
            plan.run(u8=u8)
            _ffi.check(lib.ssd_decode_nms(_ffi.ptr(priors), _ffi.ptr(plan.deltas), _ffi.ptr(plan.logits), B, N, L, var, 1,
                                          dec.score_threshold, dec.iou_threshold, T, 0, _ffi.ptr(st["boxes"]),
                                          _ffi.ptr(st["labels"]), _ffi.ptr(st["scores"]), _ffi.ptr(st["valid"]),
                                          _ffi.ptr(st["ws"]), st["ws"].numel(), _ffi.stream()), "ssd_decode_nms")

        st["enqueue"] = enqueue
        self._state[(B, slot)] = st
        return st

    @staticmethod
    def _unpack(packed: torch.Tensor, B: int, T: int) -> Dict[str, torch.Tensor]:
        """The four result tensors as views of one packed buffer (device or pinned host)."""
        n = B * T
        return {"boxes": packed[:4 * n].view(B, T, 4), "labels": packed[4 * n:5 * n].view(B, T),
                "scores": packed[5 * n:6 * n].view(B, T), "valid": packed[6 * n:].view(torch.int32)}

    def _graph(self, st: Dict[str, Any], u8: bool):
        """The captured forward + decode + NMS graph of one slot for one input dtype (captured on first use)."""
        g = st["graph"].get(u8)
        if g is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):                      # warm-up: function attributes, lazy module load, workspaces
                    st["enqueue"](u8)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with _no_gc_during_capture():
                with torch.cuda.graph(g):
                    st["enqueue"](u8)
            st["graph"][u8] = g
        return g

    def launches_per_batch(self, B: int) -> int:
        """libssd_b200 kernels per forward+decode: plan steps + candidate pass + per-image NMS."""
        return self.base_model.plan(B).n_launches + 2

    def run_resident(self, B: int, slot: int = 0, u8: bool = False) -> Dict[str, Any]:
        """Replay on whatever already sits in the slot's (float32 or uint8) image buffer (device-resident timing)."""
        st = self._prepare(B, slot)
        if self.use_cuda_graph:
            self._graph(st, u8).replay()
        else:
            st["enqueue"](u8)
        return st

    def __call__(self, images: Any):
        B = int(images.shape[0])
        st = self._prepare(B)
        u8 = self.base_model._to_image_buffer(st["plan"], images)
        self.run_resident(B, 0, u8)
        self.decoder.last_valid_detections = st["valid"]
        return st["boxes"].clone(), st["labels"].clone(), st["scores"].clone()

    def predict(self, data: Iterable[Any], steps: Optional[int] = None, verbose: int = 0):
        """predictor.py:93-97: three NumPy arrays concatenated over the batches.
        Host batches go through pinned staging buffers; copies run on a second
        stream and overlap the neighbouring batches' compute."""
        dev = _ffi.require_cuda()
        if self._copy_stream is None:
            self._copy_stream = (torch.cuda.Stream(), torch.cuda.Stream())      # H2D, D2H
        (cs, ds), ms = self._copy_stream, torch.cuda.current_stream()
        T = self.decoder.max_total_size
        ob, ol, os_ = [], [], []
        pending: List[Tuple[Dict[str, Any], torch.cuda.Event]] = []
        host: Dict[Tuple[int, int], Dict[str, torch.Tensor]] = {}
        out: Dict[str, Any] = {"rows": 0, "cap": 0}    # preallocated result arrays when ``steps`` is given (no concatenation)

        def drain(entry):
            st, hb, done = entry
            done.synchronize()
            if int(hb["valid"].min()) < 0:               # cannot happen for softmax outputs (at most one class per anchor > 0.5)
                raise _ffi.SsdB200Error("ssd_decode_nms reported a candidate overflow (valid = -1)")
            n = hb["boxes"].shape[0]
            if out["rows"] + n <= out["cap"]:          # known number of batches: results land in the final arrays directly
                r0 = out["rows"]
                out["b"][r0:r0 + n] = hb["boxes"].numpy(); out["l"][r0:r0 + n] = hb["labels"].numpy()
                out["s"][r0:r0 + n] = hb["scores"].numpy()
                out["rows"] = r0 + n
            else:
                res = self._unpack(hb["packed"].clone(), n, T)                         # one host copy, then views
                ob.append(res["boxes"].numpy()); ol.append(res["labels"].numpy()); os_.append(res["scores"].numpy())

        for i, batch in enumerate(data):
            if steps is not None and i >= steps:
                break
            img = batch[0] if isinstance(batch, (tuple, list)) else batch
            B = int(img.shape[0])
            if i == 0 and steps is not None:
                out.update(cap=steps * B, b=np.empty((steps * B, T, 4), np.float32), l=np.empty((steps * B, T), np.float32),
                           s=np.empty((steps * B, T), np.float32))
            slot = i % self.N_SLOTS
            st = self._prepare(B, slot)
            if len(pending) >= self.N_SLOTS:           # the slot's previous user must be fully drained
                drain(pending.pop(0))
            u8 = SSDModel._is_u8(img)
            dst_img = st["plan"].image_u8 if u8 else st["plan"].image
            hb = host.get((B, slot, u8))
            if hb is None:
                hb = {"img": torch.empty(tuple(dst_img.shape), dtype=dst_img.dtype, pin_memory=True),
                      "packed": torch.empty((B * T * 6 + B,), dtype=torch.float32, pin_memory=True)}
                hb.update(self._unpack(hb["packed"], B, T))
                host[(B, slot, u8)] = hb
            if isinstance(img, torch.Tensor) and img.is_cuda:
                with torch.cuda.stream(cs):
                    cs.wait_stream(ms)
                    dst_img.copy_(img if u8 else img.to(torch.float32), non_blocking=True)
            else:
                src = img if isinstance(img, torch.Tensor) else torch.from_numpy(
                    np.ascontiguousarray(img, np.uint8 if u8 else np.float32))
                if tuple(src.shape) != tuple(hb["img"].shape):
                    raise ValueError(f"expected images {tuple(hb['img'].shape)} (NHWC float32 or uint8), got {tuple(src.shape)}")
                if src.dtype != dst_img.dtype:
                    src = src.to(dst_img.dtype)
                if not src.is_pinned():
                    hb["img"].copy_(src)
                    src = hb["img"]
                with torch.cuda.stream(cs):
                    dst_img.copy_(src, non_blocking=True)
            ms.wait_stream(cs)
            self.run_resident(B, slot, u8)
            ds.wait_stream(ms)
            with torch.cuda.stream(ds):
                hb["packed"].copy_(st["packed"], non_blocking=True)        # boxes | labels | scores | valid in ONE copy
                done = torch.cuda.Event()
                done.record(ds)
            pending.append((st, hb, done))
        while pending:
            drain(pending.pop(0))
        if out["rows"]:
            head = (out["b"][:out["rows"]], out["l"][:out["rows"]], out["s"][:out["rows"]])
            if not ob:
                return head
            return (np.concatenate([head[0]] + ob, 0), np.concatenate([head[1]] + ol, 0), np.concatenate([head[2]] + os_, 0))
        if not ob:
            z = np.zeros((0, T), np.float32)
            return np.zeros((0, T, 4), np.float32), z, z.copy()
        return np.concatenate(ob, 0), np.concatenate(ol, 0), np.concatenate(os_, 0)


def smoke_forward(check: Optional[Callable[..., None]] = None) -> None:
    """Tiny forward of the MobileNetV2 graph (used by ``__graft_entry__.smoke``); ``check(model, hyper_params, images,
    deltas, logits)`` lets the caller compare against its oracle (the product never imports ``oracle/``)."""
    from tf_ssd_b200.utils import train_utils
    for backbone in ("mobilenet_v2",):
        hp = train_utils.get_hyper_params(backbone)
        hp["total_labels"] = 21
        model = SSDModel(backbone, hp, seed=1)
        x = np.random.default_rng(0).random((1, hp["img_size"], hp["img_size"], 3), dtype=np.float32)
        d, p = model(x)
        torch.cuda.synchronize()
        assert d.shape == (1, model.n_anchors, 4) and p.shape == (1, model.n_anchors, 21)
        assert bool(torch.isfinite(d).all()) and bool(torch.isfinite(p).all())
        if check is not None:
            dd, zz = model.forward_logits(x)
            torch.cuda.synchronize()
            check(model, hp, x, dd.cpu().numpy(), zz.cpu().numpy())
