"""torch-CPU restatement of the reference's two SSD graphs (forward only).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The float32 forward is
pinned (1e-4 of the output range, variable names and shapes exact) by fixtures
that the reference's own ``models/ssd_vgg16.py``, ``models/ssd_mobilenet_v2.py``
and ``models/header.py`` produced on the Keras stand-in of ``tests/tf_shim``
(``tests/golden/ref_net.npz``).  Still ``[TF-recall]``: the MobileNetV2 backbone
constructor (keras-applications 1.0.8, ``environment.yml:32`` -- third-party, not
under /root/reference) and BatchNormalization's epsilon.

Weights are a flat dict keyed ``"<keras layer name>/<variable>"`` holding
float32 NumPy arrays in Keras layouts: ``kernel`` HWIO, ``depthwise_kernel``
``[3,3,C,1]``, ``bias``, ``gamma``, ``beta``, ``moving_mean``,
``moving_variance``, and ``scale`` for the L2Normalization layer.

Two arithmetic modes:

* ``mode="fp32"``   -- what the reference computes: fp32 everywhere, BatchNorm
  applied as its own op after the convolution.
* ``mode="fp16sim"`` -- the same graph with the storage roundings of the
  B200 path made explicit: BatchNorm folded into the preceding kernel/bias in
  fp32, folded kernels rounded to fp16, fp32 accumulation, every activation
  tensor rounded to fp16 when it is written.  Head outputs stay fp32.  Used to
  separate "fp16 storage" error from kernel bugs in the parity tests.
"""

from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3          # [TF-recall] keras_applications.mobilenet_v2 BatchNormalization(epsilon=1e-3)

# (expansion t, out channels c, stride s) per inverted-residual block id 0..16
# [TF-recall] keras_applications/mobilenet_v2.py, alpha = 1.0
MNV2_BLOCKS: List[Tuple[int, int, int]] = [
    (1, 16, 1),
    (6, 24, 2), (6, 24, 1),
    (6, 32, 2), (6, 32, 1), (6, 32, 1),
    (6, 64, 2), (6, 64, 1), (6, 64, 1), (6, 64, 1),
    (6, 96, 1), (6, 96, 1), (6, 96, 1),
    (6, 160, 2), (6, 160, 1), (6, 160, 1),
    (6, 320, 1),
]


def _q16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float16).to(torch.float32)


def same_pad(size: int, k: int, s: int, d: int = 1) -> Tuple[int, int]:
    """[TF-recall] TensorFlow SAME: the odd pixel goes AFTER."""
    out = -(-size // s)
    total = max((out - 1) * s + (k - 1) * d + 1 - size, 0)
    return total // 2, total - total // 2


def correct_pad(size: int, k: int = 3) -> Tuple[int, int]:
    """[TF-recall] keras_applications ``correct_pad`` used before stride-2 VALID convs."""
    adjust = 1 - size % 2
    c = k // 2
    return c - adjust, c


class _Net:
    def __init__(self, weights: Dict[str, np.ndarray], mode: str):
        assert mode in ("fp32", "fp16sim")
        self.w = {k: torch.from_numpy(np.ascontiguousarray(v)).float() for k, v in weights.items()}
        self.sim = mode == "fp16sim"

    # x is NCHW float32 throughout
    def act(self, x: torch.Tensor, kind: str) -> torch.Tensor:
        if kind == "relu6":
            x = torch.clamp(x, 0.0, 6.0)
        elif kind == "relu":
            x = torch.relu(x)
        return _q16(x) if self.sim else x

    def _bn_params(self, name: str):
        g, b = self.w[name + "/gamma"], self.w[name + "/beta"]
        m, v = self.w[name + "/moving_mean"], self.w[name + "/moving_variance"]
        scale = g / torch.sqrt(v + BN_EPS)
        return scale, b - m * scale

    def conv(self, x, name, stride=1, padding="same", dilation=1, act="none", bn=None, depthwise=False,
             pads=None, keep_fp32=False):
        if depthwise:
            k = self.w[name + "/depthwise_kernel"]                # [kh,kw,C,1]
            wt = k.permute(2, 3, 0, 1).contiguous()               # [C,1,kh,kw]
            groups = wt.shape[0]
        else:
            k = self.w[name + "/kernel"]                          # HWIO
            wt = k.permute(3, 2, 0, 1).contiguous()               # OIHW
            groups = 1
        kh = wt.shape[2]
        bias = self.w.get(name + "/bias")
        if pads is None:
            if padding == "same":
                ph = same_pad(x.shape[2], kh, stride, dilation)
                pw = same_pad(x.shape[3], kh, stride, dilation)
            else:
                ph = pw = (0, 0)
        else:
            ph, pw = pads
        x = F.pad(x, (pw[0], pw[1], ph[0], ph[1]))
        if self.sim:
            b_eff = bias if bias is not None else torch.zeros(wt.shape[0])
            if bn is not None:
                scale, shift = self._bn_params(bn)
                wt = wt * scale.view(-1, 1, 1, 1)
                b_eff = b_eff * scale + shift
            y = F.conv2d(x, _q16(wt), b_eff, stride=stride, dilation=dilation, groups=groups)
        else:
            y = F.conv2d(x, wt, bias, stride=stride, dilation=dilation, groups=groups)
            if bn is not None:
                scale, shift = self._bn_params(bn)
                y = y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
        if keep_fp32:
            return y
        return self.act(y, act)

    def maxpool(self, x, k, s):
        """Keras MaxPool2D(padding='same'): pad with -inf, odd pixel after."""
        ph = same_pad(x.shape[2], k, s)
        pw = same_pad(x.shape[3], k, s)
        x = F.pad(x, (pw[0], pw[1], ph[0], ph[1]), value=float("-inf"))
        return F.max_pool2d(x, k, s)

    def head(self, hyper_params, taps: Sequence[torch.Tensor], return_logits: bool):
        """models/header.py:54-90 -- per-map 3x3 SAME conv pair, reshape+concat
        (``HeadWrapper.call`` :46-51), softmax over L (:88).  Returns
        ``(pred_deltas, pred_labels)`` in that order (:90)."""
        L = hyper_params["total_labels"]
        labels, boxes = [], []
        for i, t in enumerate(taps):
            lab = self.conv(t, f"{i + 1}_conv_label_output", keep_fp32=True)
            box = self.conv(t, f"{i + 1}_conv_boxes_output", keep_fp32=True)
            B = t.shape[0]
            labels.append(lab.permute(0, 2, 3, 1).reshape(B, -1, L))
            boxes.append(box.permute(0, 2, 3, 1).reshape(B, -1, 4))
        logits = torch.cat(labels, dim=1)
        deltas = torch.cat(boxes, dim=1)
        if return_logits:
            return deltas.numpy(), logits.numpy()
        return deltas.numpy(), torch.softmax(logits, dim=-1).numpy()


def _to_nchw(images: np.ndarray, sim: bool) -> torch.Tensor:
    x = torch.from_numpy(np.ascontiguousarray(images, dtype=np.float32)).permute(0, 3, 1, 2).contiguous()
    return _q16(x) if sim else x


def mobilenet_v2_ssd_forward(weights, hyper_params, images: np.ndarray, mode: str = "fp32",
                             return_logits: bool = False, return_taps: bool = False):
    """models/ssd_mobilenet_v2.py:15-47 over [TF-recall] keras_applications MobileNetV2
    (alpha 1.0, include_top=False).  ``images`` is NHWC float32."""
    n = _Net(weights, mode)
    with torch.no_grad():
        x = _to_nchw(images, n.sim)
        p = correct_pad(x.shape[2])
        x = n.conv(x, "Conv1", stride=2, pads=(p, correct_pad(x.shape[3])), act="relu6", bn="bn_Conv1")
        taps = []
        for bid, (t, c, s) in enumerate(MNV2_BLOCKS):
            prefix = f"block_{bid}_" if bid else "expanded_conv_"
            inp = x
            cin = x.shape[1]
            if bid:
                x = n.conv(x, prefix + "expand", act="relu6", bn=prefix + "expand_BN")
                if bid == 13:
                    taps.append(x)                                  # block_13_expand_relu  :27
            if s == 2:
                pads = (correct_pad(x.shape[2]), correct_pad(x.shape[3]))
                x = n.conv(x, prefix + "depthwise", stride=2, pads=pads, act="relu6",
                           bn=prefix + "depthwise_BN", depthwise=True)
            else:
                x = n.conv(x, prefix + "depthwise", act="relu6", bn=prefix + "depthwise_BN", depthwise=True)
            x = n.conv(x, prefix + "project", act="none", bn=prefix + "project_BN", keep_fp32=True)
            if s == 1 and cin == c:
                x = x + inp                                         # block_i_add
            x = n.act(x, "none")
        x = n.conv(x, "Conv_1", act="relu6", bn="Conv_1_bn")
        taps.append(x)                                              # out_relu  :28
        for i in range(1, 5):                                       # extras :31-41
            x = n.conv(x, f"extra{i}_1", padding="valid", act="relu")
            x = n.conv(x, f"extra{i}_2", stride=2, padding="same", act="relu")
            taps.append(x)
        out = n.head(hyper_params, taps, return_logits)
        if return_taps:
            return out, [t.permute(0, 2, 3, 1).numpy() for t in taps]
        return out


def vgg16_ssd_forward(weights, hyper_params, images: np.ndarray, mode: str = "fp32",
                      return_logits: bool = False, return_taps: bool = False):
    """models/ssd_vgg16.py:66-121 (+ L2Normalization :15-63).  With
    ``hyper_params["img_size"] == 512`` and 7 feature maps the SSD512 extension
    of SURVEY.md Appendix C is built instead (stride-2 SAME conv10_2/conv11_2
    and an extra conv12 block) -- that variant has no reference graph."""
    n = _Net(weights, mode)
    ssd512 = len(hyper_params["feature_map_shapes"]) == 7
    with torch.no_grad():
        x = _to_nchw(images, n.sim)
        cfg = [("conv1", 2), ("conv2", 2), ("conv3", 3), ("conv4", 3), ("conv5", 3)]
        conv4_3 = None
        for bname, reps in cfg:
            for r in range(1, reps + 1):
                x = n.conv(x, f"{bname}_{r}", act="relu")
            if bname == "conv4":
                conv4_3 = x
            x = n.maxpool(x, 2, 2) if bname != "conv5" else n.maxpool(x, 3, 1)   # :82-101
        x = n.conv(x, "conv6", dilation=6, act="relu")              # :103
        conv7 = n.conv(x, "conv7", act="relu")                      # :104
        x = n.conv(conv7, "conv8_1", padding="valid", act="relu")
        conv8_2 = n.conv(x, "conv8_2", stride=2, act="relu")
        x = n.conv(conv8_2, "conv9_1", padding="valid", act="relu")
        conv9_2 = n.conv(x, "conv9_2", stride=2, act="relu")
        x = n.conv(conv9_2, "conv10_1", padding="valid", act="relu")
        if ssd512:
            conv10_2 = n.conv(x, "conv10_2", stride=2, act="relu")
            x = n.conv(conv10_2, "conv11_1", padding="valid", act="relu")
            conv11_2 = n.conv(x, "conv11_2", stride=2, act="relu")
            x = n.conv(conv11_2, "conv12_1", padding="valid", act="relu")
            conv12_2 = n.conv(x, "conv12_2", stride=2, act="relu")
        else:
            conv10_2 = n.conv(x, "conv10_2", padding="valid", act="relu")    # :111
            x = n.conv(conv10_2, "conv11_1", padding="valid", act="relu")
            conv11_2 = n.conv(x, "conv11_2", padding="valid", act="relu")    # :113
        # L2Normalization.call :63 -- [TF-recall] l2_normalize = x * rsqrt(max(sum x^2, 1e-12))
        ss = torch.sum(conv4_3 * conv4_3, dim=1, keepdim=True)
        norm = conv4_3 * torch.rsqrt(torch.clamp(ss, min=1e-12)) * n.w["l2_normalization/scale"].view(1, -1, 1, 1)
        norm = n.act(norm, "none")
        taps = [norm, conv7, conv8_2, conv9_2, conv10_2, conv11_2] + ([conv12_2] if ssd512 else [])
        out = n.head(hyper_params, taps, return_logits)
        if return_taps:
            return out, [t.permute(0, 2, 3, 1).numpy() for t in taps]
        return out


def forward(backbone: str, weights, hyper_params, images, **kw):
    if backbone == "mobilenet_v2":
        return mobilenet_v2_ssd_forward(weights, hyper_params, images, **kw)
    return vgg16_ssd_forward(weights, hyper_params, images, **kw)
