"""torch-CPU autograd restatement of one Keras training step of the reference
(trainer.py:86-127): forward (models/ssd_vgg16.py, models/ssd_mobilenet_v2.py in training mode), the two CustomLoss terms
(ssd_loss.py:26-91) reduced like Keras ``compile(loss=[...])`` does (mean over the batch, summed),
the l2(5e-4) kernel regulariser (ssd_vgg16.py:76) and Adam (trainer.py:92, Keras defaults).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED against TensorFlow:
the reduction rule, the regulariser placement and Adam's epsilon handling are ``[TF-recall]``.
The hard-negative mask is taken from ``box_oracle`` (it is not differentiable); everything else is
differentiated by torch autograd in float32.
"""

from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from oracle import box_oracle as bo
from oracle.net_oracle import BN_EPS, MNV2_BLOCKS, correct_pad, same_pad

L2_REG = 5e-4


def _conv(x, w, name, stride=1, padding="same", dilation=1, relu=True):
    k = w[name + "/kernel"]                                   # HWIO
    wt = k.permute(3, 2, 0, 1)
    kh = wt.shape[2]
    if padding == "same":
        ph, pw = same_pad(x.shape[2], kh, stride, dilation), same_pad(x.shape[3], kh, stride, dilation)
        x = F.pad(x, (pw[0], pw[1], ph[0], ph[1]))
    y = F.conv2d(x, wt, w[name + "/bias"], stride=stride, dilation=dilation)
    return torch.relu(y) if relu else y


def _pool(x, k, s):
    ph, pw = same_pad(x.shape[2], k, s), same_pad(x.shape[3], k, s)
    return F.max_pool2d(F.pad(x, (pw[0], pw[1], ph[0], ph[1]), value=float("-inf")), k, s)


def vgg16_forward_torch(w: Dict[str, torch.Tensor], hyper_params, images: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """models/ssd_vgg16.py:78-121, differentiable; returns (pred_deltas, logits)."""
    x = images.permute(0, 3, 1, 2)
    conv4_3 = None
    for bname, reps in [("conv1", 2), ("conv2", 2), ("conv3", 3), ("conv4", 3), ("conv5", 3)]:
        for r in range(1, reps + 1):
            x = _conv(x, w, f"{bname}_{r}")
        if bname == "conv4":
            conv4_3 = x
        x = _pool(x, 2, 2) if bname != "conv5" else _pool(x, 3, 1)
    x = _conv(x, w, "conv6", dilation=6)
    conv7 = _conv(x, w, "conv7")
    x = _conv(conv7, w, "conv8_1", padding="valid")
    conv8_2 = _conv(x, w, "conv8_2", stride=2)
    x = _conv(conv8_2, w, "conv9_1", padding="valid")
    conv9_2 = _conv(x, w, "conv9_2", stride=2)
    x = _conv(conv9_2, w, "conv10_1", padding="valid")
    conv10_2 = _conv(x, w, "conv10_2", padding="valid")
    x = _conv(conv10_2, w, "conv11_1", padding="valid")
    conv11_2 = _conv(x, w, "conv11_2", padding="valid")
    ss = torch.sum(conv4_3 * conv4_3, dim=1, keepdim=True)
    norm = conv4_3 * torch.rsqrt(torch.clamp(ss, min=1e-12)) * w["l2_normalization/scale"].view(1, -1, 1, 1)
    L = hyper_params["total_labels"]
    labels, boxes = [], []
    for i, t in enumerate([norm, conv7, conv8_2, conv9_2, conv10_2, conv11_2]):
        lab = _conv(t, w, f"{i + 1}_conv_label_output", relu=False)
        box = _conv(t, w, f"{i + 1}_conv_boxes_output", relu=False)
        B = t.shape[0]
        labels.append(lab.permute(0, 2, 3, 1).reshape(B, -1, L))
        boxes.append(box.permute(0, 2, 3, 1).reshape(B, -1, 4))
    return torch.cat(boxes, 1), torch.cat(labels, 1)


def _q16(x, sim):
    """``sim``: round to fp16 like the device's activation storage, with a straight-through gradient, so that
    activation masks (ReLU6 knees) are evaluated on the values the device actually stored."""
    return x + (x.to(torch.float16).to(torch.float32) - x).detach() if sim else x


def _bn_train(y, w, name, stats=None):
    """[TF-recall] keras BatchNormalization(epsilon=1e-3) with training=True: batch mean and BIASED variance
    over (B, H, W); ``stats`` (optional dict) receives them for the moving-average check."""
    mean = y.mean(dim=(0, 2, 3), keepdim=True)
    var = ((y - mean) ** 2).mean(dim=(0, 2, 3), keepdim=True)
    if stats is not None:
        stats[name] = (mean.detach().reshape(-1).numpy(), var.detach().reshape(-1).numpy(), y.shape[0] * y.shape[2] * y.shape[3])
    return (y - mean) * torch.rsqrt(var + BN_EPS) * w[name + "/gamma"].view(1, -1, 1, 1) + w[name + "/beta"].view(1, -1, 1, 1)


def _conv_bn(x, w, name, bn, stride=1, pads=None, depthwise=False, relu6=True, stats=None, sim=False, res=None):
    if depthwise:
        wt = w[name + "/depthwise_kernel"].permute(2, 3, 0, 1)          # [C,1,3,3]
        groups = wt.shape[0]
    else:
        wt = w[name + "/kernel"].permute(3, 2, 0, 1)
        groups = 1
    kh = wt.shape[2]
    if pads is None:
        pads = (same_pad(x.shape[2], kh, stride), same_pad(x.shape[3], kh, stride))
    x = F.pad(x, (pads[1][0], pads[1][1], pads[0][0], pads[0][1]))
    y = _bn_train(_q16(F.conv2d(x, wt, None, stride=stride, groups=groups), sim), w, bn, stats)
    y = torch.clamp(y, 0.0, 6.0) if relu6 else y
    if res is not None:
        y = y + res
    return _q16(y, sim)


def mobilenet_v2_forward_torch(w: Dict[str, torch.Tensor], hyper_params, images: torch.Tensor, stats=None, sim=False):
    """models/ssd_mobilenet_v2.py:24-46 over [TF-recall] keras_applications MobileNetV2, TRAINING mode
    (BatchNorm on batch statistics), differentiable; returns (pred_deltas, logits).  ``sim=True`` additionally
    rounds every stored activation to fp16 (straight-through), mirroring the device's storage format."""
    x = _q16(images.permute(0, 3, 1, 2), sim)
    x = _conv_bn(x, w, "Conv1", "bn_Conv1", stride=2, pads=(correct_pad(x.shape[2]), correct_pad(x.shape[3])), stats=stats, sim=sim)
    taps = []
    for bid, (t, c, s) in enumerate(MNV2_BLOCKS):
        prefix = f"block_{bid}_" if bid else "expanded_conv_"
        inp = x
        if bid:
            x = _conv_bn(x, w, prefix + "expand", prefix + "expand_BN", stats=stats, sim=sim)
            if bid == 13:
                taps.append(x)
        pads = (correct_pad(x.shape[2]), correct_pad(x.shape[3])) if s == 2 else None
        x = _conv_bn(x, w, prefix + "depthwise", prefix + "depthwise_BN", stride=s, pads=pads, depthwise=True, stats=stats, sim=sim)
        x = _conv_bn(x, w, prefix + "project", prefix + "project_BN", relu6=False, stats=stats, sim=sim,
                     res=inp if (s == 1 and inp.shape[1] == c) else None)
    x = _conv_bn(x, w, "Conv_1", "Conv_1_bn", stats=stats, sim=sim)
    taps.append(x)
    for i in range(1, 5):
        x = _q16(_conv(x, w, f"extra{i}_1", padding="valid"), sim)
        x = _q16(_conv(x, w, f"extra{i}_2", stride=2), sim)
        taps.append(x)
    L = hyper_params["total_labels"]
    labels, boxes = [], []
    for i, t in enumerate(taps):
        lab = _conv(t, w, f"{i + 1}_conv_label_output", relu=False)
        box = _conv(t, w, f"{i + 1}_conv_boxes_output", relu=False)
        B = t.shape[0]
        labels.append(lab.permute(0, 2, 3, 1).reshape(B, -1, L))
        boxes.append(box.permute(0, 2, 3, 1).reshape(B, -1, 4))
    return torch.cat(boxes, 1), torch.cat(labels, 1)


def losses_torch(actual_deltas, actual_labels, pred_deltas, logits, neg_pos_ratio=3.0, alpha=1.0):
    """ssd_loss.py:26-91 on tensors (from logits); per-image ``(loc [B], conf [B])``."""
    ad, al = torch.as_tensor(actual_deltas), torch.as_tensor(actual_labels)
    e = (pred_deltas - ad).abs()
    hub = torch.where(e <= 1.0, 0.5 * e * e, e - 0.5).sum(-1)                      # Huber mean * 4 == sum over coords
    pos = (ad != 0).any(-1).float()
    n_pos = pos.sum(1)
    loc = (hub * pos).sum(1) / torch.where(n_pos == 0, torch.ones_like(n_pos), n_pos) * alpha
    ce = -(al * torch.log_softmax(logits, -1)).sum(-1)
    # the hard-negative selection is a constant of the step (argsort ranks, ssd_loss.py:78-84)
    with torch.no_grad():
        masked = (ce * al[..., 0]).numpy()
        posc = (al[..., 1:] != 0).any(-1).numpy().astype(np.float32)
        n_posc = posc.sum(1)
        n_neg = (n_posc * np.float32(neg_pos_ratio)).astype(np.int32)
        neg = np.stack([(bo.hard_negative_rank(masked[b]) < n_neg[b]) for b in range(masked.shape[0])]).astype(np.float32)
        final = torch.from_numpy(posc + neg)
        div = torch.from_numpy(np.where(n_posc == 0, 1.0, n_posc).astype(np.float32))
    conf = (final * ce).sum(1) / div
    return loc, conf


def train_step(weights: Dict[str, np.ndarray], hyper_params, images: np.ndarray, actual_deltas: np.ndarray,
               actual_labels: np.ndarray, l2_kernels: Sequence[str], neg_pos_ratio=3.0, alpha=1.0,
               backbone: str = "vgg16", stats=None, fp16sim: bool = False):
    """Returns ``(loss dict, grads dict)`` -- gradients of mean_B(loc) + mean_B(conf) + reg w.r.t. every variable."""
    w = {k: torch.tensor(v, dtype=torch.float32, requires_grad=not k.split("/")[-1].startswith("moving_"))
         for k, v in weights.items()}
    img = torch.from_numpy(np.ascontiguousarray(images, np.float32))
    if backbone == "mobilenet_v2":
        pd, z = mobilenet_v2_forward_torch(w, hyper_params, img, stats, sim=fp16sim)
    else:
        pd, z = vgg16_forward_torch(w, hyper_params, img)
    loc, conf = losses_torch(actual_deltas, actual_labels, pd, z, neg_pos_ratio, alpha)
    reg = sum((L2_REG * (w[k] ** 2).sum() for k in l2_kernels), torch.zeros(()))
    total = loc.mean() + conf.mean() + reg
    total.backward()
    grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros_like(weights[k])) for k, v in w.items()}
    return dict(loss=float(total.detach()), loc=loc.detach().numpy(), conf=conf.detach().numpy(), reg=float(reg.detach())), grads


def adam_update(w, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
    """[TF-recall] Keras Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); w -= lr_t * m / (sqrt(v) + eps)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    return w - lr_t * m / (np.sqrt(v) + eps), m, v
