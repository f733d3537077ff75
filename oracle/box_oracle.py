"""NumPy restatement of the reference's box math, targets, loss and decode+NMS.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Pinned by fixtures that
the reference's own source produced on the NumPy ``tensorflow`` stand-in
(``tests/golden/ref_*.npz``, ``tests/test_ref_golden.py``: bit-exact for priors,
IoU, targets and NMS selection); rules that come from TensorFlow itself and that
its documentation does not pin are tagged ``[TF-recall]``.

All arithmetic is float32 with one rounding per elementary operation, exactly
as a chain of separate TensorFlow eager ops would produce it (TensorFlow never
contracts ``a*b+c`` across ops), so integer/index results can be compared
bit-for-bit with the CUDA kernels, which are compiled with ``-fmad=false``.
"""

from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------
# priors  (utils/bbox_utils.py:131-214)
# --------------------------------------------------------------------------
def scale_for_feature_map(k: int, m: int = 6, scale_min: float = 0.2, scale_max: float = 0.9) -> float:
    """utils/bbox_utils.py:131-148 -- python float64 arithmetic."""
    return scale_min + ((scale_max - scale_min) / (m - 1)) * (k - 1)


def base_prior_boxes(aspect_ratios: Sequence[float], fm_index: int, total_maps: int) -> np.ndarray:
    """utils/bbox_utils.py:151-176.

    ``tf.sqrt(python_float)`` makes a float32 tensor, so sqrt/div/mul are
    float32; the extra square box takes ``s_k * s_{k+1}`` in float64, rounds
    that product to float32 and then takes a float32 sqrt.  The extra box is
    appended LAST.
    """
    s_cur = scale_for_feature_map(fm_index, m=total_maps)
    s_next = scale_for_feature_map(fm_index + 1, m=total_maps)
    rows = []
    for ar in aspect_ratios:
        root = np.sqrt(F32(ar))
        h = F32(s_cur) / root
        w = F32(s_cur) * root
        rows.append([-h / F32(2), -w / F32(2), h / F32(2), w / F32(2)])
    side = np.sqrt(F32(s_cur * s_next))
    rows.append([-side / F32(2), -side / F32(2), side / F32(2), side / F32(2)])
    return np.asarray(rows, dtype=F32)


def prior_boxes(feature_map_shapes: Sequence[int], aspect_ratios: Sequence[Sequence[float]]) -> np.ndarray:
    """utils/bbox_utils.py:179-214.

    Cell centres are float64 (``int32 range / int`` is float64 in TF) and only
    then rounded to float32 (:198-201).  ``meshgrid`` is xy-indexed so the
    flattened cell order is y-major (:202-205); anchors are cell-major,
    anchor-minor (:207-211); maps are concatenated and clipped to [0,1]
    (:213-214).
    """
    out = []
    total = len(feature_map_shapes)
    for i, fm in enumerate(feature_map_shapes):
        base = base_prior_boxes(aspect_ratios[i], i + 1, total)
        stride = 1 / fm
        centres = (np.arange(fm, dtype=np.int32).astype(np.float64) / fm + stride / 2).astype(F32)
        gx, gy = np.meshgrid(centres, centres)
        fx, fy = gx.reshape(-1), gy.reshape(-1)
        grid = np.stack([fy, fx, fy, fx], axis=-1)
        out.append((base.reshape(1, -1, 4) + grid.reshape(-1, 1, 4)).reshape(-1, 4))
    return np.clip(np.concatenate(out, axis=0), F32(0), F32(1)).astype(F32)


# --------------------------------------------------------------------------
# IoU, encode, decode  (utils/bbox_utils.py:24-128)
# --------------------------------------------------------------------------
def iou_map(bboxes: np.ndarray, gt_boxes: np.ndarray, transpose_perm: Optional[Sequence[int]] = None) -> np.ndarray:
    """utils/bbox_utils.py:24-55.  0/0 -> NaN, exactly like the reference.

    Shape modes: ``[N,4] x [B,G,4] -> [B,N,G]``, ``[B,M,4] x [B,G,4] ->
    [B,M,G]`` and, with ``transpose_perm=[1,0]``, ``[N,4] x [G,4] -> [N,G]``.
    """
    perm = list(transpose_perm) if transpose_perm else [0, 2, 1]
    bboxes = np.asarray(bboxes, dtype=F32)
    gt_boxes = np.asarray(gt_boxes, dtype=F32)
    expand_axis = gt_boxes.ndim - 2
    by1, bx1, by2, bx2 = np.split(bboxes, 4, axis=-1)
    gy1, gx1, gy2, gx2 = np.split(gt_boxes, 4, axis=-1)
    g_area = np.squeeze((gy2 - gy1) * (gx2 - gx1), axis=-1)
    b_area = np.squeeze((by2 - by1) * (bx2 - bx1), axis=-1)
    x_top = np.maximum(bx1, np.transpose(gx1, perm))
    y_top = np.maximum(by1, np.transpose(gy1, perm))
    x_bot = np.minimum(bx2, np.transpose(gx2, perm))
    y_bot = np.minimum(by2, np.transpose(gy2, perm))
    inter = np.maximum(x_bot - x_top, F32(0)) * np.maximum(y_bot - y_top, F32(0))
    union = np.expand_dims(b_area, -1) + np.expand_dims(g_area, expand_axis) - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / union).astype(F32)


def boxes_from_deltas(priors: np.ndarray, deltas: np.ndarray) -> np.ndarray:
    """utils/bbox_utils.py:58-82.  Note ``y2 = h + y1`` (not ``cy + h/2``)."""
    priors = np.asarray(priors, dtype=F32)
    deltas = np.asarray(deltas, dtype=F32)
    pw = priors[..., 3] - priors[..., 1]
    ph = priors[..., 2] - priors[..., 0]
    pcx = priors[..., 1] + F32(0.5) * pw
    pcy = priors[..., 0] + F32(0.5) * ph
    w = np.exp(deltas[..., 3]) * pw
    h = np.exp(deltas[..., 2]) * ph
    cx = deltas[..., 1] * pw + pcx
    cy = deltas[..., 0] * ph + pcy
    y1 = cy - F32(0.5) * h
    x1 = cx - F32(0.5) * w
    return np.stack([y1, x1, h + y1, w + x1], axis=-1).astype(F32)


def deltas_from_boxes(priors: np.ndarray, gt: np.ndarray) -> np.ndarray:
    """utils/bbox_utils.py:85-128.  Zero prior extent -> 1e-3; zero GT extent -> delta 0."""
    priors = np.asarray(priors, dtype=F32)
    gt = np.asarray(gt, dtype=F32)
    pw = priors[..., 3] - priors[..., 1]
    ph = priors[..., 2] - priors[..., 0]
    pcx = priors[..., 1] + F32(0.5) * pw
    pcy = priors[..., 0] + F32(0.5) * ph
    gw = gt[..., 3] - gt[..., 1]
    gh = gt[..., 2] - gt[..., 0]
    gcx = gt[..., 1] + F32(0.5) * gw
    gcy = gt[..., 0] + F32(0.5) * gh
    pw = np.where(pw == 0, F32(1e-3), pw)
    ph = np.where(ph == 0, F32(1e-3), ph)
    zero = np.zeros_like(gw)
    with np.errstate(divide="ignore", invalid="ignore"):
        dx = np.where(gw == 0, zero, (gcx - pcx) / pw)
        dy = np.where(gh == 0, zero, (gcy - pcy) / ph)
        dw = np.where(gw == 0, zero, np.log(gw / pw))
        dh = np.where(gh == 0, zero, np.log(gh / ph))
    return np.stack([dy, dx, dh, dw], axis=-1).astype(F32)


# --------------------------------------------------------------------------
# target assignment  (utils/train_utils.py:102-136)
# --------------------------------------------------------------------------
def match_encode(priors, gt_boxes, gt_labels, total_labels: int, iou_threshold: float = 0.5,
                 variances=(0.1, 0.1, 0.2, 0.2), return_aux: bool = False):
    """utils/train_utils.py:102-136.

    argmax over G takes the FIRST maximum [TF-recall: tf.argmax]; the positive
    test is a strict ``>``; there is no bipartite best-prior-per-GT step.
    ``tf.one_hot`` of an out-of-range label (the -1 padding) is an all-zero row.
    """
    gt_boxes = np.asarray(gt_boxes, dtype=F32)
    gt_labels = np.asarray(gt_labels, dtype=np.int32)
    iou = iou_map(priors, gt_boxes)                                     # :123
    idx = np.argmax(iou, axis=2).astype(np.int32)                       # :124
    best = np.max(iou, axis=2)                                          # :125
    pos = best > F32(iou_threshold)                                     # :126
    picked = np.take_along_axis(gt_boxes, idx[..., None], axis=1)       # :129
    picked = np.where(pos[..., None], picked, F32(0))                   # :130
    deltas = deltas_from_boxes(priors, picked) / np.asarray(variances, dtype=F32)   # :131
    lab = np.take_along_axis(gt_labels, idx, axis=1)                    # :133
    lab = np.where(pos, lab, 0).astype(np.int32)                        # :134
    onehot = (lab[..., None] == np.arange(total_labels, dtype=np.int32)).astype(F32)  # :135
    if return_aux:
        return deltas.astype(F32), onehot, idx, best, lab
    return deltas.astype(F32), onehot


# --------------------------------------------------------------------------
# loss  (ssd_loss.py:26-91)
# --------------------------------------------------------------------------
def _huber_sum4(actual: np.ndarray, pred: np.ndarray) -> np.ndarray:
    """Keras Huber(delta=1, reduction=NONE) as normalised by ssd_loss.py:36-43.

    [TF-recall] TF 2.0 returns the element-wise ``[B,N,4]`` tensor (the
    reference then reduce_sums it, :41); TF >= 2.1 returns the mean over the
    last axis and the reference multiplies by 4 (:42).  Both equal the sum of
    the four Huber terms (x/4*4 is exact in binary floating point).
    """
    err = pred - actual
    a = np.abs(err)
    quad = np.minimum(a, F32(1))
    lin = a - quad
    return np.sum(F32(0.5) * (quad * quad) + F32(1) * lin, axis=-1, dtype=F32)


def loc_loss(actual_deltas, pred_deltas, loc_loss_alpha: float = 1.0) -> np.ndarray:
    """ssd_loss.py:26-57 -> per-image ``[B]``."""
    actual = np.asarray(actual_deltas, dtype=F32)
    pred = np.asarray(pred_deltas, dtype=F32)
    per_anchor = _huber_sum4(actual, pred)
    pos = np.any(actual != 0, axis=2).astype(F32)                       # :46-47
    n_pos = np.sum(pos, axis=1, dtype=F32)                              # :48
    total = np.sum(pos * per_anchor, axis=-1, dtype=F32)                # :50
    n_pos = np.where(n_pos == 0, F32(1), n_pos)                         # :51-55
    return (total / n_pos * F32(loc_loss_alpha)).astype(F32)


def categorical_ce_from_probs(y: np.ndarray, p: np.ndarray) -> np.ndarray:
    """[TF-recall] Keras ``categorical_crossentropy`` probability path:
    renormalise, clip to [1e-7, 1-1e-7], ``-sum(y*log p)``  (ssd_loss.py:69-70)."""
    p = np.asarray(p, dtype=F32)
    p = p / np.sum(p, axis=-1, keepdims=True, dtype=F32)
    p = np.clip(p, F32(1e-7), F32(1) - F32(1e-7))
    return (-np.sum(np.asarray(y, dtype=F32) * np.log(p), axis=-1, dtype=F32)).astype(F32)


def categorical_ce_from_logits(y: np.ndarray, z: np.ndarray) -> np.ndarray:
    """[TF-recall] ``softmax_cross_entropy_with_logits`` -- what Keras graph
    mode substitutes when the prediction is the direct output of a Softmax op
    (the ``fit`` path of trainer.py:120-127)."""
    z = np.asarray(z, dtype=F32)
    zmax = np.max(z, axis=-1, keepdims=True)
    lse = np.log(np.sum(np.exp(z - zmax), axis=-1, keepdims=True, dtype=F32)) + zmax
    return (-np.sum(np.asarray(y, dtype=F32) * (z - lse), axis=-1, dtype=F32)).astype(F32)


def hard_negative_rank(masked_loss: np.ndarray) -> np.ndarray:
    """ssd_loss.py:79-80: ``argsort(argsort(x, DESCENDING))`` along the last
    axis.  [TF-recall] descending argsort goes through top_k, which puts the
    lower index first among equal values."""
    order = np.argsort(-masked_loss.astype(np.float64), axis=-1, kind="stable")
    return np.argsort(order, axis=-1, kind="stable").astype(np.int32)


def conf_loss(actual_labels, pred_labels, neg_pos_ratio: float = 3.0, from_logits: bool = False,
              return_aux: bool = False):
    """ssd_loss.py:59-91 -> per-image ``[B]``.

    Ranks run over ALL anchors (positives carry masked loss 0), so when
    ``3*n_pos`` exceeds the number of non-zero-loss negatives the zero-loss
    anchors -- positives included, lowest index first -- are also selected and
    ``final_mask`` can reach 2 on a positive (:78-85).
    """
    y = np.asarray(actual_labels, dtype=F32)
    ce = categorical_ce_from_logits(y, pred_labels) if from_logits else categorical_ce_from_probs(y, pred_labels)
    pos = np.any(y[..., 1:] != 0, axis=2).astype(F32)                   # :72-73
    n_pos = np.sum(pos, axis=1, dtype=F32)                              # :74
    n_neg = (n_pos * F32(neg_pos_ratio)).astype(np.int32)               # :75
    masked = ce * y[..., 0]                                             # :78
    rank = hard_negative_rank(masked)                                   # :79-80
    neg = (rank < n_neg[:, None]).astype(F32)                           # :81-82
    final = pos + neg                                                   # :84
    total = np.sum(final * ce, axis=-1, dtype=F32)                      # :85
    n_div = np.where(n_pos == 0, F32(1), n_pos)                         # :86-90
    out = (total / n_div).astype(F32)
    if return_aux:
        return out, ce, final, rank
    return out


def loss_grads(actual_deltas, pred_deltas, actual_labels, pred_logits, neg_pos_ratio=3.0, loc_loss_alpha=1.0):
    """Gradient of ``mean_B(loc) + mean_B(conf)`` (Keras SUM_OVER_BATCH_SIZE of
    each per-image loss, trainer.py:91-94 [TF-recall]) w.r.t. pred_deltas and the
    pre-softmax logits, with the mining mask treated as a constant (argsort
    has no gradient).  float64 internally; used to check the CUDA backward."""
    a = np.asarray(actual_deltas, dtype=np.float64)
    p = np.asarray(pred_deltas, dtype=np.float64)
    y = np.asarray(actual_labels, dtype=np.float64)
    z = np.asarray(pred_logits, dtype=np.float64)
    B = a.shape[0]
    pos = np.any(a != 0, axis=2).astype(np.float64)
    n_pos = np.maximum(pos.sum(axis=1), 1.0)
    g_d = np.clip(p - a, -1.0, 1.0) * pos[..., None] * (loc_loss_alpha / n_pos)[:, None, None] / B
    _, _, final, _ = conf_loss(actual_labels, pred_logits, neg_pos_ratio, from_logits=True, return_aux=True)
    cpos = np.any(y[..., 1:] != 0, axis=2).astype(np.float64)
    c_npos = np.maximum(cpos.sum(axis=1), 1.0)
    sm = np.exp(z - z.max(axis=-1, keepdims=True))
    sm /= sm.sum(axis=-1, keepdims=True)
    g_z = (sm * y.sum(axis=-1, keepdims=True) - y) * final.astype(np.float64)[..., None] / c_npos[:, None, None] / B
    return g_d, g_z


# --------------------------------------------------------------------------
# decode + combined NMS  (models/decoder.py:60-93, utils/bbox_utils.py:10-21)
# --------------------------------------------------------------------------
def softmax(z: np.ndarray) -> np.ndarray:
    """models/header.py:88 -- Keras softmax over the last axis, float32."""
    z = np.asarray(z, dtype=F32)
    e = np.exp(z - np.max(z, axis=-1, keepdims=True))
    return (e / np.sum(e, axis=-1, keepdims=True, dtype=F32)).astype(F32)


def _nms_iou(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """[TF-recall] IoU of tensorflow/core/kernels/non_max_suppression_op.cc:
    corners canonicalised with min/max, 0 when either area <= 0.
    ``a`` is one box ``[4]``, ``b`` is ``[K,4]``."""
    aymin, aymax = np.minimum(a[0], a[2]), np.maximum(a[0], a[2])
    axmin, axmax = np.minimum(a[1], a[3]), np.maximum(a[1], a[3])
    bymin, bymax = np.minimum(b[:, 0], b[:, 2]), np.maximum(b[:, 0], b[:, 2])
    bxmin, bxmax = np.minimum(b[:, 1], b[:, 3]), np.maximum(b[:, 1], b[:, 3])
    area_a = (aymax - aymin) * (axmax - axmin)
    area_b = (bymax - bymin) * (bxmax - bxmin)
    ih = np.maximum(np.minimum(aymax, bymax) - np.maximum(aymin, bymin), F32(0))
    iw = np.maximum(np.minimum(axmax, bxmax) - np.maximum(axmin, bxmin), F32(0))
    inter = ih * iw
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = inter / (area_a + area_b - inter)
    return np.where((area_a <= 0) | (area_b <= 0), F32(0), iou).astype(F32)


def combined_nms(boxes, scores, max_output_size_per_class: int, max_total_size: int,
                 iou_threshold: float = 0.5, score_threshold: float = float("-inf"),
                 clip_boxes: bool = True, return_indices: bool = False):
    """[TF-recall] ``tf.image.combined_non_max_suppression`` with
    ``pad_per_class=False`` (the call at utils/bbox_utils.py:21 <- decoder.py:86-92).

    ``boxes`` is ``[B,N,q,4]`` with q in {1, L}; ``scores`` is ``[B,N,L]``.
    Per class: candidates are ``score > score_threshold`` (strict), visited in
    descending score, kept iff IoU with every kept box of that class is
    ``<= iou_threshold``; at most ``max_output_size_per_class`` kept.  All
    kept boxes of an image are then ordered by score and the first
    ``max_total_size`` are returned, zero padded, boxes clipped to [0,1].

    Equal scores: TensorFlow's order is implementation defined
    (std::priority_queue / std::sort).  ORACLE RULE: lower anchor index first
    inside a class; (score desc, class asc, anchor index asc) in the merge.
    """
    boxes = np.asarray(boxes, dtype=F32)
    scores = np.asarray(scores, dtype=F32)
    B, N, q, _ = boxes.shape
    L = scores.shape[2]
    T = max_total_size
    out_b = np.zeros((B, T, 4), F32)
    out_s = np.zeros((B, T), F32)
    out_c = np.zeros((B, T), F32)
    out_i = np.full((B, T), -1, np.int32)
    valid = np.zeros((B,), np.int32)
    thr = F32(iou_threshold)
    for b in range(B):
        kept: List[Tuple[float, int, int]] = []
        for c in range(L):
            sc = scores[b, :, c]
            cand = np.nonzero(sc > F32(score_threshold))[0]
            if cand.size == 0:
                continue
            cand = cand[np.lexsort((cand, -sc[cand].astype(np.float64)))]
            bx = boxes[b, :, c if q > 1 else 0, :]
            sel: List[int] = []
            for i in cand:
                if len(sel) >= max_output_size_per_class:
                    break
                if sel and np.any(_nms_iou(bx[i], bx[np.asarray(sel)]) > thr):
                    continue
                sel.append(int(i))
            kept.extend((float(sc[i]), c, i) for i in sel)
        kept.sort(key=lambda t: (-t[0], t[1], t[2]))
        kept = kept[:T]
        valid[b] = len(kept)
        for r, (s, c, i) in enumerate(kept):
            bb = boxes[b, i, c if q > 1 else 0, :]
            out_b[b, r] = np.clip(bb, F32(0), F32(1)) if clip_boxes else bb
            out_s[b, r] = s
            out_c[b, r] = c
            out_i[b, r] = i
    if return_indices:
        return out_b, out_s, out_c, valid, out_i
    return out_b, out_s, out_c, valid


def ssd_decode(priors, variances, pred_deltas, pred_label_probs, max_total_size: int = 200,
               score_threshold: float = 0.5, return_aux: bool = False):
    """models/decoder.py:60-93 -> ``(boxes, labels, scores)`` (note the order).

    An anchor whose argmax over all L columns is 0 has its WHOLE score row
    zeroed; any other row is passed intact, background column included (:78-83).
    NMS iou_threshold is TensorFlow's default 0.5 (not passed at :86-92).
    """
    d = np.asarray(pred_deltas, dtype=F32) * np.asarray(variances, dtype=F32)      # :74
    boxes = boxes_from_deltas(np.asarray(priors, dtype=F32), d)                     # :75
    probs = np.asarray(pred_label_probs, dtype=F32)
    amax = np.argmax(probs, axis=-1)                                                # :78
    scores = np.where((amax != 0)[..., None], probs, F32(0))                        # :79-83
    boxes4 = boxes.reshape(boxes.shape[0], -1, 1, 4)                                # :84
    fb, fs, fc, valid, fi = combined_nms(boxes4, scores, max_total_size, max_total_size,
                                         iou_threshold=0.5, score_threshold=score_threshold,
                                         clip_boxes=True, return_indices=True)
    if return_aux:
        return fb, fc, fs, valid, fi
    return fb, fc, fs


# --------------------------------------------------------------------------
# input pipeline  (utils/data_utils.py:33-37, augmentation.py:119-139)
# --------------------------------------------------------------------------
def preprocess_image(img_u8: np.ndarray, out_h: int, out_w: int, flip: bool = False) -> np.ndarray:
    """``tf.image.convert_image_dtype(img, tf.float32)`` then ``tf.image.resize(img, (out_h, out_w))``
    (utils/data_utils.py:36-37), optionally ``tf.image.flip_left_right`` (augmentation.py:127).  [TF-recall]
    convert = cast * float32(1/255); resize = bilinear with half-pixel centres
    (``in = (out + 0.5) * scale - 0.5``, lower = max(floor, 0), upper = min(ceil, size-1), lerp = in - floor),
    one float32 rounding per operation like the TensorFlow kernel."""
    img = np.asarray(img_u8, dtype=np.uint8)
    H, W = img.shape[:2]
    x = img.astype(F32) * F32(1.0 / 255.0)
    sy, sx = F32(F32(H) / F32(out_h)), F32(F32(W) / F32(out_w))
    in_y = ((np.arange(out_h, dtype=F32) + F32(0.5)) * sy - F32(0.5)).astype(F32)
    in_x = ((np.arange(out_w, dtype=F32) + F32(0.5)) * sx - F32(0.5)).astype(F32)
    fy, fx = np.floor(in_y), np.floor(in_x)
    y0 = np.maximum(fy.astype(np.int64), 0); y1 = np.minimum(np.ceil(in_y).astype(np.int64), H - 1)
    x0 = np.maximum(fx.astype(np.int64), 0); x1 = np.minimum(np.ceil(in_x).astype(np.int64), W - 1)
    ly = (in_y - fy).astype(F32)[:, None, None]
    lx = (in_x - fx).astype(F32)[None, :, None]
    tl, tr = x[y0][:, x0], x[y0][:, x1]
    bl, br = x[y1][:, x0], x[y1][:, x1]
    top = (tl + ((tr - tl).astype(F32) * lx).astype(F32)).astype(F32)
    bot = (bl + ((br - bl).astype(F32) * lx).astype(F32)).astype(F32)
    out = (top + ((bot - top).astype(F32) * ly).astype(F32)).astype(F32)
    return out[:, ::-1].copy() if flip else out


def flip_boxes(gt_boxes: np.ndarray) -> np.ndarray:
    """augmentation.py:128-137: ``[y1, 1 - x2, y2, 1 - x1]`` (all-zero padding rows are left untouched)."""
    b = np.asarray(gt_boxes, dtype=F32)
    out = np.stack([b[..., 0], F32(1.0) - b[..., 3], b[..., 2], F32(1.0) - b[..., 1]], axis=-1).astype(F32)
    pad = ~b.any(axis=-1)
    out[pad] = 0
    return out

