"""tf.image ops used by the reference (NumPy; see the package docstring)."""

from __future__ import annotations

import numpy as np

import tensorflow as tf


def _nms_iou(b: np.ndarray, i: int, j: int) -> np.float32:
    """[TF-recall] core/kernels/non_max_suppression_op.cc IOU(): corners canonicalised with min/max, 0 when either
    area is <= 0, float32 arithmetic."""
    f = np.float32
    ymin_i, xmin_i = min(b[i, 0], b[i, 2]), min(b[i, 1], b[i, 3])
    ymax_i, xmax_i = max(b[i, 0], b[i, 2]), max(b[i, 1], b[i, 3])
    ymin_j, xmin_j = min(b[j, 0], b[j, 2]), min(b[j, 1], b[j, 3])
    ymax_j, xmax_j = max(b[j, 0], b[j, 2]), max(b[j, 1], b[j, 3])
    area_i = f(f(ymax_i - ymin_i) * f(xmax_i - xmin_i))
    area_j = f(f(ymax_j - ymin_j) * f(xmax_j - xmin_j))
    if area_i <= 0 or area_j <= 0:
        return f(0)
    iy0, ix0 = max(ymin_i, ymin_j), max(xmin_i, xmin_j)
    iy1, ix1 = min(ymax_i, ymax_j), min(xmax_i, xmax_j)
    inter = f(max(f(iy1 - iy0), f(0)) * max(f(ix1 - ix0), f(0)))
    return f(inter / f(f(area_i + area_j) - inter))


def combined_non_max_suppression(boxes, scores, max_output_size_per_class, max_total_size, iou_threshold=0.5,
                                 score_threshold=float("-inf"), pad_per_class=False, clip_boxes=True, name=None):
    """tf.image.combined_non_max_suppression.

    Documented: boxes ``[batch, num_boxes, q, 4]`` (q = 1: shared by all classes, or q = num_classes), scores
    ``[batch, num_boxes, num_classes]``; per class greedy selection in descending score order, pruning boxes whose IoU
    with an already selected box exceeds ``iou_threshold``; boxes scoring no more than ``score_threshold`` are
    removed; at most ``max_output_size_per_class`` per class and ``max_total_size`` per image (the best by score over
    all classes); outputs padded with zeros; ``clip_boxes`` clips coordinates to [0, 1]; returns
    ``(nmsed_boxes, nmsed_scores, nmsed_classes, valid_detections)``.

    [TF-recall] (kernel BatchedNonMaxSuppressionOp): the score filter is a strict ``>``; suppression is a strict
    ``iou > iou_threshold``; equal scores inside a class are visited in ascending box index; the cross-class merge is
    a sort by descending score (equal scores: class-major order here -- implementation-defined in TensorFlow)."""
    b = tf.convert_to_tensor(boxes).numpy().astype(np.float32)
    s = tf.convert_to_tensor(scores).numpy().astype(np.float32)
    B, N, q, _ = b.shape
    L = s.shape[2]
    if q not in (1, L):
        raise ValueError("boxes.shape[2] must be 1 or num_classes")
    per_class, total = int(max_output_size_per_class), int(max_total_size)
    if pad_per_class:
        total = min(total, per_class * L)
    thr, iou_thr = np.float32(score_threshold), np.float32(iou_threshold)
    out_b = np.zeros((B, total, 4), np.float32)
    out_s = np.zeros((B, total), np.float32)
    out_c = np.zeros((B, total), np.float32)
    valid = np.zeros((B,), np.int32)
    for n in range(B):
        merged = []                                     # (score, class, box index)
        for c in range(L):
            cb = b[n, :, 0 if q == 1 else c, :]
            cand = [i for i in range(N) if s[n, i, c] > thr]
            cand.sort(key=lambda i: (-float(s[n, i, c]), i))
            kept = []
            for i in cand:
                if len(kept) >= per_class:
                    break
                ok = True
                for j in reversed(kept):
                    if _nms_iou(cb, i, j) > iou_thr:
                        ok = False
                        break
                if ok:
                    kept.append(i)
            merged += [(float(s[n, i, c]), c, i) for i in kept]
        merged.sort(key=lambda t: -t[0])                # stable: equal scores stay class-major
        merged = merged[:total]
        valid[n] = len(merged)
        for k, (sc, c, i) in enumerate(merged):
            box = b[n, i, 0 if q == 1 else c, :]
            out_b[n, k] = np.clip(box, 0, 1) if clip_boxes else box
            out_s[n, k] = sc
            out_c[n, k] = c
    return tf.Tensor(out_b), tf.Tensor(out_s), tf.Tensor(out_c), tf.Tensor(valid)
