"""Seeded synthetic VOC-shaped inputs.

There is no network for datasets, so benchmarks and tests use these
generators: 20 classes + background, ground truth padded with box 0 / label -1
(the reference's ``padded_batch`` values, utils/data_utils.py:140-155).
Pure NumPy host code; no arithmetic of the hot path lives here.
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def make_images(batch: int, size: int, seed: int = 1234) -> np.ndarray:
    """NHWC float32 in [0,1) like ``convert_image_dtype`` output (data_utils.py:36)."""
    rng = np.random.default_rng(seed)
    return rng.random((batch, size, size, 3), dtype=np.float32)


def make_ground_truth(batch: int, padded: int = 16, max_boxes: int = 8, n_classes: int = 20, seed: int = 1234,
                      snap: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
    """``gt_boxes [B,G,4]`` float32, ``gt_labels [B,G]`` int32 (-1 padding).

    Per image ``g ~ U{1..max_boxes}`` boxes, centres U(0,1), sides U(0.1,0.6),
    corners clipped to [0,1].  ``snap=k`` rounds corners to multiples of 1/k so
    that exact IoU ties and IoU == 0.5 occur (integer-tie variant)."""
    rng = np.random.default_rng(seed)
    boxes = np.zeros((batch, padded, 4), np.float32)
    labels = np.full((batch, padded), -1, np.int32)
    for b in range(batch):
        g = int(rng.integers(1, min(max_boxes, padded) + 1))
        c = rng.random((g, 2))
        wh = rng.uniform(0.1, 0.6, (g, 2))
        y1, x1 = c[:, 0] - wh[:, 0] / 2, c[:, 1] - wh[:, 1] / 2
        y2, x2 = c[:, 0] + wh[:, 0] / 2, c[:, 1] + wh[:, 1] / 2
        bx = np.clip(np.stack([y1, x1, y2, x2], -1), 0.0, 1.0)
        if snap:
            bx = np.round(bx * snap) / snap
            bx[:, 2] = np.maximum(bx[:, 2], bx[:, 0] + 1.0 / snap)
            bx[:, 3] = np.maximum(bx[:, 3], bx[:, 1] + 1.0 / snap)
            bx = np.clip(bx, 0.0, 1.0)
        boxes[b, :g] = bx.astype(np.float32)
        labels[b, :g] = rng.integers(1, n_classes + 1, g).astype(np.int32)
    return boxes, labels


def make_head_outputs(batch: int, n_anchors: int, n_labels: int = 21, seed: int = 1234, hot_fraction: float = 0.02,
                      hot_boost: float = 8.0, background_bias: float = 0.0) -> Tuple[np.ndarray, np.ndarray]:
    """``(pred_deltas [B,N,4], logits [B,N,L])``: logits N(0, 2^2) with
    ``hot_fraction`` of the anchors boosted on one random foreground class so a
    realistic number of NMS candidates (tens to hundreds per image) appears."""
    rng = np.random.default_rng(seed)
    deltas = rng.standard_normal((batch, n_anchors, 4), dtype=np.float32)
    logits = (2.0 * rng.standard_normal((batch, n_anchors, n_labels), dtype=np.float32)).astype(np.float32)
    hot = rng.random((batch, n_anchors)) < hot_fraction
    cls = rng.integers(1, n_labels, (batch, n_anchors))
    bi, ni = np.nonzero(hot)
    logits[bi, ni, cls[bi, ni]] += np.float32(hot_boost)
    if background_bias:
        # detector-like: background dominates everywhere but on the hot anchors
        logits[..., 0] += np.float32(background_bias)
        logits[bi, ni, 0] -= np.float32(background_bias)
    return deltas, logits
