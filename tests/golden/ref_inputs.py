"""Seeded inputs shared by ``make_ref_golden.py`` (which feeds them to the reference's own source under the NumPy
``tensorflow`` shim) and by the tests (which feed the same inputs to the oracle and to the CUDA path).  Pure NumPy; no
arithmetic of the hot path lives here."""

from __future__ import annotations

import zlib
from typing import Dict, Tuple

import numpy as np

AR3 = [1., 2., 1. / 2.]
AR5 = [1., 2., 1. / 2., 3., 1. / 3.]
# utils/train_utils.py:13-34 plus the SSD512 extension -- generate_prior_boxes is generic over it
PRIOR_CONFIGS = {
    "mobilenet_v2": ([19, 10, 5, 3, 2, 1], [AR3, AR5, AR5, AR5, AR3, AR3]),
    "vgg16": ([38, 19, 10, 5, 3, 1], [AR3, AR5, AR5, AR5, AR3, AR3]),
    "vgg16_512": ([64, 32, 16, 8, 4, 2, 1], [AR3, AR5, AR5, AR5, AR5, AR3, AR3]),
}
VARIANCES = [0.1, 0.1, 0.2, 0.2]
NET_SEED = 20261017
# (tag, prior config, batch, seed, background logit bias) of the SSDDecoder fixtures: the first saturates
# max_total_size = 200, the other two leave fewer than 200 detections (zero padding is exercised)
DECODE_CASES = (("mnv2_dense", "mobilenet_v2", 3, 12, 2.0), ("mnv2_sparse", "mobilenet_v2", 2, 14, 7.0),
                ("vgg16", "vgg16", 1, 13, 8.0))


def variable(name: str, shape: Tuple[int, ...], seed: int = NET_SEED) -> np.ndarray:
    """Deterministic value of the Keras variable ``name`` (order independent: the stream is keyed by the name)."""
    rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
    leaf = name.rsplit("/", 1)[-1]
    if leaf in ("kernel", "depthwise_kernel"):
        fan_in = int(np.prod(shape[:3])) if leaf == "kernel" else int(np.prod(shape[:2]))
        return (rng.standard_normal(shape) * np.sqrt(2.0 / fan_in)).astype(np.float32)
    if leaf == "bias":
        return (0.05 * rng.standard_normal(shape)).astype(np.float32)
    if leaf == "gamma":
        return rng.uniform(0.6, 1.4, shape).astype(np.float32)
    if leaf in ("beta", "moving_mean"):
        return (0.1 * rng.standard_normal(shape)).astype(np.float32)
    if leaf == "moving_variance":
        return rng.uniform(0.5, 1.5, shape).astype(np.float32)
    if leaf == "scale":
        return (20.0 + rng.standard_normal(shape)).astype(np.float32)
    raise KeyError(name)


def weights_for(shapes: Dict[str, Tuple[int, ...]], seed: int = NET_SEED) -> Dict[str, np.ndarray]:
    return {k: variable(k, tuple(s), seed) for k, s in shapes.items()}


def image(batch: int, size: int, seed: int = NET_SEED) -> np.ndarray:
    return np.random.default_rng([seed, size]).random((batch, size, size, 3), dtype=np.float32)


def ground_truth(batch: int, padded: int, seed: int, snap: int = 0, max_boxes: int = 6):
    """``[B,G,4]`` float32 boxes, ``[B,G]`` int32 labels padded with box 0 / label -1 (utils/data_utils.py:140-155)."""
    rng = np.random.default_rng(seed)
    boxes = np.zeros((batch, padded, 4), np.float32)
    labels = np.full((batch, padded), -1, np.int32)
    for b in range(batch):
        g = int(rng.integers(1, min(max_boxes, padded) + 1))
        c = rng.random((g, 2))
        wh = rng.uniform(0.1, 0.6, (g, 2))
        bx = np.clip(np.concatenate([c - wh / 2, c + wh / 2], -1), 0.0, 1.0)
        if snap:
            bx = np.round(bx * snap) / snap
            bx[:, 2:] = np.maximum(bx[:, 2:], bx[:, :2] + 1.0 / snap)
            bx = np.clip(bx, 0.0, 1.0)
        boxes[b, :g] = bx.astype(np.float32)
        labels[b, :g] = rng.integers(1, 21, g).astype(np.int32)
    return boxes, labels


def tie_case():
    """Anchors and ground truth on the k/4 lattice: IoU exactly 0.5, exact IoU ties, a duplicated ground-truth box
    (first maximum must win), a zero-area box and padded rows."""
    grid = [k / 4.0 for k in range(5)]
    lattice = np.array([[y1, x1, y2, x2] for y1 in grid for y2 in grid if y2 > y1
                        for x1 in grid for x2 in grid if x2 > x1], np.float32)
    gt = np.zeros((2, 6, 4), np.float32)
    gt[0, :5] = [[0, 0, .5, .5], [0, 0, .5, 1], [0, 0, .5, .5], [.25, .25, .75, .75], [.5, .5, .5, 1]]
    gt[1, :3] = [[0, 0, 1, 1], [0, .5, 1, 1], [0, 0, 1, .5]]
    lab = np.array([[3, 7, 9, 1, 2, -1], [5, 6, 4, -1, -1, -1]], np.int32)
    return lattice, gt, lab


def head_outputs(batch: int, n_anchors: int, seed: int, hot_fraction: float = 0.03, n_labels: int = 21,
                 background_bias: float = 2.0):
    """``(pred_deltas, probabilities)``; probabilities are a float32 softmax computed HERE (an input, not a result)."""
    rng = np.random.default_rng(seed)
    deltas = rng.standard_normal((batch, n_anchors, 4), dtype=np.float32)
    z = (2.0 * rng.standard_normal((batch, n_anchors, n_labels), dtype=np.float32)).astype(np.float32)
    hot = rng.random((batch, n_anchors)) < hot_fraction
    cls = rng.integers(1, n_labels, (batch, n_anchors))
    bi, ni = np.nonzero(hot)
    z[bi, ni, cls[bi, ni]] += np.float32(8.0)
    z[..., 0] += np.float32(background_bias)
    e = np.exp(z - z.max(-1, keepdims=True))
    return deltas, (e / e.sum(-1, keepdims=True, dtype=np.float32)).astype(np.float32), z


def loss_edge_inputs():
    """[B,N,L] targets / probabilities for the hard-negative-mining corner cases of ssd_loss.py:59-91:
    image 0 no positives; 1 all positives; 2 quirk (3 * n_pos > number of negatives, so positives are "mined" too and
    count twice); 3 exactly tied cross-entropies; 4 a probability below the Keras clip; 5 un-normalised rows."""
    rng = np.random.default_rng(404)
    B, N, L = 6, 300, 21
    lab = np.zeros((B, N), np.int64)
    lab[1] = rng.integers(1, L, N)
    lab[2, 40:140] = rng.integers(1, L, 100)
    lab[3, :5] = 3
    lab[4, ::7] = 2
    lab[5, ::9] = rng.integers(1, L, len(range(0, N, 9)))
    y = np.eye(L, dtype=np.float32)[lab]
    z = rng.standard_normal((B, N, L)).astype(np.float32)
    e = np.exp(z - z.max(-1, keepdims=True))
    p = (e / e.sum(-1, keepdims=True, dtype=np.float32)).astype(np.float32)
    p[3, 5:] = p[3, 5]
    p[4, ::7, 2] = 0.0
    p[4, 3, 0] = 1e-9
    p[5] *= np.float32(3.0)
    ad = (rng.standard_normal((B, N, 4)) * (lab > 0)[..., None]).astype(np.float32)
    ad[1, 7] = 0.0                      # a "positive" whose four deltas are exactly zero: loc_loss_fn drops it (:46)
    pd = (3 * rng.standard_normal((B, N, 4))).astype(np.float32)
    return y, p, ad, pd


def nms_tie_inputs(n_cls: int = 6):
    """Lattice boxes (IoU exactly 0.5 occurs) with DISTINCT scores: every decision is TF-verifiable (no equal-score
    order involved), suppression at exactly the threshold is exercised (strict >)."""
    rng = np.random.default_rng(12)
    k, N, B = 32, 320, 2
    y1 = rng.integers(0, k - 8, N); x1 = rng.integers(0, k - 8, N)
    h = rng.choice([4, 8], N); w = rng.choice([4, 8], N)
    boxes = (np.stack([y1, x1, y1 + h, x1 + w], -1) / k).astype(np.float32)
    cls = rng.integers(0, n_cls, (B, N))
    perm = np.stack([rng.permutation(N) for _ in range(B)])
    sc = (0.5 + (perm + 1) / np.float32(2 * N + 2)).astype(np.float32)          # distinct, in (0.5, 1)
    scores = np.zeros((B, N, n_cls), np.float32)
    np.put_along_axis(scores, cls[..., None], sc[..., None], axis=2)
    return np.broadcast_to(boxes[None, :, None, :], (B, N, 1, 4)).copy(), scores


# ---- augmentation.py fixtures: every random decision of ``augmentation.apply`` spelled out -------------------------
# patch: None (skipped) or {"expand": None | (u_ratio, u_left, u_top), "overlap": index into [.1,.3,.5,.7,.9],
# "crop": window as fractions (y0, x0, h, w) of the current canvas}; flip: bool; brightness / contrast / hue /
# saturation: None (skipped) or the [0,1) sample behind the op's tf.random.uniform call.
AUGMENT_SIZE = (40, 56)                 # H, W of the fixture image (already "resized": utils/data_utils.py:37)
AUGMENT_CASES = (
    {"patch": {"expand": (0.37, 0.61, 0.22), "overlap": 2, "crop": (0.13, 0.21, 0.55, 0.62)}, "flip": True,
     "brightness": 0.83, "contrast": 0.12, "hue": 0.71, "saturation": 0.90},
    {"patch": {"expand": None, "overlap": 0, "crop": (0.05, 0.30, 0.80, 0.45)}, "flip": False,
     "brightness": None, "contrast": 0.95, "hue": None, "saturation": None},
    {"patch": None, "flip": True, "brightness": None, "contrast": None, "hue": 0.08, "saturation": None},
    {"patch": None, "flip": False, "brightness": 0.02, "contrast": None, "hue": None, "saturation": 0.31},
    {"patch": None, "flip": False, "brightness": None, "contrast": None, "hue": None, "saturation": None},
    {"patch": {"expand": (0.999, 0.0, 0.999), "overlap": 4, "crop": (0.0, 0.0, 1.0, 1.0)}, "flip": False,
     "brightness": 0.5, "contrast": 0.5, "hue": 0.5, "saturation": 0.5},
)


def augment_inputs(seed: int = 77):
    """One float image in [0,1] (exact u8/255 values, some saturated pixels) and its boxes."""
    rng = np.random.default_rng(seed)
    H, W = AUGMENT_SIZE
    img = rng.integers(0, 256, (H, W, 3)).astype(np.float32) * np.float32(1.0 / 255.0)
    img[:4, :6] = 0.0                    # grey / black / white patches: zero saturation, zero value
    img[4:8, :6] = 1.0
    img[8:12, :6] = np.float32(0.5)
    boxes = np.array([[0.10, 0.20, 0.55, 0.70], [0.40, 0.05, 0.95, 0.45], [0.0, 0.0, 1.0, 1.0], [0.62, 0.58, 0.70, 0.66]],
                     np.float32)
    return img, boxes


def augment_queue(case) -> list:
    """The uniform samples ``augmentation.apply`` consumes for ``case``, in the order it draws them
    (augmentation.py:26-31, 205-222, 177-181): a gate sample per operation (> 0.5 applies it), then the operation's own."""
    gate = lambda on: 0.75 if on else 0.25
    q = [gate(case["patch"] is not None)]
    if case["patch"] is not None:
        ex = case["patch"]["expand"]
        q.append(gate(ex is not None))
        if ex is not None:
            q.extend(ex)
        q.append(int(case["patch"]["overlap"]))                  # get_random_min_overlap: an int32 draw
    q.append(gate(case["flip"]))
    for k in ("brightness", "contrast", "hue", "saturation"):
        q.append(gate(case[k] is not None))
        if case[k] is not None:
            q.append(case[k])
    return q
