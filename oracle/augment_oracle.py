"""NumPy restatement of the reference's training-time augmentation (``augmentation.py``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Pinned by ``tests/golden/ref_augment.npz``: the UNMODIFIED
``/root/reference/augmentation.py`` executed on the NumPy ``tensorflow`` stand-in with the random draws and the crop
window fed in (``tests/golden/make_ref_golden.py:augment_fixture``).

The reference draws every random number from TensorFlow's global stream; here each function takes the draws as
arguments (``u`` = the uniform [0,1) sample behind one ``tf.random.uniform`` call), in the order the reference makes
them, so the device path can be fed the same numbers.  float32, one rounding per elementary operation.

[TF-recall] ``tf.image.adjust_hue`` / ``adjust_saturation`` are restated as RGB -> HSV -> RGB with the documented
formulas (TensorFlow's fused kernels agree to float tolerance, not bit for bit); ``tf.image.resize`` = bilinear,
half-pixel centres; ``sample_distorted_bounding_box`` is a random search whose RESULT (the window) is an input here.
"""

from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32


def channel_mean(img: np.ndarray) -> np.ndarray:
    """Per-channel mean over height and width, rounded once to float32.  TensorFlow reduces in float32 in an order of
    its own (Eigen's tree), NumPy's stand-in row by row, the device in fixed-order partial sums: the three agree to
    ~1e-6, which the hue / saturation round trip amplifies on near-grey pixels (tolerances in the tests say so)."""
    return np.mean(np.asarray(img, F32), axis=(0, 1), dtype=np.float64).astype(F32)


def _round_half_even(x) -> np.float32:
    return F32(np.round(F32(x)))           # tf.round: half to even, like np.round


# --------------------------------------------------------------------------
# boxes  (utils/bbox_utils.py:217-233)
# --------------------------------------------------------------------------
def renormalize_boxes(boxes: np.ndarray, min_max: Sequence[float]) -> np.ndarray:
    """utils/bbox_utils.py:217-233: ``clip((b - [ymin,xmin,ymin,xmin]) / [ymax-ymin, xmax-xmin, ...], 0, 1)``."""
    b = np.asarray(boxes, F32)
    y_min, x_min, y_max, x_max = [F32(v) for v in min_max]
    lo = np.array([y_min, x_min, y_min, x_min], F32)
    span = np.array([y_max - y_min, x_max - x_min, y_max - y_min, x_max - x_min], F32)
    return np.clip(((b - lo).astype(F32) / span).astype(F32), F32(0), F32(1))


def flip_boxes(boxes: np.ndarray) -> np.ndarray:
    """augmentation.py:128-137."""
    b = np.asarray(boxes, F32)
    return np.stack([b[..., 0], F32(1) - b[..., 3], b[..., 2], F32(1) - b[..., 1]], -1).astype(F32)


# --------------------------------------------------------------------------
# geometric  (augmentation.py:119-220)
# --------------------------------------------------------------------------
def resolve_expand(height: int, width: int, u_ratio: float, u_left: float, u_top: float) -> Dict[str, int]:
    """augmentation.py:177-184: the whole-pixel canvas geometry from the three uniform samples."""
    h, w = F32(height), F32(width)
    ratio = F32(F32(u_ratio) * F32(3) + F32(1))                       # uniform(minval=1, maxval=4)
    final_h = _round_half_even(h * ratio)
    final_w = _round_half_even(w * ratio)
    pad_left = _round_half_even(F32(F32(u_left) * F32(final_w - w)) + F32(0))
    pad_top = _round_half_even(F32(F32(u_top) * F32(final_h - h)) + F32(0))
    return {"pad_top": int(pad_top), "pad_left": int(pad_left), "canvas_h": int(final_h), "canvas_w": int(final_w)}


def expand_image(img: np.ndarray, boxes: np.ndarray, geom: Dict[str, int]) -> Tuple[np.ndarray, np.ndarray]:
    """augmentation.py:164-202: the image on a larger canvas filled with its per-channel mean; boxes renormalised."""
    img = np.asarray(img, F32)
    H, W = img.shape[:2]
    pt, pl, ch, cw = geom["pad_top"], geom["pad_left"], geom["canvas_h"], geom["canvas_w"]
    pb, pr = ch - (H + pt), cw - (W + pl)
    mean = channel_mean(img)
    canvas = np.empty((ch, cw, img.shape[2]), F32)
    canvas[...] = mean
    canvas[pt:pt + H, pl:pl + W] = np.where(img == F32(-1), mean, img)            # the -1 sentinel of :187-193
    h, w = F32(H), F32(W)
    min_max = [F32(-pt) / h, F32(-pl) / w, F32(F32(pb) + h) / h, F32(F32(pr) + w) / w]
    return canvas, renormalize_boxes(boxes, min_max)


def resize_bilinear(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """``tf.image.resize`` (bilinear, half-pixel centres), float32 in, float32 out."""
    x = np.asarray(img, F32)
    H, W = x.shape[:2]
    sy, sx = F32(F32(H) / F32(out_h)), F32(F32(W) / F32(out_w))
    in_y = ((np.arange(out_h, dtype=F32) + F32(0.5)) * sy - F32(0.5)).astype(F32)
    in_x = ((np.arange(out_w, dtype=F32) + F32(0.5)) * sx - F32(0.5)).astype(F32)
    fy, fx = np.floor(in_y), np.floor(in_x)
    y0 = np.maximum(fy.astype(np.int64), 0); y1 = np.minimum(np.ceil(in_y).astype(np.int64), H - 1)
    x0 = np.maximum(fx.astype(np.int64), 0); x1 = np.minimum(np.ceil(in_x).astype(np.int64), W - 1)
    ly = (in_y - fy).astype(F32)[:, None, None]
    lx = (in_x - fx).astype(F32)[None, :, None]
    tl, tr, bl, br = x[y0][:, x0], x[y0][:, x1], x[y1][:, x0], x[y1][:, x1]
    top = (tl + ((tr - tl).astype(F32) * lx).astype(F32)).astype(F32)
    bot = (bl + ((br - bl).astype(F32) * lx).astype(F32)).astype(F32)
    return (top + ((bot - top).astype(F32) * ly).astype(F32)).astype(F32)


def patch(img: np.ndarray, boxes: np.ndarray, expand: Optional[Dict[str, int]],
          crop: Tuple[int, int, int, int]) -> Tuple[np.ndarray, np.ndarray]:
    """augmentation.py:205-234: optional expand, the window ``crop = (y0, x0, h, w)`` (pixels of the current canvas)
    cut out and resized back to the original resolution; boxes renormalised to the window."""
    img = np.asarray(img, F32)
    H, W = img.shape[:2]
    if expand is not None:
        img, boxes = expand_image(img, boxes, expand)
    ch, cw = img.shape[:2]
    y0, x0, h, w = crop
    window = [F32(y0) / F32(ch), F32(x0) / F32(cw), F32(y0 + h) / F32(ch), F32(x0 + w) / F32(cw)]
    out = resize_bilinear(img[y0:y0 + h, x0:x0 + w], H, W)
    return out, renormalize_boxes(boxes, window)


# --------------------------------------------------------------------------
# photometric  (augmentation.py:67-116)
# --------------------------------------------------------------------------
def adjust_brightness(img: np.ndarray, delta: float) -> np.ndarray:
    return (np.asarray(img, F32) + F32(delta)).astype(F32)


def adjust_contrast(img: np.ndarray, factor: float) -> np.ndarray:
    x = np.asarray(img, F32)
    mean = channel_mean(x)[None, None, :]
    return (((x - mean).astype(F32) * F32(factor)).astype(F32) + mean).astype(F32)


def rgb_to_hsv(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, F32)
    r, g, b = x[..., 0], x[..., 1], x[..., 2]
    v = np.maximum(np.maximum(r, g), b)
    spread = (v - np.minimum(np.minimum(r, g), b)).astype(F32)
    with np.errstate(all="ignore"):
        s = np.where(v > 0, (spread / v).astype(F32), F32(0)).astype(F32)
        norm = (F32(1) / (F32(6) * spread).astype(F32)).astype(F32)
        h_r = (norm * (g - b).astype(F32)).astype(F32)
        h_g = ((norm * (b - r).astype(F32)).astype(F32) + F32(2.0 / 6.0)).astype(F32)
        h_b = ((norm * (r - g).astype(F32)).astype(F32) + F32(4.0 / 6.0)).astype(F32)
    h = np.where(r == v, h_r, np.where(g == v, h_g, h_b))
    h = np.where(spread > 0, h, F32(0)).astype(F32)
    h = np.where(h < 0, (h + F32(1)).astype(F32), h).astype(F32)
    return np.stack([h, s, v], -1)


def hsv_to_rgb(x: np.ndarray) -> np.ndarray:
    h, s, v = x[..., 0], x[..., 1], x[..., 2]
    dh = (h * F32(6)).astype(F32)
    dr = np.clip((np.abs(dh - F32(3)) - F32(1)).astype(F32), 0, 1)
    dg = np.clip((F32(2) - np.abs(dh - F32(2))).astype(F32), 0, 1)
    db = np.clip((F32(2) - np.abs(dh - F32(4))).astype(F32), 0, 1)
    oms = (F32(1) - s).astype(F32)
    return np.stack([((oms + (s * d).astype(F32)).astype(F32) * v).astype(F32) for d in (dr, dg, db)], -1)


def adjust_hue(img: np.ndarray, delta: float) -> np.ndarray:
    hsv = rgb_to_hsv(img)
    h = (hsv[..., 0] + F32(delta)).astype(F32)
    hsv[..., 0] = (h - np.floor(h)).astype(F32)
    return hsv_to_rgb(hsv)


def adjust_saturation(img: np.ndarray, factor: float) -> np.ndarray:
    hsv = rgb_to_hsv(img)
    hsv[..., 1] = np.clip((hsv[..., 1] * F32(factor)).astype(F32), 0, 1)
    return hsv_to_rgb(hsv)


def uniform(u: float, lo: float, hi: float) -> np.float32:
    """``tf.random.uniform((), lo, hi)`` from its [0,1) sample: ``u * (hi - lo) + lo`` in float32."""
    return F32(F32(F32(u) * F32(F32(hi) - F32(lo))) + F32(lo))


# --------------------------------------------------------------------------
# the pipeline  (augmentation.py:16-33)
# --------------------------------------------------------------------------
def apply(img: np.ndarray, boxes: np.ndarray, plan: Dict[str, object]) -> Tuple[np.ndarray, np.ndarray]:
    """augmentation.py:16-33 with every random decision resolved in ``plan``:

    ``patch``: None or ``{"expand": None | resolve_expand(...), "crop": (y0, x0, h, w)}``; ``flip``: bool;
    ``brightness`` / ``contrast`` / ``hue`` / ``saturation``: None or the drawn delta / factor.
    Order: patch, flip, brightness, contrast, hue, saturation, clip to [0, 1]."""
    img = np.asarray(img, F32)
    boxes = np.asarray(boxes, F32)
    if plan.get("patch") is not None:
        img, boxes = patch(img, boxes, plan["patch"].get("expand"), plan["patch"]["crop"])
    if plan.get("flip"):
        img, boxes = img[:, ::-1].copy(), flip_boxes(boxes)
    if plan.get("brightness") is not None:
        img = adjust_brightness(img, plan["brightness"])
    if plan.get("contrast") is not None:
        img = adjust_contrast(img, plan["contrast"])
    if plan.get("hue") is not None:
        img = adjust_hue(img, plan["hue"])
    if plan.get("saturation") is not None:
        img = adjust_saturation(img, plan["saturation"])
    return np.clip(img, F32(0), F32(1)), boxes
