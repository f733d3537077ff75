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


# ------------------------------------------------------------------------------------------------------------------
# Ops used by the reference's input pipeline (utils/data_utils.py:33-37) and augmentation.py, from their documented
# behaviour.  float32 images in [0, 1], HWC.
# ------------------------------------------------------------------------------------------------------------------
def convert_image_dtype(image, dtype, saturate=False, name=None):
    """uint8 -> float32: ``cast(image) * (1 / 255)`` (documented scaling to [0, 1]); float -> float: identity."""
    a = tf.convert_to_tensor(image).numpy()
    dt = np.dtype(dtype)
    if a.dtype == np.uint8 and dt.kind == "f":
        return tf.Tensor(a.astype(dt) * dt.type(1.0 / 255.0))
    if a.dtype.kind == "f" and dt.kind == "f":
        return tf.Tensor(a.astype(dt))
    raise NotImplementedError((a.dtype, dt))


def resize(images, size, method="bilinear", preserve_aspect_ratio=False, antialias=False, name=None):
    """tf.image.resize (TF2): bilinear, half-pixel centres, no antialiasing.  Documented sampling rule:
    ``in = (out + 0.5) * scale - 0.5`` with ``scale = in_size / out_size``; the two nearest source pixels (clamped to the
    image) are blended with weight ``in - floor(in)``.  [TF-recall] the order of the two lerps (x first, then y) and the
    float32 evaluation order follow resize_bilinear_op.cc."""
    if method != "bilinear" or antialias or preserve_aspect_ratio:
        raise NotImplementedError
    a = tf.convert_to_tensor(images).numpy().astype(np.float32)
    batched = a.ndim == 4
    if not batched:
        a = a[None]
    oh, ow = [int(float(tf.convert_to_tensor(s).numpy())) for s in size]
    B, H, W, C = a.shape
    f = np.float32

    def axis(out_n, in_n):
        scale = f(in_n) / f(out_n)
        src = (np.arange(out_n, dtype=np.float32) + f(0.5)) * scale - f(0.5)
        lo_f = np.floor(src)
        lo = np.maximum(lo_f.astype(np.int64), 0)
        hi = np.minimum(np.ceil(src).astype(np.int64), in_n - 1)
        return lo, hi, (src - lo_f).astype(np.float32)

    y0, y1, ly = axis(oh, H)
    x0, x1, lx = axis(ow, W)
    lx = lx[None, None, :, None]
    ly = ly[None, :, None, None]
    top = a[:, y0][:, :, x0] + (a[:, y0][:, :, x1] - a[:, y0][:, :, x0]) * lx
    bot = a[:, y1][:, :, x0] + (a[:, y1][:, :, x1] - a[:, y1][:, :, x0]) * lx
    out = (top + (bot - top) * ly).astype(np.float32)
    return tf.Tensor(out if batched else out[0])


def flip_left_right(image):
    return tf.Tensor(np.ascontiguousarray(tf.convert_to_tensor(image).numpy()[..., ::-1, :]))


def adjust_brightness(image, delta):
    """Documented: ``delta`` is added to all components (float images: no clipping)."""
    a = tf.convert_to_tensor(image)
    return a + tf.convert_to_tensor(delta, dtype_hint=a.dtype)


def adjust_contrast(images, contrast_factor):
    """Documented: per channel ``(x - mean) * contrast_factor + mean`` with the mean over height and width."""
    a = tf.convert_to_tensor(images).numpy()
    mean = np.mean(a, axis=(-3, -2), keepdims=True, dtype=np.float32)
    cf = tf.convert_to_tensor(contrast_factor, dtype_hint=a.dtype).numpy()
    return tf.Tensor(((a - mean) * cf + mean).astype(np.float32))


def rgb_to_hsv(images, name=None):
    """Documented conversion: v = max, s = (max - min) / max, h in [0, 1) by the usual piecewise formula."""
    a = tf.convert_to_tensor(images).numpy().astype(np.float32)
    r, g, b = a[..., 0], a[..., 1], a[..., 2]
    v = np.maximum(np.maximum(r, g), b)
    mn = np.minimum(np.minimum(r, g), b)
    rng = v - mn
    with np.errstate(all="ignore"):
        s = np.where(v > 0, rng / v, np.float32(0)).astype(np.float32)
        norm = np.float32(1) / (np.float32(6) * rng)
        h = np.where(r == v, norm * (g - b), np.where(g == v, norm * (b - r) + np.float32(2.0 / 6.0), norm * (r - g) + np.float32(4.0 / 6.0)))
    h = np.where(rng > 0, h, np.float32(0)).astype(np.float32)
    h = np.where(h < 0, h + np.float32(1), h).astype(np.float32)
    return tf.Tensor(np.stack([h, s, v], -1).astype(np.float32))


def hsv_to_rgb(images, name=None):
    """Documented inverse of rgb_to_hsv (piecewise linear in the hue sextant)."""
    a = tf.convert_to_tensor(images).numpy().astype(np.float32)
    h, s, v = a[..., 0], a[..., 1], a[..., 2]
    f = np.float32
    dh = h * f(6)
    dr = np.clip(np.abs(dh - f(3)) - f(1), 0, 1)
    dg = np.clip(f(2) - np.abs(dh - f(2)), 0, 1)
    db = np.clip(f(2) - np.abs(dh - f(4)), 0, 1)
    one_minus_s = f(1) - s
    rgb = np.stack([(one_minus_s + s * dr) * v, (one_minus_s + s * dg) * v, (one_minus_s + s * db) * v], -1)
    return tf.Tensor(rgb.astype(np.float32))


def adjust_hue(image, delta, name=None):
    """Documented: RGB -> HSV, ``delta`` added to the hue channel (wrapping around 1), HSV -> RGB."""
    hsv = rgb_to_hsv(image).numpy()
    d = tf.convert_to_tensor(delta, dtype_hint=np.float32).numpy().astype(np.float32)
    h = hsv[..., 0] + d
    h = (h - np.floor(h)).astype(np.float32)
    return hsv_to_rgb(np.stack([h, hsv[..., 1], hsv[..., 2]], -1))


def adjust_saturation(image, saturation_factor, name=None):
    """Documented: RGB -> HSV, the saturation channel multiplied by ``saturation_factor`` (kept inside [0, 1]), HSV -> RGB."""
    hsv = rgb_to_hsv(image).numpy()
    sf = tf.convert_to_tensor(saturation_factor, dtype_hint=np.float32).numpy().astype(np.float32)
    s = np.clip(hsv[..., 1] * sf, 0, 1).astype(np.float32)
    return hsv_to_rgb(np.stack([hsv[..., 0], s, hsv[..., 2]], -1))


def random_brightness(image, max_delta, seed=None):
    """Documented: ``delta`` drawn uniformly from ``[-max_delta, max_delta)``."""
    delta = tf.random.uniform([], -max_delta, max_delta)
    return adjust_brightness(image, delta)


def random_contrast(image, lower, upper, seed=None):
    return adjust_contrast(image, tf.random.uniform([], lower, upper))


def random_hue(image, max_delta, seed=None):
    return adjust_hue(image, tf.random.uniform([], -max_delta, max_delta))


def random_saturation(image, lower, upper, seed=None):
    return adjust_saturation(image, tf.random.uniform([], lower, upper))


#: crop windows handed to the next ``sample_distorted_bounding_box`` calls: (y0, x0, height, width) in pixels.  The
#: sampler inside TensorFlow is a random search whose stream cannot be reproduced, so fixtures choose the window.
crop_queue = []
#: every window handed out, in pixels (so a fixture can record what the reference code was given)
crop_log = []


def sample_distorted_bounding_box(image_size, bounding_boxes, seed=None, min_object_covered=0.1, aspect_ratio_range=None,
                                  area_range=None, max_attempts=None, use_image_if_no_bounding_boxes=None, name=None):
    """Documented outputs: ``begin = [y0, x0, 0]``, ``size = [h, w, -1]`` and ``bboxes [1, 1, 4]`` = the window in
    normalised ``[y_min, x_min, y_max, x_max]`` coordinates.  The window itself comes from ``crop_queue``."""
    shp = tf.convert_to_tensor(image_size).numpy()
    H, W = int(shp[0]), int(shp[1])
    y0, x0, h, w = crop_queue.pop(0)
    if any(isinstance(v, float) for v in (y0, x0, h, w)):        # fractions of the (possibly expanded) canvas
        y0, x0 = int(y0 * H), int(x0 * W)
        h, w = max(1, min(int(h * H), H - y0)), max(1, min(int(w * W), W - x0))
    crop_log.append((y0, x0, h, w, H, W))
    f = np.float32
    box = np.array([[[f(y0) / f(H), f(x0) / f(W), f(y0 + h) / f(H), f(x0 + w) / f(W)]]], np.float32)
    return (tf.Tensor(np.array([y0, x0, 0], np.int32)), tf.Tensor(np.array([h, w, -1], np.int32)), tf.Tensor(box))
