"""Training-time augmentation with the reference's names (``augmentation.py`` of FurkanOM/tf-ssd), on the device.

The reference chains TensorFlow image ops per example inside ``tf.data`` (augmentation.py:16-33: patch, flip, then
brightness / contrast / hue / saturation, each behind a coin flip, then clip to [0, 1]).  Here the random decisions are
drawn on the host IN THE REFERENCE'S ORDER into one 16-word plan per image and a whole batch is transformed by
``ssd_augment_batch`` (three HBM-bound passes + a box kernel, ``csrc/augment_kernels.cu``).  ``apply(img, gt_boxes)``
keeps the reference's call shape -- one ``[H,W,3]`` image and its ``[G,4]`` boxes -- and also takes a batch
``[B,H,W,3]`` / ``[B,G,4]`` (zero rows = batch padding, left untouched).

The random stream is a ``Draws`` object: ``RandomDraws`` (NumPy generator; the stream of ``tf.random`` is not
reproducible anyway) or ``ReplayDraws`` (explicit samples: tests replay the draws the reference's code consumed).
[TF-recall] ``tf.image.sample_distorted_bounding_box`` is restated from its documentation as a bounded random search.
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32
PATCH, EXPAND, FLIP, BRIGHTNESS, CONTRAST, HUE, SATURATION, NO_CLIP = 1, 2, 4, 8, 16, 32, 64, 128
MIN_OVERLAPS = (0.1, 0.3, 0.5, 0.7, 0.9)                 # augmentation.py:148-161
Plan = Dict[str, Any]


# ------------------------------------------------------------------ draws --
class RandomDraws(object):
    """Host random stream: ``uniform()`` in [0,1), ``integer(n)`` in [0,n), ``crop(...)`` = the distorted-box sampler."""

    def __init__(self, seed: Any = None, max_attempts: int = 100):
        self.rng = seed if isinstance(seed, np.random.Generator) else np.random.default_rng(seed)
        self.max_attempts = max_attempts

    def uniform(self) -> float:
        return float(self.rng.random(dtype=np.float32))

    def integer(self, n: int) -> int:
        return int(self.rng.integers(0, n))

    def crop(self, canvas_h: int, canvas_w: int, boxes: np.ndarray, min_object_covered: float,
             aspect_ratio_range: Sequence[float] = (0.5, 2.0), area_range: Sequence[float] = (0.05, 1.0)) -> Tuple[int, int, int, int]:
        """``tf.image.sample_distorted_bounding_box`` (augmentation.py:222-227), from its documentation: a window with
        aspect ratio in ``aspect_ratio_range`` and area fraction in ``area_range`` that contains at least
        ``min_object_covered`` of one of the boxes; the whole image after ``max_attempts`` failures."""
        boxes = np.asarray(boxes, np.float64).reshape(-1, 4)
        boxes = boxes[(boxes[:, 2] > boxes[:, 0]) & (boxes[:, 3] > boxes[:, 1])]
        area_px = boxes.copy()
        area_px[:, [0, 2]] *= canvas_h
        area_px[:, [1, 3]] *= canvas_w
        for _ in range(self.max_attempts):
            aspect = self.rng.uniform(aspect_ratio_range[0], aspect_ratio_range[1])
            area = self.rng.uniform(area_range[0], area_range[1]) * canvas_h * canvas_w
            h = int(round(np.sqrt(area / aspect)))
            w = int(round(h * aspect))
            if h < 1 or w < 1 or h > canvas_h or w > canvas_w:
                continue
            y0 = int(self.rng.integers(0, canvas_h - h + 1))
            x0 = int(self.rng.integers(0, canvas_w - w + 1))
            if len(area_px) == 0:
                return y0, x0, h, w
            ih = np.minimum(area_px[:, 2], y0 + h) - np.maximum(area_px[:, 0], y0)
            iw = np.minimum(area_px[:, 3], x0 + w) - np.maximum(area_px[:, 1], x0)
            inter = np.clip(ih, 0, None) * np.clip(iw, 0, None)
            box_area = (area_px[:, 2] - area_px[:, 0]) * (area_px[:, 3] - area_px[:, 1])
            if np.any(inter >= min_object_covered * box_area):
                return y0, x0, h, w
        return 0, 0, canvas_h, canvas_w


class ReplayDraws(object):
    """Replays explicit samples: ``samples`` in the order ``make_plan`` draws them (gates and values; the min-overlap
    index as an int), ``crops`` = the windows ``(y0, x0, h, w)`` in canvas pixels."""

    def __init__(self, samples: Sequence[float], crops: Sequence[Tuple[int, int, int, int]] = ()):
        self.samples, self.crops = list(samples), list(crops)

    def uniform(self) -> float:
        return float(self.samples.pop(0))

    def integer(self, n: int) -> int:
        v = int(self.samples.pop(0))
        if not 0 <= v < n:
            raise ValueError(f"replayed integer {v} outside [0, {n})")
        return v

    def crop(self, canvas_h: int, canvas_w: int, boxes: Any, min_object_covered: float, **kwargs) -> Tuple[int, int, int, int]:
        y0, x0, h, w = [int(v) for v in self.crops.pop(0)]
        if not (0 <= y0 and 0 <= x0 and h >= 1 and w >= 1 and y0 + h <= canvas_h and x0 + w <= canvas_w):
            raise ValueError(f"replayed window {(y0, x0, h, w)} outside the {canvas_h}x{canvas_w} canvas")
        return y0, x0, h, w


# ------------------------------------------------------------------- plans --
def _round(x) -> np.float32:
    return F32(np.round(F32(x)))                 # tf.round: half to even


def _uniform(u: float, lo: float, hi: float) -> np.float32:
    """``tf.random.uniform((), lo, hi)`` from its [0,1) sample: ``u * (hi - lo) + lo`` in float32."""
    return F32(F32(F32(u) * F32(F32(hi) - F32(lo))) + F32(lo))


def get_random_bool(draws) -> bool:
    """augmentation.py:36-42."""
    return F32(draws.uniform()) > F32(0.5)


def get_random_min_overlap(draws) -> float:
    """augmentation.py:148-161."""
    return MIN_OVERLAPS[draws.integer(len(MIN_OVERLAPS))]


def expand_geometry(height: int, width: int, draws) -> Dict[str, int]:
    """augmentation.py:177-184: expansion ratio in [1,4), rounded canvas size, rounded random offsets (float32)."""
    h, w = F32(height), F32(width)
    ratio = _uniform(draws.uniform(), 1, 4)
    final_h, final_w = _round(h * ratio), _round(w * ratio)
    pad_left = _round(_uniform(draws.uniform(), 0, F32(final_w - w)))
    pad_top = _round(_uniform(draws.uniform(), 0, F32(final_h - h)))
    return {"pad_top": int(pad_top), "pad_left": int(pad_left), "canvas_h": int(final_h), "canvas_w": int(final_w)}


def _boxes_on_canvas(boxes: np.ndarray, height: int, width: int, geom: Optional[Dict[str, int]]) -> np.ndarray:
    """The boxes the sampler sees (augmentation.py:196-201): renormalised to the expanded canvas (host float64 is
    enough -- they only steer the random search, the device recomputes them in float32)."""
    b = np.asarray(boxes, np.float64).reshape(-1, 4)
    b = b[np.any(b != 0, axis=1)]
    if geom is None or len(b) == 0:
        return b
    out = b.copy()
    out[:, [0, 2]] = (b[:, [0, 2]] * height + geom["pad_top"]) / geom["canvas_h"]
    out[:, [1, 3]] = (b[:, [1, 3]] * width + geom["pad_left"]) / geom["canvas_w"]
    return np.clip(out, 0.0, 1.0)


def make_plan(height: int, width: int, gt_boxes: Any, draws) -> Plan:
    """One example's random decisions in the order ``augmentation.apply`` makes them (augmentation.py:26-31):
    patch [expand [ratio, left, top], min overlap, window], flip, brightness, contrast, hue, saturation."""
    plan: Plan = {"patch": None, "flip": False, "brightness": None, "contrast": None, "hue": None, "saturation": None}
    if get_random_bool(draws):                                              # patch (:205-234)
        geom = expand_geometry(height, width, draws) if get_random_bool(draws) else None
        min_overlap = get_random_min_overlap(draws)
        ch, cw = (geom["canvas_h"], geom["canvas_w"]) if geom else (height, width)
        window = draws.crop(ch, cw, _boxes_on_canvas(gt_boxes, height, width, geom), min_overlap)
        plan["patch"] = {"expand": geom, "crop": tuple(int(v) for v in window), "min_overlap": min_overlap}
    plan["flip"] = bool(get_random_bool(draws))                             # flip_horizontally (:119-139)
    if get_random_bool(draws):
        plan["brightness"] = float(_uniform(draws.uniform(), -0.12, 0.12))   # random_brightness (:67-78)
    if get_random_bool(draws):
        plan["contrast"] = float(_uniform(draws.uniform(), 0.5, 1.5))        # random_contrast (:81-90)
    if get_random_bool(draws):
        plan["hue"] = float(_uniform(draws.uniform(), -0.08, 0.08))          # random_hue (:93-104)
    if get_random_bool(draws):
        plan["saturation"] = float(_uniform(draws.uniform(), 0.5, 1.5))      # random_saturation (:107-116)
    return plan


def pack_plans(plans: Sequence[Plan], height: int, width: int) -> np.ndarray:
    """``[B,16]`` int32 words in the layout ``ssd_augment_batch`` documents (include/ssd_b200.h)."""
    words = np.zeros((len(plans), 16), np.int32)
    values = words.view(np.float32)
    for i, p in enumerate(plans):
        flags = 0
        geom = {"pad_top": 0, "pad_left": 0, "canvas_h": height, "canvas_w": width}
        crop = (0, 0, height, width)
        if p.get("patch") is not None:
            flags |= PATCH
            if p["patch"].get("expand") is not None:
                flags |= EXPAND
                geom = p["patch"]["expand"]
            crop = p["patch"]["crop"]
            y0, x0, h, w = crop
            if not (0 <= y0 and 0 <= x0 and h >= 1 and w >= 1 and y0 + h <= geom["canvas_h"] and x0 + w <= geom["canvas_w"]):
                raise ValueError(f"crop window {crop} outside the canvas {geom}")
        flags |= FLIP if p.get("flip") else 0
        flags |= NO_CLIP if p.get("no_clip") else 0
        words[i, 1:5] = [geom["pad_top"], geom["pad_left"], geom["canvas_h"], geom["canvas_w"]]
        words[i, 5:9] = crop
        for bit, key, slot in ((BRIGHTNESS, "brightness", 9), (CONTRAST, "contrast", 10), (HUE, "hue", 11), (SATURATION, "saturation", 12)):
            if p.get(key) is not None:
                flags |= bit
                values[i, slot] = F32(p[key])
        words[i, 0] = flags
    return words


# ------------------------------------------------------------------ device --
_default_draws = RandomDraws()
_NO_OP: Plan = {"patch": None, "flip": False, "brightness": None, "contrast": None, "hue": None, "saturation": None}


def _as_batch(img: Any, gt_boxes: Any):
    from tf_ssd_b200 import _ffi
    x, boxes = _ffi.to_dev(img), _ffi.to_dev(gt_boxes)
    single = x.dim() == 3
    if single:
        x, boxes = x[None], boxes[None]
    if x.dim() != 4 or x.shape[3] != 3 or boxes.dim() != 3 or boxes.shape[0] != x.shape[0] or boxes.shape[2] != 4:
        raise ValueError("expected an image [H,W,3] with boxes [G,4], or a batch [B,H,W,3] with boxes [B,G,4]")
    return x.contiguous(), boxes.contiguous(), single


def apply_plans(img: Any, gt_boxes: Any, plans: Sequence[Plan], out_size: Optional[Tuple[int, int]] = None):
    """Runs resolved plans on the device: ``img`` ``[B,H,W,3]`` float32, ``gt_boxes`` ``[B,G,4]`` -> new tensors.
    ``out_size`` (default: the input size, what ``patch`` resizes back to, augmentation.py:231) must be the input size
    unless every plan has a patch."""
    import torch
    from tf_ssd_b200 import _ffi
    x, boxes, single = _as_batch(img, gt_boxes)
    boxes = boxes.clone()
    B, H, W = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
    Ho, Wo = (H, W) if out_size is None else (int(out_size[0]), int(out_size[1]))
    if len(plans) != B:
        raise ValueError(f"{len(plans)} plans for a batch of {B}")
    if (Ho, Wo) != (H, W) and any(p.get("patch") is None for p in plans):
        raise ValueError("an output size different from the input needs a patch in every plan")
    d_plans = torch.from_numpy(pack_plans(plans, H, W)).to(x.device)
    out = torch.empty((B, Ho, Wo, 3), dtype=torch.float32, device=x.device)
    lib = _ffi.lib()
    ws = _ffi.workspace(int(lib.ssd_augment_workspace_bytes(B, H, W, Ho, Wo)))
    _ffi.check(lib.ssd_augment_batch(_ffi.ptr(x), _ffi.ptr(out), _ffi.ptr(boxes) if boxes.numel() else None, B, H, W, Ho, Wo,
                                     int(boxes.shape[1]), _ffi.ptr(d_plans), _ffi.ptr(ws), ws.numel() * ws.element_size(),
                                     _ffi.stream()), "ssd_augment_batch")
    return (out[0], boxes[0]) if single else (out, boxes)


def apply(img: Any, gt_boxes: Any, draws: Any = None):
    """augmentation.py:16-33.  ``img`` ``[H,W,3]`` + ``gt_boxes`` ``[G,4]`` (the reference's call) or a batch
    ``[B,H,W,3]`` + ``[B,G,4]``; float32 in [0,1] -> augmented image(s) and adjusted boxes (CUDA tensors)."""
    draws = draws or _default_draws
    x, boxes, single = _as_batch(img, gt_boxes)
    host_boxes = boxes.detach().cpu().numpy()
    plans = [make_plan(int(x.shape[1]), int(x.shape[2]), host_boxes[i], draws) for i in range(x.shape[0])]
    out, out_boxes = apply_plans(x, boxes, plans)
    return (out[0], out_boxes[0]) if single else (out, out_boxes)


def randomly_apply_operation(operation, img: Any, gt_boxes: Any, *args, draws: Any = None):
    """augmentation.py:45-64."""
    draws = draws or _default_draws
    if get_random_bool(draws):
        return operation(img, gt_boxes, *args, draws=draws)
    from tf_ssd_b200 import _ffi
    return _ffi.to_dev(img), _ffi.to_dev(gt_boxes)


def _one_op(img: Any, gt_boxes: Any, **fields):
    x, boxes, single = _as_batch(img, gt_boxes)
    plan = dict(_NO_OP, no_clip=True, **fields)           # the single operations do not clip; only apply() does (:32)
    out, out_boxes = apply_plans(x, boxes, [plan] * int(x.shape[0]))
    return (out[0], out_boxes[0]) if single else (out, out_boxes)


def random_brightness(img: Any, gt_boxes: Any, max_delta: float = 0.12, draws: Any = None):
    """augmentation.py:67-78: ``delta`` uniform in [-max_delta, max_delta) added to every component."""
    return _one_op(img, gt_boxes, brightness=float(_uniform((draws or _default_draws).uniform(), -max_delta, max_delta)))


def random_contrast(img: Any, gt_boxes: Any, lower: float = 0.5, upper: float = 1.5, draws: Any = None):
    """augmentation.py:81-90: ``(x - mean) * factor + mean`` with the per-channel mean over height and width."""
    return _one_op(img, gt_boxes, contrast=float(_uniform((draws or _default_draws).uniform(), lower, upper)))


def random_hue(img: Any, gt_boxes: Any, max_delta: float = 0.08, draws: Any = None):
    """augmentation.py:93-104."""
    return _one_op(img, gt_boxes, hue=float(_uniform((draws or _default_draws).uniform(), -max_delta, max_delta)))


def random_saturation(img: Any, gt_boxes: Any, lower: float = 0.5, upper: float = 1.5, draws: Any = None):
    """augmentation.py:107-116."""
    return _one_op(img, gt_boxes, saturation=float(_uniform((draws or _default_draws).uniform(), lower, upper)))


def flip_horizontally(img: Any, gt_boxes: Any, draws: Any = None):
    """augmentation.py:119-139."""
    return _one_op(img, gt_boxes, flip=True)


def expand_image(img: Any, gt_boxes: Any, height: Any = None, width: Any = None, draws: Any = None):
    """augmentation.py:164-202: the image on a ``canvas_h x canvas_w`` canvas filled with its per-channel mean, boxes
    renormalised.  One example at a time (every example gets its own canvas size); ``patch`` fuses it for batches."""
    x, boxes, single = _as_batch(img, gt_boxes)
    if x.shape[0] != 1:
        raise ValueError("expand_image produces a different canvas size per example: pass one image (or use patch)")
    H, W = int(x.shape[1]), int(x.shape[2])
    geom = expand_geometry(H if height is None else int(height), W if width is None else int(width), draws or _default_draws)
    plan = dict(_NO_OP, no_clip=True, patch={"expand": geom, "crop": (0, 0, geom["canvas_h"], geom["canvas_w"])})
    out, out_boxes = apply_plans(x, boxes, [plan], out_size=(geom["canvas_h"], geom["canvas_w"]))
    return (out[0], out_boxes[0]) if single else (out, out_boxes)


def patch(img: Any, gt_boxes: Any, draws: Any = None):
    """augmentation.py:205-234 (always applied; ``apply`` gates it with a coin flip): optional expand, a distorted
    window that keeps ``min_overlap`` of some object, resized back to the input resolution; boxes follow."""
    draws = draws or _default_draws
    x, boxes, single = _as_batch(img, gt_boxes)
    H, W = int(x.shape[1]), int(x.shape[2])
    host_boxes = boxes.detach().cpu().numpy()
    plans: List[Plan] = []
    for i in range(x.shape[0]):
        geom = expand_geometry(H, W, draws) if get_random_bool(draws) else None
        min_overlap = get_random_min_overlap(draws)
        ch, cw = (geom["canvas_h"], geom["canvas_w"]) if geom else (H, W)
        window = draws.crop(ch, cw, _boxes_on_canvas(host_boxes[i], H, W, geom), min_overlap)
        plans.append(dict(_NO_OP, no_clip=True, patch={"expand": geom, "crop": window}))
    out, out_boxes = apply_plans(x, boxes, plans)
    return (out[0], out_boxes[0]) if single else (out, out_boxes)
