"""Box helpers with the reference's names and argument meanings
(``utils/bbox_utils.py`` of FurkanOM/tf-ssd), computed by sm_100a kernels.

Inputs may be NumPy arrays, CPU tensors or CUDA tensors; results are float32
CUDA tensors (``.cpu().numpy()`` gives what the reference's ``.numpy()`` gave).
Boxes are normalised ``[y1, x1, y2, x2]``.
"""

from __future__ import annotations

import ctypes as C
from typing import Any, Optional, Sequence

import numpy as np
import torch

from tf_ssd_b200 import _ffi


def non_max_suppression(pred_bboxes: Any, pred_labels: Any, **kwargs: Any):
    """utils/bbox_utils.py:10-21 -- ``tf.image.combined_non_max_suppression``.

    ``pred_bboxes`` ``[B,N,q,4]`` (q = 1 or L), ``pred_labels`` ``[B,N,L]``.
    Keyword arguments are TensorFlow's: ``max_output_size_per_class``,
    ``max_total_size``, ``iou_threshold=0.5``, ``score_threshold=-inf``,
    ``pad_per_class=False``, ``clip_boxes=True``.  Returns TensorFlow's tuple
    ``(boxes [B,T,4], scores [B,T], classes [B,T], valid_detections [B])``.
    """
    per_class = int(kwargs.pop("max_output_size_per_class"))
    total = int(kwargs.pop("max_total_size"))
    iou_thr = float(kwargs.pop("iou_threshold", 0.5))
    score_thr = float(kwargs.pop("score_threshold", float("-inf")))
    pad_per_class = bool(kwargs.pop("pad_per_class", False))
    clip_boxes = bool(kwargs.pop("clip_boxes", True))
    kwargs.pop("name", None)
    if kwargs:
        raise TypeError(f"non_max_suppression got unexpected keyword arguments {sorted(kwargs)}")
    if pad_per_class:
        raise NotImplementedError("pad_per_class=True is not used by the reference and is not implemented")
    _ffi.check_device()
    boxes = _ffi.to_dev(pred_bboxes)
    scores = _ffi.to_dev(pred_labels)
    if boxes.dim() != 4 or scores.dim() != 3 or boxes.shape[-1] != 4:
        raise ValueError("expected boxes [B,N,q,4] and scores [B,N,L]")
    B, N, q, _ = boxes.shape
    L = scores.shape[2]
    lib = _ffi.lib()
    dev = boxes.device
    out_b = torch.empty((B, total, 4), dtype=torch.float32, device=dev)
    out_s = torch.empty((B, total), dtype=torch.float32, device=dev)
    out_c = torch.empty((B, total), dtype=torch.float32, device=dev)
    valid = torch.empty((B,), dtype=torch.int32, device=dev)
    nbytes = lib.ssd_combined_nms_workspace_bytes(B, N, L, per_class, total, 0)
    ws = _ffi.workspace(nbytes)
    _ffi.check(lib.ssd_combined_nms(_ffi.ptr(boxes), _ffi.ptr(scores), B, N, q, L, per_class, total, iou_thr,
                                    score_thr, int(clip_boxes), 0, _ffi.ptr(out_b), _ffi.ptr(out_s),
                                    _ffi.ptr(out_c), _ffi.ptr(valid), _ffi.ptr(ws), ws.numel(), _ffi.stream()),
               "ssd_combined_nms")
    return out_b, out_s, out_c, valid


def generate_iou_map(bboxes: Any, gt_boxes: Any, transpose_perm: Optional[Sequence[int]] = None):
    """utils/bbox_utils.py:24-55.  ``[N,4] x [B,G,4] -> [B,N,G]``,
    ``[B,M,4] x [B,G,4] -> [B,M,G]``, or rank-2 ground truth with
    ``transpose_perm=[1,0]``: ``[N,4] x [G,4] -> [N,G]``."""
    _ffi.check_device()
    b = _ffi.to_dev(bboxes)
    g = _ffi.to_dev(gt_boxes)
    rank2 = g.dim() == 2
    if rank2:
        if transpose_perm is None or list(transpose_perm) != [1, 0]:
            raise ValueError("rank-2 gt_boxes need transpose_perm=[1, 0] (as in the reference)")
        g = g.unsqueeze(0)
    elif transpose_perm is not None and list(transpose_perm) != [0, 2, 1]:
        raise ValueError("rank-3 gt_boxes need transpose_perm=[0, 2, 1] (the reference default)")
    if g.dim() != 3 or g.shape[-1] != 4 or b.shape[-1] != 4 or b.dim() not in (2, 3):
        raise ValueError("expected bboxes [N,4] or [B,N,4] and gt_boxes [B,G,4] or [G,4]")
    B, G = g.shape[0], g.shape[1]
    batched = b.dim() == 3
    if batched and b.shape[0] != B:
        raise ValueError("batch size mismatch between bboxes and gt_boxes")
    N = b.shape[-2]
    out = torch.empty((B, N, G), dtype=torch.float32, device=g.device)
    _ffi.check(_ffi.lib().ssd_iou_map(_ffi.ptr(b), _ffi.ptr(g), B, N, G, int(batched), _ffi.ptr(out), _ffi.stream()),
               "ssd_iou_map")
    return out[0] if rank2 else out


def _pairwise(fn_name: str, priors: Any, other: Any):
    _ffi.check_device()
    p = _ffi.to_dev(priors)
    o = _ffi.to_dev(other)
    if p.shape[-1] != 4 or o.shape[-1] != 4:
        raise ValueError("last dimension must be 4")
    out_shape = torch.broadcast_shapes(p.shape, o.shape)
    if p.dim() > 2 or tuple(o.shape) != tuple(out_shape):
        # general broadcasting: materialise both operands at the output shape
        p = p.expand(out_shape).contiguous()
        o = o.expand(out_shape).contiguous()
        batched, N = 1, 1
        B = int(np.prod(out_shape[:-1]))
    else:
        N = p.shape[0] if p.dim() == 2 else 1
        B = int(np.prod(out_shape[:-1])) // max(N, 1)
        batched = 0
        if p.dim() == 1:
            p = p.reshape(1, 4)
    out = torch.empty(out_shape, dtype=torch.float32, device=o.device)
    fn = getattr(_ffi.lib(), fn_name)
    _ffi.check(fn(_ffi.ptr(p), _ffi.ptr(o), B, N, batched, _ffi.ptr(out), _ffi.stream()), fn_name)
    return out


def get_bboxes_from_deltas(prior_boxes: Any, deltas: Any):
    """utils/bbox_utils.py:58-82 (``[dy,dx,dh,dw]`` -> corners; ``y2 = h + y1``)."""
    return _pairwise("ssd_decode_boxes", prior_boxes, deltas)


def get_deltas_from_bboxes(bboxes: Any, gt_boxes: Any):
    """utils/bbox_utils.py:85-128 (zero prior extent -> 1e-3, zero GT extent -> 0)."""
    return _pairwise("ssd_encode_deltas", bboxes, gt_boxes)


def get_scale_for_nth_feature_map(k: int, m: int = 6, scale_min: float = 0.2, scale_max: float = 0.9) -> float:
    """utils/bbox_utils.py:131-148 (host scalar, float64 like the reference)."""
    return scale_min + ((scale_max - scale_min) / (m - 1)) * (k - 1)


def _flatten_aspect_ratios(aspect_ratios: Sequence[Sequence[float]]):
    flat = [float(a) for row in aspect_ratios for a in row]
    counts = [len(row) for row in aspect_ratios]
    return flat, counts


def generate_base_prior_boxes(aspect_ratios: Sequence[float], feature_map_index: int, total_feature_map: int):
    """utils/bbox_utils.py:151-176 -- the ``[A,4]`` origin-centred boxes of one
    cell (extra square box last).  Host glue with the reference's dtype path
    (float32 sqrt/div/mul, float64 scale product); the hot path never calls it:
    :func:`generate_prior_boxes` evaluates the same arithmetic inside its kernel."""
    s_cur = get_scale_for_nth_feature_map(feature_map_index, m=total_feature_map)
    s_next = get_scale_for_nth_feature_map(feature_map_index + 1, m=total_feature_map)
    f32 = np.float32
    rows = []
    for ar in aspect_ratios:
        root = np.sqrt(f32(ar))
        h, w = f32(s_cur) / root, f32(s_cur) * root
        rows.append([-h / f32(2), -w / f32(2), h / f32(2), w / f32(2)])
    side = np.sqrt(f32(s_cur * s_next))
    rows.append([-side / f32(2), -side / f32(2), side / f32(2), side / f32(2)])
    return _ffi.to_dev(np.asarray(rows, dtype=np.float32))


def generate_prior_boxes(feature_map_shapes: Sequence[int], aspect_ratios: Sequence[Sequence[float]]):
    """utils/bbox_utils.py:179-214 -> ``[N,4]`` float32, clipped to [0,1];
    y-major cells, anchor-minor, maps concatenated in order."""
    _ffi.check_device()
    lib = _ffi.lib()
    fms = _ffi.i32_array([int(f) for f in feature_map_shapes])
    flat, counts = _flatten_aspect_ratios(aspect_ratios)
    if len(counts) != len(feature_map_shapes):
        raise ValueError("feature_map_shapes and aspect_ratios must have the same length")
    ars = _ffi.f32_array(flat)
    cnt = _ffi.i32_array(counts)
    n = lib.ssd_prior_box_count(fms, len(counts), cnt)
    if n <= 0:
        raise ValueError("invalid feature map / aspect ratio specification")
    out = torch.empty((n, 4), dtype=torch.float32, device=_ffi.require_cuda())
    _ffi.check(lib.ssd_prior_boxes(fms, len(counts), ars, cnt, _ffi.ptr(out), n, _ffi.stream()), "ssd_prior_boxes")
    return out


# north_star alias: the name does not exist in the reference snapshot.
init_prior_boxes = generate_prior_boxes


def renormalize_bboxes_with_min_max(bboxes: Any, min_max: Any):
    """utils/bbox_utils.py:217-233: boxes re-expressed inside the window ``[y_min, x_min, y_max, x_max]`` and clipped to
    [0, 1].  (``ssd_augment_batch`` applies the same arithmetic inside the fused augmentation kernels; this stand-alone
    form serves the other call sites with a handful of boxes.)"""
    b = _ffi.to_dev(bboxes)
    mm = _ffi.to_dev(min_max).reshape(4)
    lo = torch.stack([mm[0], mm[1], mm[0], mm[1]])
    span = torch.stack([mm[2] - mm[0], mm[3] - mm[1], mm[2] - mm[0], mm[3] - mm[1]])
    return torch.clamp((b - lo) / span, 0.0, 1.0)


def normalize_bboxes(bboxes: Any, height: Any, width: Any):
    """utils/bbox_utils.py:236-251: pixel corners -> normalised ``[y1, x1, y2, x2]``."""
    b = _ffi.to_dev(bboxes)
    scale = torch.tensor([float(height), float(width), float(height), float(width)], dtype=torch.float32, device=b.device)
    return b / scale


def denormalize_bboxes(bboxes: Any, height: Any, width: Any):
    """utils/bbox_utils.py:254-269: normalised corners -> rounded pixel corners (``tf.round``: half to even)."""
    b = _ffi.to_dev(bboxes)
    scale = torch.tensor([float(height), float(width), float(height), float(width)], dtype=torch.float32, device=b.device)
    return torch.round(b * scale)
