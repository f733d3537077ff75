"""Training configuration and target generation with the reference's names
(``utils/train_utils.py`` of FurkanOM/tf-ssd)."""

from __future__ import annotations

import copy
import math
from typing import Any, Dict, Iterator, Tuple

import torch

from tf_ssd_b200 import _ffi

_AR3 = [1., 2., 1. / 2.]
_AR5 = [1., 2., 1. / 2., 3., 1. / 3.]

# utils/train_utils.py:13-34, plus the SSD512 extension of SURVEY.md Appendix C
# ("vgg16_512" is NOT in the reference; it follows the SSD paper's 7-map layout).
SSD = {
    "vgg16": {
        "img_size": 300,
        "feature_map_shapes": [38, 19, 10, 5, 3, 1],
        "aspect_ratios": [list(_AR3), list(_AR5), list(_AR5), list(_AR5), list(_AR3), list(_AR3)],
    },
    "mobilenet_v2": {
        "img_size": 300,
        "feature_map_shapes": [19, 10, 5, 3, 2, 1],
        "aspect_ratios": [list(_AR3), list(_AR5), list(_AR5), list(_AR5), list(_AR3), list(_AR3)],
    },
    "vgg16_512": {
        "img_size": 512,
        "feature_map_shapes": [64, 32, 16, 8, 4, 2, 1],
        "aspect_ratios": [list(_AR3), list(_AR5), list(_AR5), list(_AR5), list(_AR5), list(_AR3), list(_AR3)],
    },
}


def get_hyper_params(backbone: str, **kwargs: Any) -> Dict[str, Any]:
    """utils/train_utils.py:36-55 -- deep copy + defaults; an override applies
    only when the key already exists AND the value is truthy."""
    hyper_params = copy.deepcopy(SSD[backbone])
    hyper_params["iou_threshold"] = 0.5
    hyper_params["neg_pos_ratio"] = 3
    hyper_params["loc_loss_alpha"] = 1
    hyper_params["variances"] = [0.1, 0.1, 0.2, 0.2]
    for key, value in kwargs.items():
        if key in hyper_params and value:
            hyper_params[key] = value
    return hyper_params


def scheduler(epoch: int) -> float:
    """utils/train_utils.py:57-70."""
    if epoch < 100:
        return 1e-3
    if epoch < 125:
        return 1e-4
    return 1e-5


def get_step_size(total_items: int, batch_size: int) -> int:
    """utils/train_utils.py:72-82."""
    return math.ceil(total_items / batch_size)


def generator(dataset: Any, prior_boxes: Any, hyper_params: Dict[str, Any],
              augmentation_fn: Any = None) -> Iterator[Tuple[Any, Tuple[Any, Any]]]:
    """utils/train_utils.py:84-100 -- infinite ``(img, (deltas, labels))`` feed.  ``augmentation_fn`` (e.g.
    ``tf_ssd_b200.augmentation.apply``) transforms each padded BATCH on the device before the targets are matched: the
    batched form of the per-example ``augmentation_fn`` the reference maps over its ``tf.data`` pipeline (trainer.py:68)."""
    while True:
        for img, gt_boxes, gt_labels in dataset:
            if augmentation_fn is not None:
                img, gt_boxes = augmentation_fn(img, gt_boxes)
            actual_deltas, actual_labels = calculate_actual_outputs(prior_boxes, gt_boxes, gt_labels, hyper_params)
            yield img, (actual_deltas, actual_labels)


def calculate_actual_outputs(prior_boxes: Any, gt_boxes: Any, gt_labels: Any, hyper_params: Dict[str, Any],
                             return_indices: bool = False):
    """utils/train_utils.py:102-136 as ONE fused kernel (IoU + argmax + threshold
    + gather + encode/variances + one-hot); the ``[B,N,G]`` IoU map is never
    written.  Returns ``(bbox_deltas [B,N,4], bbox_labels [B,N,L])``; with
    ``return_indices`` also the int32 label and matched-GT index per anchor."""
    _ffi.check_device()
    total_labels = int(hyper_params["total_labels"])
    iou_threshold = float(hyper_params["iou_threshold"])
    variances = _ffi.f32_array(hyper_params["variances"])
    priors = _ffi.to_dev(prior_boxes)
    gtb = _ffi.to_dev(gt_boxes)
    gtl = _ffi.to_dev(gt_labels, dtype=torch.int32)
    if gtb.dim() != 3 or gtl.dim() != 2 or priors.dim() != 2:
        raise ValueError("expected prior_boxes [N,4], gt_boxes [B,G,4], gt_labels [B,G]")
    B, G = gtb.shape[0], gtb.shape[1]
    N = priors.shape[0]
    dev = priors.device
    if G == 0:      # an all-empty batch: pad one dummy (zero box, label -1) like padded_batch would
        gtb = torch.zeros((B, 1, 4), dtype=torch.float32, device=dev)
        gtl = torch.full((B, 1), -1, dtype=torch.int32, device=dev)
        G = 1
    deltas = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
    onehot = torch.empty((B, N, total_labels), dtype=torch.float32, device=dev)
    lab = torch.empty((B, N), dtype=torch.int32, device=dev) if return_indices else None
    idx = torch.empty((B, N), dtype=torch.int32, device=dev) if return_indices else None
    _ffi.check(_ffi.lib().ssd_match_encode(_ffi.ptr(priors), _ffi.ptr(gtb), _ffi.ptr(gtl), B, N, G, total_labels,
                                           iou_threshold, variances, _ffi.ptr(deltas), _ffi.ptr(onehot),
                                           _ffi.ptr(lab), _ffi.ptr(idx), _ffi.stream()), "ssd_match_encode")
    if return_indices:
        return deltas, onehot, lab, idx
    return deltas, onehot
