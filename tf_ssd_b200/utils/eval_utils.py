"""PASCAL-VOC style evaluation with the reference's names (``utils/eval_utils.py`` of FurkanOM/tf-ssd).

``update_stats`` is the natural consumer of the IoU kernel in its ``[B,M,4] x [B,G,4] -> [B,M,G]`` mode
(utils/eval_utils.py:57): the map is computed on the device by ``ssd_iou_map``; the greedy
one-prediction-per-ground-truth bookkeeping (a Python triple loop in the reference) runs on the host over the
``[B,200]`` reductions."""

from __future__ import annotations

from typing import Any, Dict, Sequence, Tuple

import numpy as np

from tf_ssd_b200.utils import bbox_utils


def init_stats(labels: Sequence[str]) -> Dict[int, Dict[str, Any]]:
    """utils/eval_utils.py:12-33."""
    return {i: {"label": lab, "total": 0, "tp": [], "fp": [], "scores": []} for i, lab in enumerate(labels) if i != 0}


def _np(x) -> np.ndarray:
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def update_stats(pred_bboxes: Any, pred_labels: Any, pred_scores: Any, gt_boxes: Any, gt_labels: Any,
                 stats: Dict[int, Dict[str, Any]]) -> Dict[int, Dict[str, Any]]:
    """utils/eval_utils.py:36-91."""
    iou_map = _np(bbox_utils.generate_iou_map(pred_bboxes, gt_boxes))             # [B,M,G] on the device (:57)
    pred_labels, pred_scores, gt_labels = _np(pred_labels), _np(pred_scores), _np(gt_labels)
    merged = iou_map.max(-1)                                                      # :58
    gt_of = iou_map.argmax(-1).astype(np.int32)                                   # :59 (first maximum)
    order = np.argsort(-merged, axis=-1, kind="stable")                           # :60 descending, ties by index
    uniq, counts = np.unique(gt_labels.reshape(-1), return_counts=True)           # :62-67
    for lab, c in zip(uniq, counts):
        if lab != -1:
            stats[int(lab)]["total"] += int(c)
    for b in range(merged.shape[0]):                                              # :68-90
        taken = set()
        for m in order[b]:
            lab = int(pred_labels[b, m])
            if lab == 0:
                continue
            g = int(gt_of[b, m])
            st = stats[lab]
            st["scores"].append(float(pred_scores[b, m]))
            hit = merged[b, m] >= 0.5 and lab == int(gt_labels[b, g]) and g not in taken
            st["tp"].append(1 if hit else 0)
            st["fp"].append(0 if hit else 1)
            if hit:
                taken.add(g)
    return stats


def calculate_ap(recall: np.ndarray, precision: np.ndarray) -> float:
    """utils/eval_utils.py:94-108: 11-point interpolated AP."""
    ap = 0.0
    for thr in np.arange(0, 1.1, 0.1):
        p = precision[recall >= thr]
        if len(p) > 0:
            ap += float(np.amax(p))
    return ap / 11


def calculate_mAP(stats: Dict[int, Dict[str, Any]]) -> Tuple[Dict[int, Dict[str, Any]], float]:
    """utils/eval_utils.py:111-139."""
    aps = []
    for label, st in stats.items():
        tp, fp, scores = np.array(st["tp"]), np.array(st["fp"]), np.array(st["scores"])
        ids = np.argsort(-scores, kind="stable")
        acc_tp, acc_fp = np.cumsum(tp[ids]), np.cumsum(fp[ids])
        with np.errstate(divide="ignore", invalid="ignore"):
            recall = acc_tp / st["total"] if st["total"] else np.zeros_like(acc_tp, dtype=np.float64)
            precision = acc_tp / np.maximum(acc_fp + acc_tp, 1)
        ap = calculate_ap(recall, precision)
        st["recall"], st["precision"], st["AP"] = recall, precision, ap
        aps.append(ap)
    return stats, float(np.mean(aps)) if aps else 0.0


def evaluate_predictions(dataset: Any, pred_bboxes: Any, pred_labels: Any, pred_scores: Any, labels: Sequence[str],
                         batch_size: int) -> Dict[int, Dict[str, Any]]:
    """utils/eval_utils.py:142-178."""
    stats = init_stats(labels)
    for batch_id, (_, gt_boxes, gt_labels) in enumerate(dataset):
        start, end = batch_id * batch_size, (batch_id + 1) * batch_size
        if start >= len(pred_bboxes):
            break
        stats = update_stats(pred_bboxes[start:end], pred_labels[start:end], pred_scores[start:end], gt_boxes, gt_labels, stats)
    stats, m_ap = calculate_mAP(stats)
    print("mAP: {}".format(float(m_ap)))
    return stats
