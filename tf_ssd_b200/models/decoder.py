"""Decoding layer with the reference's interface (``models/decoder.py``)."""

from __future__ import annotations

from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from tf_ssd_b200 import _ffi


class SSDDecoder(object):
    """models/decoder.py:14-93.  Callable on ``[pred_deltas, pred_label_probs]``;
    returns ``(boxes [B,T,4], labels [B,T] float32, scores [B,T])`` -- note the
    order, it is the reference's (decoder.py:93), not TensorFlow's.

    One fused pipeline on the device: variances, box decode, the background-row
    rule, per-class greedy NMS (IoU 0.5, TensorFlow's default) and the
    cross-class top-``max_total_size`` merge.  ``from_logits=True`` additionally
    fuses the head's softmax (models/header.py:88)."""

    def __init__(self, prior_boxes: Any, variances: Sequence[float], max_total_size: int = 200,
                 score_threshold: float = 0.5, iou_threshold: float = 0.5, **kwargs: Any) -> None:
        self.prior_boxes = prior_boxes
        self.variances = list(variances)
        self.max_total_size = int(max_total_size)
        self.score_threshold = float(score_threshold)
        self.iou_threshold = float(iou_threshold)
        self.name = kwargs.get("name", "ssd_decoder")
        self._priors_dev: Optional[torch.Tensor] = None
        self._ws: Optional[torch.Tensor] = None
        self._ws_key = None
        self.last_valid_detections: Optional[torch.Tensor] = None

    def get_config(self) -> Dict[str, Any]:
        """models/decoder.py:43-58."""
        priors = self.prior_boxes
        if isinstance(priors, torch.Tensor):
            priors = priors.detach().cpu().numpy()
        return {
            "name": self.name,
            "prior_boxes": np.asarray(priors),
            "variances": self.variances,
            "max_total_size": self.max_total_size,
            "score_threshold": self.score_threshold,
        }

    def _priors(self) -> torch.Tensor:
        if self._priors_dev is None or self._priors_dev.device != _ffi.require_cuda():
            self._priors_dev = _ffi.to_dev(self.prior_boxes)
        return self._priors_dev

    def candidate_capacity(self, N: int, L: int, from_logits: bool) -> int:
        """Per-image capacity of the candidate list handed to ``ssd_decode_nms`` (0 = N).  Softmax rows sum to 1, so with
        a threshold >= 0.5 at most one class per anchor can pass (models/decoder.py:78-92) and N is exact.  Below 0.5 a
        normalised row can hold floor(1 / threshold) passing classes; inputs that are not normalised at all may overflow
        any bound short of N * L -- the kernel then reports valid = -1 and ``call`` retries with the full N * L."""
        thr = self.score_threshold
        if from_logits or thr >= 0.5:
            return 0
        return N * L if thr <= 0.0 else min(N * L, N * max(1, int(1.0 / thr)))

    def decode_into(self, pred_deltas: torch.Tensor, pred_labels: torch.Tensor, from_logits: bool,
                    out_boxes: torch.Tensor, out_labels: torch.Tensor, out_scores: torch.Tensor,
                    out_valid: torch.Tensor, max_candidates: Optional[int] = None) -> None:
        """Enqueue the fused decode+NMS on the current stream into caller
        buffers (allocation-free once the workspace exists: graph-capturable)."""
        B, N, L = pred_labels.shape
        lib = _ffi.lib()
        cap = self.candidate_capacity(N, L, from_logits) if max_candidates is None else int(max_candidates)
        key = (B, N, L, self.max_total_size, cap, pred_labels.device)
        if self._ws is None or self._ws_key != key:
            self._ws = _ffi.workspace(lib.ssd_decode_nms_workspace_bytes(B, N, L, self.max_total_size, cap))
            self._ws_key = key
        _ffi.check(lib.ssd_decode_nms(_ffi.ptr(self._priors()), _ffi.ptr(pred_deltas), _ffi.ptr(pred_labels),
                                      B, N, L, _ffi.f32_array(self.variances), int(from_logits),
                                      self.score_threshold, self.iou_threshold, self.max_total_size, cap,
                                      _ffi.ptr(out_boxes), _ffi.ptr(out_labels), _ffi.ptr(out_scores),
                                      _ffi.ptr(out_valid), _ffi.ptr(self._ws), self._ws.numel(), _ffi.stream()),
                   "ssd_decode_nms")

    def call(self, inputs: Sequence[Any], from_logits: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        _ffi.check_device()
        pred_deltas = _ffi.to_dev(inputs[0])
        pred_labels = _ffi.to_dev(inputs[1])
        if pred_deltas.dim() != 3 or pred_labels.dim() != 3 or pred_deltas.shape[2] != 4:
            raise ValueError("expected pred_deltas [B,N,4] and pred_label_probs [B,N,L]")
        B, T, dev = pred_deltas.shape[0], self.max_total_size, pred_deltas.device
        boxes = torch.empty((B, T, 4), dtype=torch.float32, device=dev)
        labels = torch.empty((B, T), dtype=torch.float32, device=dev)
        scores = torch.empty((B, T), dtype=torch.float32, device=dev)
        valid = torch.empty((B,), dtype=torch.int32, device=dev)
        self.decode_into(pred_deltas, pred_labels, from_logits, boxes, labels, scores, valid)
        N, L = pred_labels.shape[1], pred_labels.shape[2]
        if not from_logits and bool((valid < 0).any()):
            # more candidates than the capacity assumed for normalised probabilities (un-normalised scores): TensorFlow's
            # combined NMS has no such limit, so run again with room for every (anchor, class) pair
            self.decode_into(pred_deltas, pred_labels, from_logits, boxes, labels, scores, valid, max_candidates=N * L)
        if bool((valid < 0).any()):
            raise _ffi.SsdB200Error("ssd_decode_nms: candidate list overflow")
        self.last_valid_detections = valid
        return boxes, labels, scores

    __call__ = call


def get_decoder_model(base_model: Any, prior_boxes: Any, hyper_params: Dict[str, Any]):
    """models/decoder.py:96-108 -- wrap an SSD model so that ``predict`` returns
    ``(boxes, labels, scores)``."""
    from tf_ssd_b200.models.engine import DecoderModel
    return DecoderModel(base_model, SSDDecoder(prior_boxes, hyper_params["variances"]))
