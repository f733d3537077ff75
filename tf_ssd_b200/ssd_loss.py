"""SSD loss with the reference's interface (``ssd_loss.py`` of FurkanOM/tf-ssd)."""

from __future__ import annotations

from typing import Any, Optional, Tuple

import torch

from tf_ssd_b200 import _ffi


class CustomLoss(object):
    """ssd_loss.py:10-91 -- ``loc_loss_fn`` / ``conf_loss_fn`` return the
    per-image ``[B]`` losses exactly like the reference's Keras loss callables."""

    def __init__(self, neg_pos_ratio: int, loc_loss_alpha: float) -> None:
        self.neg_pos_ratio = float(neg_pos_ratio)
        self.loc_loss_alpha = float(loc_loss_alpha)
        self._ws: Optional[torch.Tensor] = None

    def _workspace(self, B: int, N: int, L: int) -> torch.Tensor:
        need = _ffi.lib().ssd_loss_workspace_bytes(B, N, L)
        if self._ws is None or self._ws.numel() < need or self._ws.device != _ffi.require_cuda():
            self._ws = _ffi.workspace(need)
        return self._ws

    def _run(self, ad, pd, al, pl, from_logits: bool) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        _ffi.check_device()
        ref = ad if ad is not None else al
        B, N = ref.shape[0], ref.shape[1]
        L = al.shape[2] if al is not None else 1
        dev = ref.device
        loc = torch.empty((B,), dtype=torch.float32, device=dev) if ad is not None else None
        conf = torch.empty((B,), dtype=torch.float32, device=dev) if al is not None else None
        ws = self._workspace(B, N, L)
        _ffi.check(_ffi.lib().ssd_loss_fwd(_ffi.ptr(ad), _ffi.ptr(pd), _ffi.ptr(al), _ffi.ptr(pl), B, N, L,
                                           self.neg_pos_ratio, self.loc_loss_alpha, int(from_logits),
                                           _ffi.ptr(loc), _ffi.ptr(conf), _ffi.ptr(ws), ws.numel(), _ffi.stream()),
                   "ssd_loss_fwd")
        return loc, conf

    def loc_loss_fn(self, actual_deltas: Any, pred_deltas: Any):
        """ssd_loss.py:26-57."""
        ad, pd = _ffi.to_dev(actual_deltas), _ffi.to_dev(pred_deltas)
        if ad.shape != pd.shape or ad.dim() != 3 or ad.shape[2] != 4:
            raise ValueError("expected actual_deltas and pred_deltas of shape [B,N,4]")
        return self._run(ad, pd, None, None, False)[0]

    def conf_loss_fn(self, actual_labels: Any, pred_labels: Any, from_logits: bool = False):
        """ssd_loss.py:59-91.  ``pred_labels`` are probabilities (the public
        signature: Keras renormalise+clip path); ``from_logits=True`` evaluates
        what Keras substitutes inside ``model.fit`` (softmax CE on the logits)."""
        al, pl = _ffi.to_dev(actual_labels), _ffi.to_dev(pred_labels)
        if al.shape != pl.shape or al.dim() != 3:
            raise ValueError("expected actual_labels and pred_labels of shape [B,N,L]")
        return self._run(None, None, al, pl, from_logits)[1]

    # -- fused training entry: both losses + gradients in three kernels -----
    def forward_backward(self, actual_deltas, pred_deltas, actual_labels, pred_logits, grad_scale: Optional[float] = None):
        """Both per-image losses (logits path) and the gradients of
        ``grad_scale * (sum loc + sum conf)`` w.r.t. ``pred_deltas`` and the
        logits.  ``grad_scale`` defaults to ``1/B`` (Keras batch mean,
        trainer.py:91-94)."""
        ad, pd = _ffi.to_dev(actual_deltas), _ffi.to_dev(pred_deltas)
        al, pl = _ffi.to_dev(actual_labels), _ffi.to_dev(pred_logits)
        loc, conf = self._run(ad, pd, al, pl, True)
        B, N, L = al.shape
        gs = float(grad_scale) if grad_scale is not None else 1.0 / B
        gd = torch.empty_like(pd)
        gz = torch.empty_like(pl)
        ws = self._ws
        _ffi.check(_ffi.lib().ssd_loss_bwd(_ffi.ptr(ad), _ffi.ptr(pd), _ffi.ptr(al), _ffi.ptr(pl), B, N, L,
                                           self.loc_loss_alpha, gs, _ffi.ptr(gd), _ffi.ptr(gz), _ffi.ptr(ws),
                                           ws.numel(), _ffi.stream()), "ssd_loss_bwd")
        return loc, conf, gd, gz


def ssd_loss(actual_deltas, pred_deltas, actual_labels, pred_labels, neg_pos_ratio: int = 3,
             loc_loss_alpha: float = 1.0, from_logits: bool = False):
    """north_star alias: per-image ``loc + conf`` in one call."""
    fn = CustomLoss(neg_pos_ratio, loc_loss_alpha)
    ad, pd = _ffi.to_dev(actual_deltas), _ffi.to_dev(pred_deltas)
    al, pl = _ffi.to_dev(actual_labels), _ffi.to_dev(pred_labels)
    loc, conf = fn._run(ad, pd, al, pl, from_logits)
    return loc + conf
