"""Multibox head with the reference's names (``models/header.py``).

In the reference ``get_head_from_outputs`` adds twelve Keras Conv2D layers, two
``HeadWrapper`` reshape+concat layers and a softmax to a functional graph.
Here the head is part of the launch plan (``engine._PlanBuilder.head``): one 3x3
convolution per feature map computes the label and box channels together and
writes them at the map's anchor offset of the concatenated outputs, so the
reshape/concat costs nothing."""

from __future__ import annotations

from typing import Any, Dict, Sequence, Tuple

import torch

from tf_ssd_b200 import _ffi


class HeadWrapper(object):
    """models/header.py:11-51 -- reshape ``[B,H,W,A*C] -> [B,H*W*A,C]`` and concat
    along axis 1.  Pure view/concat of device buffers (no arithmetic)."""

    def __init__(self, last_dimension: int, **kwargs: Any) -> None:
        self.last_dimension = int(last_dimension)
        self.name = kwargs.get("name", "head_wrapper")

    def get_config(self) -> Dict[str, Any]:
        return {"name": self.name, "last_dimension": self.last_dimension}

    def call(self, inputs: Sequence[torch.Tensor]) -> torch.Tensor:
        batch = inputs[0].shape[0]
        return torch.cat([x.reshape(batch, -1, self.last_dimension) for x in inputs], dim=1)

    __call__ = call


def softmax(logits: torch.Tensor) -> torch.Tensor:
    """models/header.py:88 ``Activation("softmax")`` over the last axis."""
    _ffi.check_device()
    z = _ffi.to_dev(logits)
    out = torch.empty_like(z)
    rows = z.numel() // z.shape[-1]
    _ffi.check(_ffi.lib().ssd_softmax(_ffi.ptr(z), rows, z.shape[-1], _ffi.ptr(out), _ffi.stream()), "ssd_softmax")
    return out


def get_head_from_outputs(hyper_params: Dict[str, Any], outputs: Sequence[Any]) -> Tuple[Any, Any]:
    """models/header.py:54-90.  ``outputs`` are the tap activations of a plan under
    construction (``engine.Act``); returns ``(pred_deltas, pred_logits)`` buffers."""
    builder = getattr(outputs, "builder", None)
    if builder is None:
        raise TypeError("get_head_from_outputs is driven by the launch-plan builder (engine.SSDModel.plan)")
    return builder.head(list(outputs), hyper_params)
