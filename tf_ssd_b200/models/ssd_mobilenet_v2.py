"""MobileNetV2-SSD with the reference's entry points (``models/ssd_mobilenet_v2.py``)."""

from __future__ import annotations

from typing import Any, Dict

import numpy as np

from tf_ssd_b200.models.engine import SSDModel


def get_model(hyper_params: Dict[str, Any], seed: int = 0) -> SSDModel:
    """models/ssd_mobilenet_v2.py:15-47.  The Keras ImageNet weights the reference
    downloads are unobtainable offline: variables are seeded random
    (``seed``) until ``load_weights`` is called."""
    return SSDModel("mobilenet_v2", hyper_params, seed=seed)


def init_model(model: SSDModel) -> None:
    """models/ssd_mobilenet_v2.py:50-59 -- one dummy forward; here it also builds
    and warms the batch-1 launch plan."""
    model(np.random.default_rng(0).random((1, 300, 300, 3), dtype=np.float32))
