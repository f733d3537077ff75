"""VGG16-SSD with the reference's entry points (``models/ssd_vgg16.py``)."""

from __future__ import annotations

from typing import Any, Dict

import numpy as np

from tf_ssd_b200.models.engine import SSDModel


def get_model(hyper_params: Dict[str, Any], seed: int = 0) -> SSDModel:
    """models/ssd_vgg16.py:66-121 (L2Normalization :15-63 included).  With the
    seven-map ``vgg16_512`` hyper-parameters the SSD512 extension is built."""
    return SSDModel("vgg16", hyper_params, seed=seed)


def init_model(model: SSDModel) -> None:
    """models/ssd_vgg16.py:124-133.  The reference warms its resolution-agnostic
    graph at 512x512; plans here are per (batch, img_size), so the dummy forward
    uses the configured ``img_size``."""
    s = model.img_size
    model(np.random.default_rng(0).random((1, s, s, 3), dtype=np.float32))
