"""Filesystem / argument helpers with the reference's names (``utils/io_utils.py`` of FurkanOM/tf-ssd).

Weights are stored as ``.npz`` keyed by Keras variable names (``h5py`` is not available in this image; see
``tools/convert_h5_weights.py`` for the ``.h5`` -> ``.npz`` conversion a machine with h5py can run)."""

from __future__ import annotations

import argparse
import os
from datetime import datetime
from typing import Optional, Sequence

VALID_BACKBONES = ("mobilenet_v2", "vgg16")


def get_log_path(model_type: str, custom_postfix: str = "") -> str:
    """utils/io_utils.py:13-24: ``logs/<model><postfix>/<YYYYmmdd-HHMMSS>``."""
    stamp = datetime.now().strftime("%Y%m%d-%H%M%S")
    return os.path.join("logs", f"{model_type}{custom_postfix}", stamp)


def get_model_path(model_type: str, main_path: str = "trained") -> str:
    """utils/io_utils.py:27-39 (``.npz`` instead of ``.h5``); creates the directory."""
    os.makedirs(main_path, exist_ok=True)
    return os.path.join(main_path, f"ssd_{model_type}_model_weights.npz")


def handle_args(argv: Optional[Sequence[str]] = None) -> argparse.Namespace:
    """utils/io_utils.py:42-56: ``-handle-gpu`` and ``--backbone``; the extra options size a run on synthetic
    VOC-shaped data (there is no dataset access in this environment)."""
    parser = argparse.ArgumentParser(description="SSD: Single Shot MultiBox Detector Implementation (B200-native hot path)")
    options = [
        (("-handle-gpu",), dict(action="store_true", help="accepted for compatibility; nothing to do without TensorFlow")),
        (("--backbone",), dict(default=VALID_BACKBONES[0], metavar=str(list(VALID_BACKBONES)), help="Which backbone used for the ssd")),
        (("--epochs",), dict(type=int, default=150)),
        (("--batch-size",), dict(type=int, default=32)),
        (("--train-items",), dict(type=int, default=0, help="synthetic items per epoch (0: dataset default)")),
        (("--val-items",), dict(type=int, default=0)),
        (("--model-dir",), dict(default="trained")),
        (("--no-augmentation",), dict(action="store_true", help="skip augmentation.apply on the training batches")),
    ]
    for flags, kw in options:
        parser.add_argument(*flags, **kw)
    return parser.parse_args(argv)


def is_valid_backbone(backbone: str) -> None:
    """utils/io_utils.py:59-68."""
    assert backbone in VALID_BACKBONES, f"backbone must be one of {VALID_BACKBONES}, got {backbone!r}"


def handle_gpu_compatibility() -> None:
    """utils/io_utils.py:71-83 sets TensorFlow's memory-growth flag; there is no TensorFlow here."""
    return None
