#!/usr/bin/env python
"""Convert the reference's Keras weight file (``trained/ssd_{backbone}_model_weights.h5``, written by
``ModelCheckpoint(save_weights_only=True)`` at trainer.py:108-113, utils/io_utils.py:26-38) into the ``.npz`` this
repository loads (``SSDModel.load_weights``): one array per Keras variable, keyed ``"<layer name>/<variable>"`` in
Keras layouts (kernel HWIO, depthwise_kernel [3,3,C,1], gamma/beta/moving_mean/moving_variance, bias, scale).

Keras stores ``<layer>/<layer>/<variable>:0`` (optionally under a ``model_weights`` group).  The reference's
``L2Normalization`` creates its per-channel scale with an UNNAMED ``tf.Variable`` (models/ssd_vgg16.py:52), which Keras
saves as ``Variable:0``: a layer group holding a single variable that is none of the standard Keras names is mapped to
``<layer>/scale`` (the name ``SSDModel`` uses).

Needs ``h5py`` for real files (not present in the build image; run it wherever the .h5 file lives):

    python tools/convert_h5_weights.py trained/ssd_mobilenet_v2_model_weights.h5 trained/ssd_mobilenet_v2_model_weights.npz

``convert_tree`` works on any nested mapping with the same shape (the tests feed it a hand-built tree).
"""
import sys
from typing import Any, Dict, Mapping

import numpy as np

KERAS_VARIABLES = ("kernel", "depthwise_kernel", "bias", "gamma", "beta", "moving_mean", "moving_variance", "scale")


def _is_group(obj: Any) -> bool:
    return isinstance(obj, Mapping) or (hasattr(obj, "keys") and hasattr(obj, "__getitem__") and not hasattr(obj, "shape"))


def convert_tree(root: Any) -> Dict[str, np.ndarray]:
    """``{"<layer>/<variable>": float32 array}`` from a Keras weight tree (h5py group or nested mapping)."""
    if _is_group(root) and "model_weights" in root.keys():
        root = root["model_weights"]
    found = {}                                          # layer -> {variable: array}

    def visit(group: Any, path: tuple) -> None:
        for key in group.keys():
            obj = group[key]
            if _is_group(obj):
                visit(obj, path + (key,))
            else:
                if not path:
                    raise ValueError(f"dataset {key!r} is not inside a layer group")
                var = key.split(":")[0]
                found.setdefault(path[-1], {})[var] = np.asarray(obj, dtype=np.float32)

    visit(root, ())
    out = {}
    for layer, variables in found.items():
        unknown = [v for v in variables if v not in KERAS_VARIABLES]
        if unknown:
            if len(variables) == 1:                     # e.g. l2_normalization/Variable:0 -> l2_normalization/scale
                out[f"{layer}/scale"] = variables[unknown[0]]
                continue
            raise ValueError(f"layer {layer!r} holds variables this converter does not know: {unknown}")
        for var, arr in variables.items():
            out[f"{layer}/{var}"] = arr
    return out


def convert(src: str, dst: str) -> int:
    try:
        import h5py
    except ImportError as e:                                    # fail loudly: no silent partial conversion
        raise SystemExit("convert_h5_weights.py needs h5py: " + str(e))
    with h5py.File(src, "r") as f:
        out = convert_tree(f)
    np.savez(dst, **out)
    return len(out)


if __name__ == "__main__":
    if len(sys.argv) != 3:
        raise SystemExit(__doc__)
    print(f"{convert(sys.argv[1], sys.argv[2])} variables written to {sys.argv[2]}")
