#!/usr/bin/env python
"""Convert the reference's Keras weight file (``trained/ssd_{backbone}_model_weights.h5``, written by
``ModelCheckpoint(save_weights_only=True)`` at trainer.py:108-113, utils/io_utils.py:26-38) into the ``.npz`` this
repository loads (``SSDModel.load_weights``): one array per Keras variable, keyed ``"<layer name>/<variable>"`` in
Keras layouts (kernel HWIO, depthwise_kernel [3,3,C,1], gamma/beta/moving_mean/moving_variance, bias, scale).

Needs ``h5py`` (not present in the build image; run it wherever the .h5 file lives):

    python tools/convert_h5_weights.py trained/ssd_mobilenet_v2_model_weights.h5 trained/ssd_mobilenet_v2_model_weights.npz
"""
import sys

import numpy as np


def convert(src: str, dst: str) -> int:
    try:
        import h5py
    except ImportError as e:                                    # fail loudly: no silent partial conversion
        raise SystemExit("convert_h5_weights.py needs h5py: " + str(e))
    out = {}
    with h5py.File(src, "r") as f:
        root = f["model_weights"] if "model_weights" in f else f

        def visit(name, obj):
            if isinstance(obj, h5py.Dataset):
                parts = name.split("/")
                # Keras stores <layer>/<layer>/<variable>:0 -- keep "<layer>/<variable>"
                var = parts[-1].split(":")[0]
                out[f"{parts[-2]}/{var}"] = np.asarray(obj, dtype=np.float32)
        root.visititems(visit)
    np.savez(dst, **out)
    return len(out)


if __name__ == "__main__":
    if len(sys.argv) != 3:
        raise SystemExit(__doc__)
    print(f"{convert(sys.argv[1], sys.argv[2])} variables written to {sys.argv[2]}")
