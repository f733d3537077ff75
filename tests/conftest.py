import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# the reference's hyper-parameters, restated for tests (utils/train_utils.py:13-34)
AR3 = [1., 2., 1. / 2.]
AR5 = [1., 2., 1. / 2., 3., 1. / 3.]
CONFIGS = {
    "mobilenet_v2": ([19, 10, 5, 3, 2, 1], [AR3, AR5, AR5, AR5, AR3, AR3], 2268),
    "vgg16": ([38, 19, 10, 5, 3, 1], [AR3, AR5, AR5, AR5, AR3, AR3], 8732),
    "vgg16_512": ([64, 32, 16, 8, 4, 2, 1], [AR3, AR5, AR5, AR5, AR5, AR3, AR3], 24564),
}
VARIANCES = [0.1, 0.1, 0.2, 0.2]
