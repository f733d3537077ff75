from . import mobilenet_v2                      # noqa: F401
from .mobilenet_v2 import MobileNetV2            # noqa: F401
