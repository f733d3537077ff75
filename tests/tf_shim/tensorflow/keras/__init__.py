"""tf.keras stand-in: just enough of the functional API for the reference's model files to BUILD and RUN eagerly
(NumPy tensors; convolutions through torch-CPU float32).  See the package docstring of ``tensorflow``."""

from . import layers, models, regularizers, applications, callbacks, optimizers    # noqa: F401
from .models import Model                                                           # noqa: F401
