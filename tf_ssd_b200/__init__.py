"""B200-native (sm_100a) implementation of the tf-ssd hot path.

The package mirrors the module layout of the reference (FurkanOM/tf-ssd):

    tf_ssd_b200.utils.bbox_utils    <- utils/bbox_utils.py
    tf_ssd_b200.utils.train_utils   <- utils/train_utils.py
    tf_ssd_b200.ssd_loss            <- ssd_loss.py
    tf_ssd_b200.models.decoder      <- models/decoder.py
    tf_ssd_b200.models.header       <- models/header.py
    tf_ssd_b200.models.ssd_mobilenet_v2 / ssd_vgg16

Every tensor operation runs in hand-written CUDA kernels reached through the
C ABI of ``libssd_b200.so`` (``include/ssd_b200.h``); torch tensors are device
buffers only.  There is no CPU fallback.
"""

__version__ = "0.1.0"
