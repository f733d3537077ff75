"""keras_applications.mobilenet_v2.MobileNetV2 (1.0.8), alpha = 1.0, include_top=False -- [TF-recall].

The reference calls this third-party constructor at ``models/ssd_mobilenet_v2.py:25``; the package is not under
/root/reference and not installable, so the topology below is a restatement from memory of keras_applications 1.0.8
(and the MobileNetV2 paper's table 2): layer names, ``correct_pad`` zero padding before the stride-2 convolutions,
BatchNormalization(epsilon=1e-3, momentum=0.999), ReLU(6.), residual adds when stride 1 and in == out channels.
Everything the REFERENCE adds on top (tap selection by layer name, extras, heads) runs from the reference's own source.
"""

from __future__ import annotations

from ..layers import Add, BatchNormalization, Conv2D, DepthwiseConv2D, Input, ReLU, ZeroPadding2D
from ..models import Model


def _make_divisible(v, divisor, min_value=None):
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


def correct_pad(size, kernel_size=3):
    """keras_applications.correct_pad for a square input of ``size``: ((k//2 - adjust, k//2), same)."""
    adjust = 1 - size % 2
    correct = kernel_size // 2
    return ((correct - adjust, correct), (correct - adjust, correct))


class _Shape(object):
    def __init__(self, size, channels):
        self.size, self.channels = size, channels


def _inverted_res_block(x, shp, expansion, stride, alpha, filters, block_id):
    in_channels = shp.channels
    pointwise_filters = _make_divisible(int(filters * alpha), 8)
    inputs = x
    prefix = "block_{}_".format(block_id)
    if block_id:
        x = Conv2D(expansion * in_channels, 1, padding="same", use_bias=False, name=prefix + "expand")(x)
        x = BatchNormalization(epsilon=1e-3, momentum=0.999, name=prefix + "expand_BN")(x)
        x = ReLU(6., name=prefix + "expand_relu")(x)
    else:
        prefix = "expanded_conv_"
    if stride == 2:
        x = ZeroPadding2D(padding=correct_pad(shp.size, 3), name=prefix + "pad")(x)
    x = DepthwiseConv2D(3, strides=stride, use_bias=False, padding="same" if stride == 1 else "valid",
                        name=prefix + "depthwise")(x)
    x = BatchNormalization(epsilon=1e-3, momentum=0.999, name=prefix + "depthwise_BN")(x)
    x = ReLU(6., name=prefix + "depthwise_relu")(x)
    x = Conv2D(pointwise_filters, 1, padding="same", use_bias=False, name=prefix + "project")(x)
    x = BatchNormalization(epsilon=1e-3, momentum=0.999, name=prefix + "project_BN")(x)
    if stride == 2:
        p = correct_pad(shp.size, 3)[0]
        shp.size = (shp.size + p[0] + p[1] - 3) // 2 + 1
    shp.channels = pointwise_filters
    if in_channels == pointwise_filters and stride == 1:
        return Add(name=prefix + "add")([inputs, x])
    return x


def MobileNetV2(input_shape=None, alpha=1.0, include_top=True, weights="imagenet", input_tensor=None, pooling=None,
                classes=1000, **kwargs):
    if include_top or alpha != 1.0:
        raise NotImplementedError
    size = int(input_shape[0])
    img_input = Input(shape=input_shape)
    shp = _Shape(size, 3)
    first_block_filters = _make_divisible(32 * alpha, 8)
    x = ZeroPadding2D(padding=correct_pad(shp.size, 3), name="Conv1_pad")(img_input)
    x = Conv2D(first_block_filters, 3, strides=(2, 2), padding="valid", use_bias=False, name="Conv1")(x)
    x = BatchNormalization(epsilon=1e-3, momentum=0.999, name="bn_Conv1")(x)
    x = ReLU(6., name="Conv1_relu")(x)
    p = correct_pad(shp.size, 3)[0]
    shp.size, shp.channels = (shp.size + p[0] + p[1] - 3) // 2 + 1, first_block_filters
    cfg = [(16, 1, 1, 0), (24, 2, 6, 1), (24, 1, 6, 2), (32, 2, 6, 3), (32, 1, 6, 4), (32, 1, 6, 5), (64, 2, 6, 6),
           (64, 1, 6, 7), (64, 1, 6, 8), (64, 1, 6, 9), (96, 1, 6, 10), (96, 1, 6, 11), (96, 1, 6, 12), (160, 2, 6, 13),
           (160, 1, 6, 14), (160, 1, 6, 15), (320, 1, 6, 16)]
    for filters, stride, expansion, block_id in cfg:
        x = _inverted_res_block(x, shp, expansion, stride, alpha, filters, block_id)
    x = Conv2D(1280, 1, use_bias=False, name="Conv_1")(x)
    x = BatchNormalization(epsilon=1e-3, momentum=0.999, name="Conv_1_bn")(x)
    x = ReLU(6., name="out_relu")(x)
    return Model(img_input, x, name="mobilenetv2_1.00_{}".format(size))
