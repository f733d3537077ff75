"""tf.keras.layers stand-ins.

A layer called on a symbolic tensor (something that descends from ``Input``) records a graph node; ``Model`` evaluates
the graph eagerly.  A layer called on a concrete tensor runs immediately.  Convolution / pooling geometry follows
the documented TensorFlow rules:

* ``padding="same"``: ``out = ceil(in / stride)``; ``pad_total = max((out - 1) * stride + (k - 1) * dilation + 1 - in, 0)``;
  ``pad_before = pad_total // 2`` (the odd pixel goes AFTER);
* ``padding="valid"``: no padding, ``out = floor((in - ((k - 1) * dilation + 1)) / stride) + 1``;
* max pooling ignores padded positions (they are -inf);
* kernels are HWIO (``[kh, kw, in, out]``), depthwise kernels ``[kh, kw, in, multiplier]``, data is NHWC float32.
"""

from __future__ import annotations

import re
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

import tensorflow as tf

_NAME_COUNTS: Dict[str, int] = {}


def reset_name_counters() -> None:
    _NAME_COUNTS.clear()


def _auto_name(cls_name: str) -> str:
    snake = re.sub(r"(?<=[a-z0-9])([A-Z])", r"_\1", re.sub(r"([A-Z]+)([A-Z][a-z])", r"\1_\2", cls_name)).lower()
    n = _NAME_COUNTS.get(snake, 0)
    _NAME_COUNTS[snake] = n + 1
    return snake if n == 0 else f"{snake}_{n}"


class SymbolicTensor(object):
    """Output of a layer call during graph construction."""

    def __init__(self, layer: Optional["Layer"], inputs: Any, name: str = ""):
        self.layer, self.inputs, self.name = layer, inputs, name


def _is_symbolic(x: Any) -> bool:
    if isinstance(x, SymbolicTensor):
        return True
    if isinstance(x, (list, tuple)):
        return any(_is_symbolic(v) for v in x)
    return False


class Layer(object):
    def __init__(self, trainable: bool = True, name: Optional[str] = None, dtype: Any = None, **kwargs: Any):
        if kwargs:
            raise TypeError(f"unexpected Layer arguments {sorted(kwargs)}")
        self.name = name or _auto_name(type(self).__name__)
        self.trainable = trainable
        self.built = False
        self.output: Optional[SymbolicTensor] = None
        self.input: Any = None

    def get_config(self) -> Dict[str, Any]:
        return {"name": self.name, "trainable": self.trainable}

    def build(self, input_shape: Any) -> None:
        pass

    def call(self, inputs: Any) -> Any:
        raise NotImplementedError

    def __call__(self, inputs: Any, *args: Any, **kwargs: Any) -> Any:
        if _is_symbolic(inputs):
            self.input = inputs
            self.output = SymbolicTensor(self, inputs, self.name)
            return self.output
        return self._run(inputs)

    def _run(self, inputs: Any) -> Any:
        if isinstance(inputs, (list, tuple)):
            inputs = [tf.convert_to_tensor(v) for v in inputs]
        else:
            inputs = tf.convert_to_tensor(inputs)
        if not self.built:
            self.build([t.shape for t in inputs] if isinstance(inputs, list) else inputs.shape)
            self.built = True
        return self.call(inputs)

    # weights in Keras order
    def variables(self) -> Dict[str, tf.Variable]:
        return {}


def Input(shape: Sequence[Optional[int]] = None, batch_size=None, name: Optional[str] = None, dtype=None, **kw):
    return SymbolicTensor(None, None, name or "input")


def _same_pad(size: int, k: int, s: int, d: int):
    out = -(-size // s)
    total = max((out - 1) * s + (k - 1) * d + 1 - size, 0)
    return total // 2, total - total // 2


def _pair(v) -> tuple:
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


def _activation(name: Optional[str], x: np.ndarray) -> np.ndarray:
    if name is None or name == "linear":
        return x
    if name == "relu":
        return np.maximum(x, np.float32(0))
    if name == "softmax":
        return tf.nn.softmax(x).numpy()
    raise NotImplementedError(name)


def _nhwc_conv(x: np.ndarray, w_hwio: np.ndarray, stride, dilation, padding: str, groups: int = 1) -> np.ndarray:
    """NHWC float32 convolution through torch-CPU (float32 accumulate), explicit TF padding."""
    kh, kw = w_hwio.shape[:2]
    if padding == "same":
        ph, pw = _same_pad(x.shape[1], kh, stride[0], dilation[0]), _same_pad(x.shape[2], kw, stride[1], dilation[1])
    elif padding == "valid":
        ph = pw = (0, 0)
    else:
        raise ValueError(padding)
    t = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2)))
    t = F.pad(t, (pw[0], pw[1], ph[0], ph[1]))
    if groups == 1:
        w = torch.from_numpy(np.ascontiguousarray(w_hwio.transpose(3, 2, 0, 1)))          # OIHW
    else:                                                                                  # depthwise [kh,kw,C,1]
        w = torch.from_numpy(np.ascontiguousarray(w_hwio.transpose(2, 3, 0, 1)))          # [C,1,kh,kw]
    y = F.conv2d(t, w, None, stride=stride, dilation=dilation, groups=groups)
    return np.ascontiguousarray(y.numpy().transpose(0, 2, 3, 1))


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", dilation_rate=(1, 1), activation=None,
                 use_bias=True, kernel_initializer="glorot_uniform", bias_initializer="zeros",
                 kernel_regularizer=None, name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.filters, self.kernel_size = int(filters), _pair(kernel_size)
        self.strides, self.dilation_rate = _pair(strides), _pair(dilation_rate)
        self.padding, self.activation, self.use_bias = padding.lower(), activation, use_bias
        self.kernel_initializer, self.kernel_regularizer = kernel_initializer, kernel_regularizer
        self.kernel = self.bias = None

    def build(self, input_shape):
        cin = int(input_shape[-1])
        self.kernel = tf.Variable(np.zeros(self.kernel_size + (cin, self.filters), np.float32), name=self.name + "/kernel")
        if self.use_bias:
            self.bias = tf.Variable(np.zeros((self.filters,), np.float32), name=self.name + "/bias")

    def call(self, inputs):
        y = _nhwc_conv(inputs.numpy(), self.kernel.numpy(), self.strides, self.dilation_rate, self.padding)
        if self.use_bias:
            y = y + self.bias.numpy()
        return tf.Tensor(_activation(self.activation, y))

    def variables(self):
        v = {"kernel": self.kernel}
        if self.use_bias:
            v["bias"] = self.bias
        return v


class DepthwiseConv2D(Layer):
    def __init__(self, kernel_size, strides=(1, 1), padding="valid", depth_multiplier=1, activation=None, use_bias=True,
                 name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.kernel_size, self.strides, self.padding = _pair(kernel_size), _pair(strides), padding.lower()
        self.activation, self.use_bias = activation, use_bias
        if depth_multiplier != 1:
            raise NotImplementedError
        self.depthwise_kernel = self.bias = None

    def build(self, input_shape):
        c = int(input_shape[-1])
        self.depthwise_kernel = tf.Variable(np.zeros(self.kernel_size + (c, 1), np.float32))
        if self.use_bias:
            self.bias = tf.Variable(np.zeros((c,), np.float32))

    def call(self, inputs):
        x = inputs.numpy()
        y = _nhwc_conv(x, self.depthwise_kernel.numpy(), self.strides, (1, 1), self.padding, groups=x.shape[-1])
        if self.use_bias:
            y = y + self.bias.numpy()
        return tf.Tensor(_activation(self.activation, y))

    def variables(self):
        v = {"depthwise_kernel": self.depthwise_kernel}
        if self.use_bias:
            v["bias"] = self.bias
        return v


class BatchNormalization(Layer):
    """Inference mode (documented): ``gamma * (x - moving_mean) / sqrt(moving_variance + epsilon) + beta``."""

    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.epsilon, self.momentum = float(epsilon), float(momentum)

    def build(self, input_shape):
        c = int(input_shape[-1])
        self.gamma, self.beta = tf.Variable(np.ones(c, np.float32)), tf.Variable(np.zeros(c, np.float32))
        self.moving_mean, self.moving_variance = tf.Variable(np.zeros(c, np.float32)), tf.Variable(np.ones(c, np.float32))

    def call(self, inputs):
        x = inputs.numpy()
        inv = (self.gamma.numpy() / np.sqrt(self.moving_variance.numpy() + np.float32(self.epsilon))).astype(np.float32)
        return tf.Tensor(x * inv + (self.beta.numpy() - self.moving_mean.numpy() * inv))

    def variables(self):
        return {"gamma": self.gamma, "beta": self.beta, "moving_mean": self.moving_mean,
                "moving_variance": self.moving_variance}


class ReLU(Layer):
    def __init__(self, max_value=None, name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.max_value = max_value

    def call(self, inputs):
        y = np.maximum(inputs.numpy(), np.float32(0))
        return tf.Tensor(y if self.max_value is None else np.minimum(y, np.float32(self.max_value)))


class ZeroPadding2D(Layer):
    def __init__(self, padding=(1, 1), name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        p = padding
        self.padding = ((p, p), (p, p)) if isinstance(p, int) else tuple((q, q) if isinstance(q, int) else tuple(q) for q in p)

    def call(self, inputs):
        (t, b), (l, r) = self.padding
        return tf.Tensor(np.pad(inputs.numpy(), ((0, 0), (t, b), (l, r), (0, 0))))


class Add(Layer):
    def call(self, inputs):
        out = inputs[0]
        for t in inputs[1:]:
            out = out + t
        return out


class Activation(Layer):
    def __init__(self, activation, name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.activation = activation

    def call(self, inputs):
        return tf.Tensor(_activation(self.activation, inputs.numpy()))


class MaxPool2D(Layer):
    def __init__(self, pool_size=(2, 2), strides=None, padding="valid", name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.pool_size = _pair(pool_size)
        self.strides = _pair(strides if strides is not None else pool_size)
        self.padding = padding.lower()

    def call(self, inputs):
        x = inputs.numpy()
        if self.padding == "same":
            ph, pw = _same_pad(x.shape[1], self.pool_size[0], self.strides[0], 1), \
                _same_pad(x.shape[2], self.pool_size[1], self.strides[1], 1)
        else:
            ph = pw = (0, 0)
        t = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2)))
        t = F.pad(t, (pw[0], pw[1], ph[0], ph[1]), value=float("-inf"))
        y = F.max_pool2d(t, self.pool_size, self.strides)
        return tf.Tensor(np.ascontiguousarray(y.numpy().transpose(0, 2, 3, 1)))


MaxPooling2D = MaxPool2D
