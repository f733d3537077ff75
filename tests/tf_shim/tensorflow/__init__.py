"""A NumPy-backed stand-in for the ``tensorflow`` module -- TEST INFRASTRUCTURE ONLY.

Purpose: TensorFlow 2.0 cannot be installed in this environment (no wheel, no
network), and the reference's own unit tests stub ``tensorflow`` out without
executing any tensor op (``/root/reference/tests/test_support.py:11-68``).  This
package makes the UNMODIFIED reference sources executable: with this directory
first on ``sys.path``, ``/root/reference/utils/bbox_utils.py``,
``utils/train_utils.py``, ``ssd_loss.py``, ``models/decoder.py``,
``models/header.py``, ``models/ssd_vgg16.py`` and ``models/ssd_mobilenet_v2.py``
import and run as written, and ``tests/golden/make_ref_golden.py`` records what
they compute as fixtures (``tests/golden/ref_*.npz``).  The oracle (``oracle/``)
and the CUDA path are then both held to bytes produced by the reference's own
code.

Rules of this package:

* it never imports ``oracle/`` or ``tf_ssd_b200`` (it is an independent witness);
* every op is written from TensorFlow's documented behaviour: eager execution,
  float32 stays float32, Python scalars take the dtype of the tensor operand,
  ``int / int`` is a float64 true division, binary ops on tensors of different
  dtypes raise (as TensorFlow does);
* what the documentation does NOT pin down and comes from memory of the TF 2.0
  sources is tagged ``[TF-recall]`` where it is implemented.  These are: tie order
  of ``argmax`` / ``top_k`` (lowest index first), and inside
  ``image.combined_non_max_suppression`` the order in which equal scores are
  visited.  Fixtures avoid depending on the second one (see make_ref_golden.py).

Arithmetic is NumPy's IEEE float32 (one rounding per elementary op, like an eager
TF op).  ``exp`` / ``log`` may differ from Eigen's in the last ulp.
"""

from __future__ import annotations

import builtins
import types
from typing import Any, Callable, Optional, Sequence

import numpy as np

__version__ = "2.0.0-numpy-shim"

float16 = np.dtype("float16")
float32 = np.dtype("float32")
float64 = np.dtype("float64")
int32 = np.dtype("int32")
int64 = np.dtype("int64")
uint8 = np.dtype("uint8")
bool = np.dtype("bool")      # noqa: A001  (tf.bool)

_pybool, _pyint, _pyfloat = builtins.bool, builtins.int, builtins.float


def _np_dtype(dtype: Any) -> Optional[np.dtype]:
    if dtype is None:
        return None
    if isinstance(dtype, str):
        return np.dtype(dtype)
    return np.dtype(dtype)


class Tensor(object):
    """An eager tensor: a NumPy array with TensorFlow's operator rules."""

    __array_priority__ = 1000.0      # ndarray <op> Tensor defers to the Tensor's reflected method

    def __init__(self, value: np.ndarray):
        self._v = np.asarray(value)

    # -- introspection -------------------------------------------------------
    def numpy(self) -> np.ndarray:
        return self._v

    @property
    def dtype(self) -> np.dtype:
        return self._v.dtype

    @property
    def shape(self):
        return tuple(self._v.shape)

    @property
    def ndim(self) -> int:
        return self._v.ndim

    def __len__(self) -> int:
        return len(self._v)

    def __array__(self, dtype=None, copy=None):
        return self._v if dtype is None else self._v.astype(dtype)

    def __repr__(self) -> str:
        return f"<shim.Tensor shape={self.shape} dtype={self.dtype} numpy={self._v!r}>"

    def __bool__(self) -> builtins.bool:
        return _pybool(self._v)

    def __int__(self) -> builtins.int:
        return _pyint(self._v)

    def __float__(self) -> builtins.float:
        return _pyfloat(self._v)

    def __index__(self) -> builtins.int:
        return _pyint(self._v)

    def __hash__(self):
        return id(self)

    def __iter__(self):
        for i in builtins.range(self._v.shape[0]):
            yield Tensor(self._v[i])

    def __getitem__(self, idx):
        def conv(i):
            return _pyint(i._v) if isinstance(i, Tensor) and i._v.ndim == 0 else (i._v if isinstance(i, Tensor) else i)
        if isinstance(idx, tuple):
            idx = tuple(conv(i) for i in idx)
        else:
            idx = conv(idx)
        return Tensor(self._v[idx])

    # -- arithmetic (math_ops.py operator overloads) ---------------------------
    def __neg__(self):
        return Tensor(-self._v)

    def __abs__(self):
        return Tensor(np.abs(self._v))

    def __add__(self, o):
        return add(self, o)

    def __radd__(self, o):
        return add(o, self)

    def __sub__(self, o):
        return subtract(self, o)

    def __rsub__(self, o):
        return subtract(o, self)

    def __mul__(self, o):
        return multiply(self, o)

    def __rmul__(self, o):
        return multiply(o, self)

    def __truediv__(self, o):
        return truediv(self, o)

    def __rtruediv__(self, o):
        return truediv(o, self)

    def __floordiv__(self, o):
        a, b = _pair(self, o)
        return Tensor(np.floor_divide(a, b))

    def __lt__(self, o):
        return less(self, o)

    def __le__(self, o):
        return less_equal(self, o)

    def __gt__(self, o):
        return greater(self, o)

    def __ge__(self, o):
        return greater_equal(self, o)

    def __eq__(self, o):          # TF2 behaviour: element-wise
        return equal(self, o)

    def __ne__(self, o):
        return not_equal(self, o)

    def __and__(self, o):
        return logical_and(self, o)

    def __or__(self, o):
        return logical_or(self, o)

    def __invert__(self):
        return logical_not(self)


class Variable(Tensor):
    """tf.Variable: a mutable tensor (ssd_vgg16.py:52 creates the L2-norm scale with it)."""

    def __init__(self, initial_value, trainable: builtins.bool = True, name: Optional[str] = None, dtype=None):
        super().__init__(np.array(convert_to_tensor(initial_value, dtype=dtype)._v))
        self.trainable = trainable
        self.name = name or "Variable"

    def assign(self, value):
        v = convert_to_tensor(value, dtype=self.dtype)._v
        if v.shape != self._v.shape:
            raise ValueError(f"assign: shape {v.shape} != {self._v.shape}")
        self._v = np.array(v)
        return self


def _infer(value: Any) -> np.ndarray:
    """ops.convert_to_tensor without a dtype: Python floats -> float32, ints -> int32, NumPy data keeps its dtype."""
    if isinstance(value, np.ndarray):
        return value
    if isinstance(value, np.generic):
        return np.asarray(value)
    if isinstance(value, _pybool):
        return np.asarray(value, dtype=np.bool_)
    if isinstance(value, _pyint):
        return np.asarray(value, dtype=np.int32)
    if isinstance(value, _pyfloat):
        return np.asarray(value, dtype=np.float32)
    if isinstance(value, (list, tuple)):
        if _contains_tensor(value):                      # auto-packing of nested lists of tensors (ops.pack)
            parts = [convert_to_tensor(v)._v for v in value]
            dts = {p.dtype for p in parts}
            if len(dts) != 1:
                raise TypeError(f"cannot pack tensors of dtypes {dts}")
            return np.stack(parts)
        a = np.asarray(value)
        if a.dtype == np.float64:
            return a.astype(np.float32)
        if a.dtype == np.int64:
            return a.astype(np.int32)
        return a
    raise TypeError(f"cannot convert {type(value)} to a tensor")


def _contains_tensor(value: Any) -> builtins.bool:
    if isinstance(value, Tensor):
        return True
    if isinstance(value, (list, tuple)):
        return any(_contains_tensor(v) for v in value)
    return False


def convert_to_tensor(value: Any, dtype: Any = None, dtype_hint: Any = None) -> Tensor:
    dtype = _np_dtype(dtype)
    if isinstance(value, Tensor):
        if dtype is not None and value.dtype != dtype:
            raise TypeError(f"Tensor conversion requested dtype {dtype} for a tensor of dtype {value.dtype}")
        return value
    if dtype is None and dtype_hint is not None and not isinstance(value, (np.ndarray, np.generic)) \
            and not _contains_tensor(value):
        hint = _np_dtype(dtype_hint)
        # a Python scalar / list takes the hinted dtype when it is representable in it
        a = np.asarray(value)
        if a.dtype.kind == "f" and hint.kind != "f":
            raise TypeError(f"cannot convert a float to {hint}")
        return Tensor(a.astype(hint))
    a = _infer(value)
    if dtype is not None and a.dtype != dtype:
        if isinstance(value, (np.ndarray, np.generic)) or _contains_tensor(value):
            raise TypeError(f"expected {dtype}, got {a.dtype}")
        a = np.asarray(value).astype(dtype)
    return Tensor(a)


def constant(value: Any, dtype: Any = None, shape: Any = None, name: Optional[str] = None) -> Tensor:
    dtype = _np_dtype(dtype)
    if isinstance(value, Tensor):
        value = value._v
    a = _infer(value) if dtype is None else np.asarray(value, dtype=dtype)
    if shape is not None:
        a = np.broadcast_to(a, shape).copy()
    return Tensor(a)


def _pair(x: Any, y: Any):
    """Operand pair of a binary op: the non-tensor side takes the tensor side's dtype; two tensors must agree."""
    xt, yt = isinstance(x, Tensor), isinstance(y, Tensor)
    if xt and not yt:
        y = convert_to_tensor(y, dtype_hint=x.dtype)
    elif yt and not xt:
        x = convert_to_tensor(x, dtype_hint=y.dtype)
    elif not xt and not yt:
        x = convert_to_tensor(x)
        y = convert_to_tensor(y, dtype_hint=x.dtype)
    if x.dtype != y.dtype:
        raise TypeError(f"binary op on dtypes {x.dtype} and {y.dtype} (TensorFlow does not promote)")
    return x._v, y._v


def _binary(fn: Callable) -> Callable:
    def op(x, y, name=None):
        a, b = _pair(x, y)
        with np.errstate(all="ignore"):
            return Tensor(fn(a, b))
    return op


add = _binary(np.add)
subtract = _binary(np.subtract)
multiply = _binary(np.multiply)
maximum = _binary(np.maximum)
minimum = _binary(np.minimum)
less = _binary(np.less)
less_equal = _binary(np.less_equal)
greater = _binary(np.greater)
greater_equal = _binary(np.greater_equal)
equal = _binary(np.equal)
not_equal = _binary(np.not_equal)
logical_and = _binary(np.logical_and)
logical_or = _binary(np.logical_or)


def truediv(x, y, name=None):
    """math_ops.truediv: integer operands are cast to float64 (int32) before dividing; floats divide in their dtype."""
    a, b = _pair(x, y)
    if a.dtype.kind in "iu":
        to = np.float32 if a.dtype.itemsize <= 2 else np.float64
        a, b = a.astype(to), b.astype(to)
    with np.errstate(all="ignore"):
        return Tensor(np.true_divide(a, b))


divide = truediv


def logical_not(x, name=None):
    return Tensor(np.logical_not(convert_to_tensor(x)._v))


def _unary(fn: Callable) -> Callable:
    def op(x, name=None):
        with np.errstate(all="ignore"):
            return Tensor(fn(convert_to_tensor(x)._v))
    return op


sqrt = _unary(np.sqrt)
exp = _unary(np.exp)
abs = _unary(np.abs)          # noqa: A001
square = _unary(np.square)
floor = _unary(np.floor)
negative = _unary(np.negative)
zeros_like = _unary(np.zeros_like)
ones_like = _unary(np.ones_like)
identity = _unary(lambda a: a)


def round(x, name=None):      # noqa: A001   tf.round: half to even
    return Tensor(np.round(convert_to_tensor(x)._v))


def _log(x, name=None):
    with np.errstate(all="ignore"):
        return Tensor(np.log(convert_to_tensor(x)._v))


def cast(x, dtype, name=None) -> Tensor:
    """tf.cast; float -> int truncates toward zero."""
    dtype = _np_dtype(dtype)
    a = convert_to_tensor(x)._v
    if a.dtype.kind == "f" and dtype.kind in "iu":
        with np.errstate(all="ignore"):
            return Tensor(np.trunc(a).astype(dtype))
    return Tensor(a.astype(dtype))


def rank(x, name=None) -> Tensor:
    return Tensor(np.asarray(convert_to_tensor(x)._v.ndim, dtype=np.int32))


def shape(x, out_type=int32, name=None) -> Tensor:
    return Tensor(np.asarray(convert_to_tensor(x)._v.shape, dtype=_np_dtype(out_type)))


def size(x, name=None) -> Tensor:
    return Tensor(np.asarray(convert_to_tensor(x)._v.size, dtype=np.int32))


def _axes(axis):
    if axis is None:
        return None
    if isinstance(axis, Tensor):
        axis = axis._v.tolist()
    if isinstance(axis, (list, tuple)):
        return tuple(_pyint(a) for a in axis)
    return _pyint(axis)


def _static_shape(shp) -> tuple:
    if isinstance(shp, Tensor):
        return tuple(_pyint(s) for s in shp._v.tolist())
    return tuple(_pyint(s._v) if isinstance(s, Tensor) else _pyint(s) for s in shp)


def reshape(x, shape, name=None) -> Tensor:      # noqa: A002
    return Tensor(convert_to_tensor(x)._v.reshape(_static_shape(shape)))


def split(value, num_or_size_splits, axis=0, num=None, name=None):
    a = convert_to_tensor(value)._v
    if isinstance(num_or_size_splits, _pyint):
        return [Tensor(p) for p in np.split(a, num_or_size_splits, axis=_axes(axis))]
    idx = np.cumsum(list(num_or_size_splits))[:-1]
    return [Tensor(p) for p in np.split(a, idx, axis=_axes(axis))]


def squeeze(x, axis=None, name=None) -> Tensor:
    return Tensor(np.squeeze(convert_to_tensor(x)._v, axis=_axes(axis)))


def expand_dims(x, axis, name=None) -> Tensor:
    return Tensor(np.expand_dims(convert_to_tensor(x)._v, _axes(axis)))


def transpose(x, perm=None, name=None) -> Tensor:
    return Tensor(np.transpose(convert_to_tensor(x)._v, None if perm is None else _static_shape(perm)))


def stack(values, axis=0, name=None) -> Tensor:
    parts = [convert_to_tensor(v) for v in values]
    if len({p.dtype for p in parts}) != 1:
        raise TypeError("stack: mixed dtypes")
    return Tensor(np.stack([p._v for p in parts], axis=_axes(axis)))


def concat(values, axis, name=None) -> Tensor:
    parts = [convert_to_tensor(v) for v in values]
    if len({p.dtype for p in parts}) != 1:
        raise TypeError("concat: mixed dtypes")
    return Tensor(np.concatenate([p._v for p in parts], axis=_axes(axis)))


def where(condition, x=None, y=None, name=None) -> Tensor:
    """tf.where (v2): broadcasting select."""
    c = convert_to_tensor(condition)._v
    if c.dtype != np.bool_:
        raise TypeError("where: condition must be bool")
    if x is None and y is None:
        return Tensor(np.argwhere(c).astype(np.int64))
    a, b = _pair(x, y)
    return Tensor(np.where(c, a, b))


def clip_by_value(t, clip_value_min, clip_value_max, name=None) -> Tensor:
    """clip_ops.clip_by_value: minimum(t, max) then maximum(., min)."""
    t = convert_to_tensor(t)
    return maximum(minimum(t, clip_value_max), clip_value_min)


def range(start, limit=None, delta=1, dtype=None, name=None) -> Tensor:      # noqa: A001
    vals = [start] + ([limit] if limit is not None else []) + [delta]
    vals = [_pyfloat(v._v) if isinstance(v, Tensor) and v.dtype.kind == "f" else (_pyint(v._v) if isinstance(v, Tensor) else v)
            for v in vals]
    if dtype is None:
        dtype = np.float32 if any(isinstance(v, _pyfloat) for v in vals) else np.int32
    if limit is None:
        return Tensor(np.arange(0, vals[0], vals[1], dtype=_np_dtype(dtype)))
    return Tensor(np.arange(vals[0], vals[1], vals[2], dtype=_np_dtype(dtype)))


def meshgrid(*args, indexing="xy", name=None):
    return [Tensor(np.array(g)) for g in np.meshgrid(*[convert_to_tensor(a)._v for a in args], indexing=indexing)]


def fill(dims, value, name=None) -> Tensor:
    return Tensor(np.full(_static_shape(dims), _infer(value)))


def zeros(shape, dtype=float32, name=None) -> Tensor:      # noqa: A002
    return Tensor(np.zeros(_static_shape(shape), _np_dtype(dtype)))


def ones(shape, dtype=float32, name=None) -> Tensor:       # noqa: A002
    return Tensor(np.ones(_static_shape(shape), _np_dtype(dtype)))


def _reduce(fn: Callable) -> Callable:
    def op(x, axis=None, keepdims=False, name=None):
        with np.errstate(all="ignore"):
            return Tensor(np.asarray(fn(convert_to_tensor(x)._v, axis=_axes(axis), keepdims=keepdims)))
    return op


def _sum_keep_dtype(a, axis=None, keepdims=False):
    return np.sum(a, axis=axis, keepdims=keepdims, dtype=a.dtype)


def _mean_keep_dtype(a, axis=None, keepdims=False):
    return np.mean(a, axis=axis, keepdims=keepdims, dtype=a.dtype)


reduce_sum = _reduce(_sum_keep_dtype)
reduce_mean = _reduce(_mean_keep_dtype)
reduce_max = _reduce(np.max)
reduce_min = _reduce(np.min)
reduce_any = _reduce(np.any)
reduce_all = _reduce(np.all)


def argmax(x, axis=None, output_type=int64, name=None) -> Tensor:
    """tf.argmax.  [TF-recall] ties resolve to the lowest index (Eigen's ArgMax reducer keeps the first maximum);
    the TF documentation leaves the tie order unspecified."""
    a = convert_to_tensor(x)._v
    return Tensor(np.argmax(a, axis=0 if axis is None else _axes(axis)).astype(_np_dtype(output_type)))


def gather(params, indices, validate_indices=None, axis=None, batch_dims=0, name=None) -> Tensor:
    p, i = convert_to_tensor(params)._v, convert_to_tensor(indices)._v
    if batch_dims == 0:
        return Tensor(np.take(p, i, axis=0 if axis is None else _axes(axis)))
    if batch_dims != 1 or (axis not in (None, 1)):
        raise NotImplementedError("gather: only batch_dims=1, axis=1")
    if (i < 0).any() or (i >= p.shape[1]).any():
        raise IndexError("gather: index out of range")       # TF raises on CPU
    return Tensor(np.stack([p[b][i[b]] for b in builtins.range(p.shape[0])]))


def one_hot(indices, depth, on_value=None, off_value=None, axis=None, dtype=None, name=None) -> Tensor:
    """tf.one_hot: float32 by default; an index outside [0, depth) gives an all-off row."""
    i = convert_to_tensor(indices)._v
    depth = _pyint(depth)
    dt = _np_dtype(dtype) or np.dtype(np.float32)
    out = (i[..., None] == np.arange(depth, dtype=i.dtype)).astype(dt)
    if on_value is not None or off_value is not None:
        on = 1 if on_value is None else on_value
        off = 0 if off_value is None else off_value
        out = np.where(out != 0, np.asarray(on, dt), np.asarray(off, dt))
    return Tensor(out)


def _top_k_order(a: np.ndarray) -> np.ndarray:
    """Indices that sort the LAST axis in descending order; equal elements keep ascending index order
    (tf.math.top_k documents: "If two elements are equal, the lower-index element appears first")."""
    # stable ascending sort of the negated keys == descending with lower index first; -0.0 == +0.0 compare equal
    return np.argsort(-a if a.dtype.kind == "f" else -a.astype(np.int64), axis=-1, kind="stable")


def argsort(values, axis=-1, direction="ASCENDING", stable=False, name=None) -> Tensor:
    """tf.argsort (sort_ops.py): DESCENDING = top_k over the whole axis; ASCENDING = top_k of the negated values
    (for signed integers ``-values - 1``).  Result is int32."""
    a = convert_to_tensor(values)._v
    if _axes(axis) not in (-1, a.ndim - 1):
        raise NotImplementedError("argsort: last axis only")
    if direction == "DESCENDING":
        return Tensor(_top_k_order(a).astype(np.int32))
    if direction != "ASCENDING":
        raise ValueError(direction)
    neg = -a if a.dtype.kind == "f" else (-a.astype(np.int64) - 1)
    return Tensor(_top_k_order(neg).astype(np.int32))


def sort(values, axis=-1, direction="ASCENDING", name=None) -> Tensor:
    a = convert_to_tensor(values)._v
    idx = argsort(a, axis=axis, direction=direction)._v
    return Tensor(np.take_along_axis(a, idx, axis=-1))


def cond(pred, true_fn=None, false_fn=None, name=None):
    return true_fn() if _pybool(convert_to_tensor(pred)._v) else false_fn()


def pad(tensor, paddings, mode="CONSTANT", constant_values=0, name=None) -> Tensor:
    """tf.pad, CONSTANT mode.  The reference's ``augmentation.expand_image`` (augmentation.py:191-195) passes a nested
    tuple that mixes float32 scalar tensors (already rounded to whole numbers) with Python ints; the values are used as
    whole pixel counts here.  [TF-recall] whether TensorFlow itself accepts float paddings is version dependent."""
    a = convert_to_tensor(tensor)._v
    if mode.upper() != "CONSTANT":
        raise NotImplementedError(mode)

    def whole(v):
        f = _pyfloat(v._v) if isinstance(v, Tensor) else _pyfloat(v)
        if f != int(f):
            raise ValueError(f"padding {f} is not a whole number")
        return int(f)

    if isinstance(paddings, Tensor):
        pads = [tuple(int(v) for v in r) for r in paddings._v.tolist()]
    else:
        pads = [(whole(lo), whole(hi)) for lo, hi in paddings]
    cv = convert_to_tensor(constant_values, dtype_hint=a.dtype)._v
    return Tensor(np.pad(a, pads, mode="constant", constant_values=cv))


def slice(input_, begin, size, name=None) -> Tensor:       # noqa: A001
    a = convert_to_tensor(input_)._v
    b = [_pyint(v) for v in convert_to_tensor(begin)._v.tolist()]
    s = [_pyint(v) for v in convert_to_tensor(size)._v.tolist()]
    idx = tuple(builtins.slice(bi, None if si == -1 else bi + si) for bi, si in zip(b, s))
    return Tensor(a[idx])


# ----------------------------------------------------------------- sub-modules --
math = types.ModuleType("tensorflow.math")
math.log = _log
math.exp = exp
math.sqrt = sqrt
math.abs = abs
math.maximum = maximum
math.minimum = minimum
math.reduce_sum = reduce_sum
math.reduce_max = reduce_max
math.reduce_mean = reduce_mean
math.argmax = argmax
math.equal = equal
math.top_k = None     # set below


def _top_k(input, k=1, sorted=True, name=None):      # noqa: A002
    a = convert_to_tensor(input)._v
    idx = _top_k_order(a)[..., :_pyint(k)]
    return Tensor(np.take_along_axis(a, idx, axis=-1)), Tensor(idx.astype(np.int32))


math.top_k = _top_k


def _l2_normalize(x, axis=None, epsilon=1e-12, name=None, dim=None) -> Tensor:
    """tf.nn.l2_normalize (documented): ``output = x / sqrt(max(sum(x**2), epsilon))``, computed as
    ``x * rsqrt(max(sum(x*x), epsilon))`` (nn_impl.py)."""
    a = convert_to_tensor(x)._v
    axis = dim if axis is None else axis
    sq = np.sum(a * a, axis=_axes(axis), keepdims=True, dtype=a.dtype)
    inv = (np.asarray(1, a.dtype) / np.sqrt(np.maximum(sq, np.asarray(epsilon, a.dtype)))).astype(a.dtype)
    return Tensor(a * inv)


def _softmax(logits, axis=-1, name=None) -> Tensor:
    a = convert_to_tensor(logits)._v
    m = np.max(a, axis=_axes(axis), keepdims=True)
    e = np.exp(a - m)
    return Tensor(e / np.sum(e, axis=_axes(axis), keepdims=True, dtype=a.dtype))


def _moments(x, axes, shift=None, keepdims=False, name=None):
    a = convert_to_tensor(x)._v
    ax = _axes(axes)
    mean = np.mean(a, axis=ax, keepdims=True, dtype=a.dtype)
    var = np.mean(np.square(a - mean), axis=ax, keepdims=True, dtype=a.dtype)
    if not keepdims:
        mean, var = np.squeeze(mean, ax), np.squeeze(var, ax)
    return Tensor(mean), Tensor(var)


nn = types.ModuleType("tensorflow.nn")
nn.l2_normalize = _l2_normalize
nn.softmax = _softmax
nn.relu = _unary(lambda a: np.maximum(a, np.asarray(0, a.dtype)))
nn.relu6 = _unary(lambda a: np.minimum(np.maximum(a, np.asarray(0, a.dtype)), np.asarray(6, a.dtype)))
nn.moments = _moments


# -- tf.random: the stream is the caller's (fixtures pass their own generator) ---
class _Random(types.ModuleType):
    def __init__(self):
        super().__init__("tensorflow.random")
        self.generator = np.random.default_rng(0)
        self.queue = []          # when non-empty, uniform() pops pre-drawn values instead (augmentation fixtures)
        self.log = []            # every draw made, in order (so that a device path can be fed the same draws)

    def set_seed(self, seed):
        self.generator = np.random.default_rng(seed)

    def uniform(self, shape=(), minval=0, maxval=None, dtype=float32, seed=None, name=None) -> Tensor:
        dt = _np_dtype(dtype)
        shp = _static_shape(shape)
        lo = convert_to_tensor(minval, dtype_hint=dt)._v
        if dt.kind == "f":
            hi = convert_to_tensor(1 if maxval is None else maxval, dtype_hint=dt)._v
            u = np.asarray(self.queue.pop(0), dt).reshape(shp) if self.queue else self.generator.random(shp, dtype=dt)
            # random_ops.random_uniform: rnd * (maxval - minval) + minval
            out = (u * (hi - lo).astype(dt) + lo.astype(dt)).astype(dt)
        else:
            if maxval is None:
                raise ValueError("integer uniform needs maxval")
            hi = convert_to_tensor(maxval, dtype_hint=dt)._v
            out = (np.asarray(self.queue.pop(0), dt).reshape(shp) if self.queue
                   else self.generator.integers(_pyint(lo), _pyint(hi), size=shp).astype(dt))
        self.log.append(np.array(out))
        return Tensor(out)


random = _Random()


# -- tf.losses (Keras losses, reduction NONE) -----------------------------------
class _Reduction(object):
    NONE = "none"
    SUM = "sum"
    SUM_OVER_BATCH_SIZE = "sum_over_batch_size"
    AUTO = "auto"


class _Loss(object):
    def __init__(self, reduction=_Reduction.AUTO, name=None):
        self.reduction = reduction

    def __call__(self, y_true, y_pred, sample_weight=None):
        out = self.call(convert_to_tensor(y_true), convert_to_tensor(y_pred))
        if sample_weight is not None:
            raise NotImplementedError("sample_weight")
        if self.reduction == _Reduction.NONE:
            return out
        if self.reduction == _Reduction.SUM:
            return reduce_sum(out)
        return reduce_mean(out)


class Huber(_Loss):
    """tf.keras.losses.Huber (documented): with ``x = y_true - y_pred``: ``0.5 * x^2`` if ``|x| <= d`` else
    ``0.5 * d^2 + d * (|x| - d)``, written as in losses.py::huber_loss (quadratic = min(|x|, d);
    linear = |x| - quadratic; ``0.5 * quadratic^2 + d * linear``).

    ``mean_last_axis``: TensorFlow >= 2.2 averages over the last axis (rank drops by one); TensorFlow 2.0/2.1 return
    the element-wise tensor.  ``ssd_loss.py:38-43`` handles both; the fixtures are generated under both settings and
    must agree.  The class attribute selects the behaviour."""

    mean_last_axis = True

    def __init__(self, delta=1.0, reduction=_Reduction.AUTO, name="huber_loss"):
        super().__init__(reduction, name)
        self.delta = delta

    def call(self, y_true, y_pred):
        d = convert_to_tensor(self.delta, dtype_hint=y_pred.dtype)
        error = subtract(y_pred, y_true)
        abs_error = abs(error)
        quadratic = minimum(abs_error, d)
        linear = subtract(abs_error, quadratic)
        per_elem = add(multiply(convert_to_tensor(0.5, dtype_hint=quadratic.dtype), multiply(quadratic, quadratic)),
                       multiply(d, linear))
        return reduce_mean(per_elem, axis=-1) if self.mean_last_axis else per_elem


_KERAS_EPSILON = 1e-7      # keras.backend.epsilon()


class CategoricalCrossentropy(_Loss):
    """tf.keras.losses.CategoricalCrossentropy on PROBABILITIES (from_logits=False), as documented for
    keras.backend.categorical_crossentropy: scale the predictions so each row sums to 1, clip to
    [epsilon, 1 - epsilon] with epsilon = 1e-7, return ``-sum(target * log(output), axis=-1)``."""

    def __init__(self, from_logits=False, label_smoothing=0, reduction=_Reduction.AUTO, name="categorical_crossentropy"):
        super().__init__(reduction, name)
        if label_smoothing:
            raise NotImplementedError("label_smoothing")
        self.from_logits = from_logits

    def call(self, y_true, y_pred):
        if self.from_logits:
            y_pred = _softmax(y_pred)
        output = truediv(y_pred, reduce_sum(y_pred, axis=-1, keepdims=True))
        eps = convert_to_tensor(_KERAS_EPSILON, dtype_hint=output.dtype)
        output = clip_by_value(output, eps, subtract(convert_to_tensor(1.0, dtype_hint=output.dtype), eps))
        return negative(reduce_sum(multiply(y_true, _log(output)), axis=-1))


losses = types.ModuleType("tensorflow.losses")
losses.Huber = Huber
losses.CategoricalCrossentropy = CategoricalCrossentropy
losses.Reduction = _Reduction

from . import image      # noqa: E402,F401   (tf.image)
from . import keras      # noqa: E402,F401   (tf.keras)

keras.losses = losses

config = types.SimpleNamespace(experimental=types.SimpleNamespace(
    list_physical_devices=lambda device_type=None: [], set_memory_growth=lambda gpu, enabled: None))
