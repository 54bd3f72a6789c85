"""NumPy stand-in for the subset of TensorFlow 2.4 the reference calls (see ../README.md).

Tensors are plain ``numpy.ndarray``; ``tf.float32`` maps to the dtype selected by the environment
variable ``TFSHIM_DTYPE`` (``float64`` = truth run, ``float32`` = same precision as TF on CPU).
"""
import os as _os

import numpy as _np

__version__ = "2.4.3-numpy-shim"

_FLOAT = _np.dtype(_os.environ.get("TFSHIM_DTYPE", "float64")).type
float32 = _FLOAT
float64 = _np.float64
int32 = _np.int32
int64 = _np.int64
bool = _np.bool_          # noqa: A001  (tf.bool)
newaxis = None


def set_float_dtype(name):
    """Switch what tf.float32 means (float64 = truth run, float32 = TF-CPU precision); affects tensors and
    variables created afterwards."""
    global _FLOAT, float32
    _FLOAT = _np.dtype(name).type
    float32 = _FLOAT


class Tensor(_np.ndarray):          # only so that isinstance checks in third-party code (einops) resolve
    pass


class _Weight(_np.ndarray):
    """Keras weight: an ndarray view carrying a name (einops and numpy treat it as a plain array)."""

    def __new__(cls, value, name=None, trainable=True):
        obj = _np.array(value, dtype=_FLOAT).view(cls)
        obj.name = name
        obj.trainable = trainable
        return obj

    def __array_finalize__(self, obj):
        self.name = getattr(obj, "name", None)
        self.trainable = getattr(obj, "trainable", True)

    def assign(self, value):
        self[...] = _np.asarray(value, dtype=self.dtype).reshape(self.shape)
        return self

    def numpy(self):
        return _np.asarray(self)


class Variable:
    """tf.Variable(...) hands out a ``_Weight``; the class itself only serves annotations / isinstance probes."""

    def __new__(cls, value, name=None, trainable=True, **_kw):
        return _Weight(value, name=name, trainable=trainable)


def is_tensor(x):
    return False


def _plain(x):
    return _np.asarray(x)


def function(fn=None, **_kw):        # tf.function: eager execution
    if fn is None:
        return lambda f: f
    return fn


def shape(x):
    return _np.array(_np.shape(x), dtype=_np.int32)


def rank(x):
    return _np.ndim(x)


def ones(shape, dtype=None):         # noqa: A002
    return _np.ones(_np.asarray(shape, dtype=_np.int64), dtype=dtype or _FLOAT)


def zeros(shape, dtype=None):        # noqa: A002
    return _np.zeros(_np.asarray(shape, dtype=_np.int64), dtype=dtype or _FLOAT)


def concat(values, axis=0):
    return _np.concatenate([_np.atleast_1d(_plain(v)) for v in values], axis=axis)


def reshape(x, shape):               # noqa: A002
    return _np.reshape(_plain(x), [int(s) for s in shape])


def transpose(x, perm=None):
    return _np.transpose(_plain(x), perm)


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _plain(a), _plain(b)
    if transpose_a:
        a = _np.swapaxes(a, -1, -2)
    if transpose_b:
        b = _np.swapaxes(b, -1, -2)
    return _np.matmul(a, b)


def cast(x, dtype):
    return _plain(x).astype(dtype)


def expand_dims(x, axis):
    return _np.expand_dims(_plain(x), axis)


def range(start, limit=None, delta=1, dtype=None):     # noqa: A001
    if limit is None:
        start, limit = 0, start
    return _np.arange(start, limit, delta, dtype=dtype)


def not_equal(a, b):
    return _np.not_equal(a, b)


def logical_and(a, b):
    return _np.logical_and(a, b)


def norm(x, axis=None):
    """tf.norm(ord='euclidean'): sqrt(reduce_sum(x * x, axis))."""
    x = _plain(x)
    return _np.sqrt(_np.sum(x * x, axis=axis))


def reduce_sum(x, axis=None, keepdims=False):
    return _np.sum(_plain(x), axis=axis, keepdims=keepdims)


def minimum(a, b):
    return _np.minimum(_plain(a), b)


def maximum(a, b):
    return _np.maximum(_plain(a), b)


def tile(x, multiples):
    return _np.tile(_plain(x), tuple(int(m) for m in multiples))


class _Linalg:
    @staticmethod
    def cross(a, b):
        return _np.cross(_plain(a), _plain(b))


linalg = _Linalg()


def reduce_mean(x, axis=None):
    return _np.mean(_plain(x), axis=axis)


class _Math:
    @staticmethod
    def sqrt(x):
        return _np.sqrt(x)

    @staticmethod
    def floor(x):
        return _np.floor(x)


math = _Math()


class _NN:
    @staticmethod
    def softmax(logits, axis=-1):
        """tf.nn.softmax: exp(x - max) / sum(exp(x - max)) along ``axis`` in the tensor's dtype."""
        x = _plain(logits)
        e = _np.exp(x - _np.max(x, axis=axis, keepdims=True))
        return e / _np.sum(e, axis=axis, keepdims=True)

    @staticmethod
    def relu(x):
        return _np.maximum(x, 0)


nn = _NN()


class _Random:
    def __init__(self):
        self._rng = _np.random.default_rng(0)

    def set_seed(self, seed):
        self._rng = _np.random.default_rng(seed)

    def uniform(self, shape, minval=0.0, maxval=1.0, dtype=None):    # noqa: A002
        return self._rng.uniform(minval, maxval, size=[int(s) for s in _np.atleast_1d(shape)]).astype(dtype or _FLOAT)


random = _Random()

from . import keras  # noqa: E402,F401
