"""tf.keras subset: layers, Model, activations, initializers (NumPy; see ../../README.md)."""
import numpy as _np

from . import layers  # noqa: F401
from .layers import Layer as _Layer


class _Activations:
    @staticmethod
    def gelu(x, approximate=False):
        """keras.activations.gelu, approximate=False: 0.5 x (1 + erf(x / sqrt(2)))."""
        from scipy.special import erf
        x = _np.asarray(x)
        if approximate:
            return 0.5 * x * (1.0 + _np.tanh(_np.sqrt(2.0 / _np.pi) * (x + 0.044715 * x ** 3)))
        return (0.5 * x * (1.0 + erf(x / _np.sqrt(x.dtype.type(2.0))))).astype(x.dtype)

    @staticmethod
    def relu(x):
        return _np.maximum(x, 0)

    @staticmethod
    def linear(x):
        return x


activations = _Activations()


class _Backend:
    @staticmethod
    def is_keras_tensor(x):
        return False


backend = _Backend()


class _TruncatedNormal:
    """keras.initializers.TruncatedNormal: N(mean, stddev) redrawn outside two standard deviations."""

    def __init__(self, mean=0.0, stddev=0.05, seed=None):
        self.mean, self.stddev, self.seed = mean, stddev, seed

    def __call__(self, shape, rng):
        out = rng.normal(self.mean, self.stddev, size=shape)
        bad = _np.abs(out - self.mean) > 2 * self.stddev
        while bad.any():
            out[bad] = rng.normal(self.mean, self.stddev, size=int(bad.sum()))
            bad = _np.abs(out - self.mean) > 2 * self.stddev
        return out


class _Initializers:
    TruncatedNormal = _TruncatedNormal


initializers = _Initializers()


class Model(_Layer):
    """Subclassed keras.Model: variables are created by tracing ``call`` once in ``build``."""

    _is_graph_network = False
    _distribution_strategy = None

    def build(self, input_shape):
        import tensorflow as tf
        if isinstance(input_shape, list):
            x = _np.zeros(input_shape[0], dtype=tf.float32)
            m = _np.ones(input_shape[1], dtype=_np.bool_)
            self([x, m], training=False)
        else:
            self(_np.zeros(input_shape, dtype=tf.float32), training=False)
        self.built = True

    def _assert_weights_created(self):
        if not self.built:
            raise ValueError("Weights for model %s have not yet been created." % self.name)

    @property
    def layers(self):
        """Top-level layers in attribute-tracking order, lists flattened (Model.layers)."""
        return list(self._tracked_layers())

    def get_weights(self):
        return [_np.array(w) for w in self.weights]

    def set_weights(self, values):
        ws = self.weights
        assert len(ws) == len(values)
        for w, v in zip(ws, values):
            w.assign(v)

    def count_params(self):
        return int(sum(w.size for w in self.weights))
