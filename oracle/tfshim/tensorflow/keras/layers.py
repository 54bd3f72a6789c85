"""tf.keras.layers subset in NumPy.  Weight-tracking order follows Keras: a layer's own weights
first, then its tracked sub-layers in attribute-assignment order (lists flattened), trainable
before non-trainable (``_legacy_weights``)."""
import inspect
import re

import numpy as _np

_UID = {}
_TRAINING = [None]          # call-context `training` value, inherited by nested calls (Keras call context)
_INIT_RNG = _np.random.default_rng(12345)


def _snake(name):
    s = re.sub("(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    s = re.sub("([a-z])([A-Z])", r"\1_\2", s).lower()
    return "private" + s if s[0] == "_" else s


def _unique_name(base):
    n = _UID.get(base, 0)
    _UID[base] = n + 1
    return base if n == 0 else "%s_%d" % (base, n)


def reset_uids():
    _UID.clear()


class Layer:
    def __init__(self, trainable=True, name=None, dtype=None, **kwargs):
        object.__setattr__(self, "_attr_order", [])
        self._own_weights = []
        self.built = False
        self.trainable = True
        self.name = name if name else _unique_name(_snake(type(self).__name__))

    # ---- tracking -------------------------------------------------------------------------------
    def __setattr__(self, key, value):
        order = self.__dict__.get("_attr_order")
        if order is not None and key not in order and not key.startswith("_"):
            if isinstance(value, (Layer, list)):
                order.append(key)
        object.__setattr__(self, key, value)

    def _tracked_layers(self):
        for key in self._attr_order:
            v = self.__dict__.get(key)
            if isinstance(v, Layer):
                yield v
            elif isinstance(v, list):
                for e in v:
                    if isinstance(e, Layer):
                        yield e

    def add_weight(self, name=None, shape=None, dtype=None, initializer=None, trainable=True, **kwargs):
        import tensorflow as tf
        shape = tuple(int(s) for s in shape)
        if initializer is None or initializer == "zeros":
            value = _np.zeros(shape)
        elif initializer == "ones":
            value = _np.ones(shape)
        elif initializer == "glorot_uniform":
            # keras GlorotUniform: limit = sqrt(6 / (fan_in + fan_out)), conv fans scaled by the receptive field
            rf = int(_np.prod(shape[:-2])) if len(shape) > 2 else 1
            fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
            lim = _np.sqrt(6.0 / (fan_in + fan_out))
            value = _INIT_RNG.uniform(-lim, lim, size=shape)
        else:
            value = initializer(shape, _INIT_RNG)
        v = tf.Variable(value, name=(name or "weight") + ":0", trainable=trainable)
        self._own_weights.append(v)
        return v

    @property
    def trainable_weights(self):
        out = [w for w in self._own_weights if w.trainable]
        for l in self._tracked_layers():
            out += l.trainable_weights
        return out

    @property
    def non_trainable_weights(self):
        out = [w for w in self._own_weights if not w.trainable]
        for l in self._tracked_layers():
            out += l.non_trainable_weights
        return out

    @property
    def weights(self):
        return self.trainable_weights + self.non_trainable_weights

    @property
    def trainable_variables(self):
        return self.trainable_weights

    # ---- call -----------------------------------------------------------------------------------
    def build(self, input_shape):
        self.built = True

    def __call__(self, *args, **kwargs):
        params = inspect.signature(self.call).parameters
        pushed = False
        if "training" in kwargs and kwargs["training"] is not None:
            _TRAINING.append(kwargs["training"])
            pushed = True
        elif "training" in params and "training" not in kwargs:
            names = list(params)
            if names.index("training") >= len(args):      # not passed positionally: inherit the call context
                kwargs["training"] = _TRAINING[-1]
        if "training" in kwargs and "training" not in params:
            kwargs.pop("training")
        try:
            if not self.built:
                first = args[0] if args else next(iter(kwargs.values()))
                if not isinstance(first, (list, tuple)):
                    self.build(_np.shape(first))
                self.built = True
            return self.call(*args, **kwargs)
        finally:
            if pushed:
                _TRAINING.pop()

    def call(self, inputs, *args, **kwargs):
        return inputs


def _f(x):
    import tensorflow as tf
    return _np.asarray(x, dtype=tf.float32) if _np.asarray(x).dtype.kind == "f" else _np.asarray(x)


class Dense(Layer):
    """keras.layers.Dense: tensordot(x, kernel(in, out)) + bias on the last axis."""

    def __init__(self, units, activation=None, use_bias=True, **kwargs):
        super().__init__(**kwargs)
        self.units, self.use_bias, self.activation = int(units), use_bias, activation

    def build(self, input_shape):
        self.kernel = self.add_weight(self.name + "/kernel", (input_shape[-1], self.units), initializer="glorot_uniform")
        self.bias = self.add_weight(self.name + "/bias", (self.units,), initializer="zeros") if self.use_bias else None

    def call(self, inputs):
        y = _np.matmul(_f(inputs), _np.asarray(self.kernel))
        if self.bias is not None:
            y = y + _np.asarray(self.bias)
        return self.activation(y) if self.activation else y


class Conv1D(Layer):
    """keras.layers.Conv1D, channels_last, padding 'valid': cross-correlation,
    out[b, t] = sum_k x[b, t*stride + k] @ kernel[k] + bias, kernel (k, in, out)."""

    def __init__(self, filters, kernel_size, strides=1, padding="valid", use_bias=True, **kwargs):
        super().__init__(**kwargs)
        assert padding == "valid"
        self.filters, self.kernel_size, self.strides, self.use_bias = int(filters), int(kernel_size), int(strides), use_bias

    def build(self, input_shape):
        self.kernel = self.add_weight(self.name + "/kernel", (self.kernel_size, input_shape[-1], self.filters),
                                      initializer="glorot_uniform")
        self.bias = self.add_weight(self.name + "/bias", (self.filters,), initializer="zeros") if self.use_bias else None

    def call(self, inputs):
        x = _f(inputs)
        k, s = self.kernel_size, self.strides
        n_out = (x.shape[1] - k) // s + 1
        y = 0
        for i in range(k):
            y = y + _np.matmul(x[:, i:i + (n_out - 1) * s + 1:s], _np.asarray(self.kernel)[i])
        if self.bias is not None:
            y = y + _np.asarray(self.bias)
        return y


class ZeroPadding1D(Layer):
    def __init__(self, padding=1, **kwargs):
        super().__init__(**kwargs)
        self.padding = (padding, padding) if isinstance(padding, int) else (int(padding[0]), int(padding[1]))

    def call(self, inputs):
        return _np.pad(_f(inputs), ((0, 0), self.padding, (0, 0)))


class MaxPool1D(Layer):
    """keras.layers.MaxPool1D, padding 'valid': out[t] = max(x[t*strides : t*strides + pool_size])."""

    def __init__(self, pool_size=2, strides=None, padding="valid", **kwargs):
        super().__init__(**kwargs)
        assert padding == "valid"
        self.pool_size, self.strides = int(pool_size), int(strides or pool_size)

    def call(self, inputs):
        x = _f(inputs)
        n_out = (x.shape[1] - self.pool_size) // self.strides + 1
        return _np.stack([x[:, t * self.strides:t * self.strides + self.pool_size].max(axis=1) for t in range(n_out)], axis=1)


MaxPooling1D = MaxPool1D


class LayerNormalization(Layer):
    """keras.layers.LayerNormalization over the last axis (non-fused path, taken for epsilon < 1.001e-5):
    mean/variance = tf.nn.moments (biased), y = x * (gamma * rsqrt(var + eps)) + (beta - mean * gamma * rsqrt(var + eps))."""

    def __init__(self, axis=-1, epsilon=1e-3, **kwargs):
        super().__init__(**kwargs)
        self.epsilon = epsilon

    def build(self, input_shape):
        self.gamma = self.add_weight(self.name + "/gamma", (input_shape[-1],), initializer="ones")
        self.beta = self.add_weight(self.name + "/beta", (input_shape[-1],), initializer="zeros")

    def call(self, inputs):
        x = _f(inputs)
        mean = x.mean(axis=-1, keepdims=True)
        var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)
        inv = _np.asarray(self.gamma) / _np.sqrt(var + x.dtype.type(self.epsilon))
        return x * inv + (_np.asarray(self.beta) - mean * inv)


class Dropout(Layer):
    def __init__(self, rate, **kwargs):
        super().__init__(**kwargs)
        self.rate = rate

    def call(self, inputs, training=None):
        if training and self.rate > 0:
            raise NotImplementedError("tfshim: dropout with rate > 0 in training mode is not restated")
        return inputs


class Activation(Layer):
    def __init__(self, activation, **kwargs):
        super().__init__(**kwargs)
        self.activation = activation

    def call(self, inputs):
        return self.activation(inputs)


class BatchNormalization(Layer):
    def __init__(self, *a, **kw):
        raise NotImplementedError("tfshim: OUTPUT_BN is false in every shipped config")
