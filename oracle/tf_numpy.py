"""A numpy-eager stand-in for the ~25 TensorFlow / Keras symbols ``/root/reference/kgcn/layers.py`` touches, so the
reference's OWN layer code (its loops, op order, indexing, bias placement, channel sums, padding rules) can be
executed unchanged in the authoring container and its outputs committed as golden vectors
(``oracle/make_layer_golden.py`` -> ``tests/golden/layers_*.npz``).

TEST INFRASTRUCTURE ONLY.  Nothing under ``kgcn_b200/`` imports this module.

What is and is not the reference here: every line of control flow is the reference's; the TensorFlow PRIMITIVES it
calls are restated below in float32 numpy with their documented semantics (TensorFlow 1.15 is not installable in this
image).  Each primitive cites the TF op it stands for:

  tf.matmul                         a @ b
  tf.sparse_tensor_dense_matmul     out[i, :] += v * b[j, :] per stored entry, in storage order (the CPU kernel's loop)
  SparseTensor * dense vector       sparse_dense_cwise_mul: the dense operand broadcasts against the LAST axis
  tf.sparse_tensor_to_dense         zeros + scatter (validate_indices: sorted, unique)
  tf.add_n / stack / unstack / concat / split / pad / reshape / reduce_sum / reduce_max / shape / nn.relu / nn.bias_add
  keras Layer (add_weight / build / __call__), Dense, BatchNormalization (learning phase 0: moving statistics,
  exactly what the reference trainer runs; ``tf.layers.batch_normalization(training=True)``:
  batch statistics)

``install_model_level`` adds the symbols the reference's MODEL files touch on top of the layers (example_model/model.py:
placeholders, tf.sigmoid, softmax, softmax cross-entropy, reduce_mean, cast / equal / argmax, contrib.keras Dense and
Dropout at learning phase 0), so ``oracle/make_c1_golden.py`` can execute a model file unchanged.
"""
import collections
import sys
import types

import numpy as np

F32 = np.float32
_RNG = np.random.default_rng(0)


def seed(value):
    """Seed the initialisers (glorot_uniform) of the stand-in."""
    global _RNG
    _RNG = np.random.default_rng(value)


class Tensor(np.ndarray):
    """float32 ndarray with the two TF tensor methods the layers call."""

    def set_shape(self, shape):   # static-shape hint only
        return None

    def get_shape(self):
        return self.shape


def T(x, dtype=F32):
    return np.asarray(x, dtype).view(Tensor)


class SparseTensorValue(collections.namedtuple("SparseTensorValue", ["indices", "values", "dense_shape"])):
    """tf.SparseTensorValue / tf.SparseTensor: what feed.py builds (feed.py:122,126) and the layers consume."""

    def __mul__(self, dense):
        # sparse_dense_cwise_mul: `dense` broadcasts to the sparse operand's shape; a rank-1 operand lines up with
        # the last axis, i.e. entry (i, j) is scaled by dense[j]  (kgcn/layers.py:143)
        dense = np.asarray(dense, F32)
        idx = np.asarray(self.indices).reshape(-1, len(self.dense_shape))
        if dense.ndim != 1:
            raise NotImplementedError("only the rank-1 broadcast of kgcn/layers.py:143 is restated")
        return SparseTensorValue(idx, (np.asarray(self.values, F32) * dense[idx[:, -1]]).astype(F32), self.dense_shape)


def _as_sparse(sp):
    if isinstance(sp, SparseTensorValue):
        return sp
    return SparseTensorValue(*sp)


def sparse_tensor_dense_matmul(sp, b, adjoint_a=False, adjoint_b=False):
    sp = _as_sparse(sp)
    b = np.asarray(b, F32)
    if adjoint_b:
        b = b.T
    idx = np.asarray(sp.indices).reshape(-1, 2)
    rows, cols = int(sp.dense_shape[0]), int(sp.dense_shape[1])
    if adjoint_a:
        idx = idx[:, ::-1]
        rows, cols = cols, rows
    if idx.size and (idx.min() < 0 or idx[:, 0].max() >= rows or idx[:, 1].max() >= cols):
        raise ValueError("InvalidArgumentError: sparse index out of range")    # TF raises at run time
    out = np.zeros((rows, b.shape[1]), F32)
    vals = np.asarray(sp.values, F32)
    for k in range(idx.shape[0]):                         # storage order, like the CPU kernel
        out[idx[k, 0]] = out[idx[k, 0]] + vals[k] * b[idx[k, 1]]
    return T(out)


def sparse_tensor_to_dense(sp, default_value=0, validate_indices=True):
    sp = _as_sparse(sp)
    idx = np.asarray(sp.indices).reshape(-1, 2)
    if validate_indices and idx.shape[0] > 1:
        flat = idx[:, 0] * int(sp.dense_shape[1]) + idx[:, 1]
        if (np.diff(flat) <= 0).any():
            raise ValueError("InvalidArgumentError: indices are out of order or repeated")
    out = np.full(tuple(int(v) for v in sp.dense_shape), default_value, F32)
    out[idx[:, 0], idx[:, 1]] = np.asarray(sp.values, F32)
    return T(out)


def _init(initializer, shape):
    if callable(initializer):
        return T(initializer(shape))
    if initializer in ("zeros", "zero"):
        return T(np.zeros(shape, F32))
    if initializer in ("ones", "one"):
        return T(np.ones(shape, F32))
    if initializer == "glorot_uniform":
        fan_in, fan_out = (shape[0], shape[-1]) if len(shape) >= 2 else ((shape[0], shape[0]) if shape else (1, 1))
        limit = np.sqrt(6.0 / (fan_in + fan_out))
        return T(_RNG.uniform(-limit, limit, size=shape))
    raise ValueError("initializer %r is not restated" % (initializer,))


class Layer:
    """keras.layers.Layer, eager: weights are created on the first call from the input shape."""

    def __init__(self, name=None, trainable=True, **kwargs):
        self.name, self.trainable, self.built = name, trainable, False
        self.weights = []

    def add_weight(self, name=None, shape=(), initializer="zeros", trainable=True, **kwargs):
        shape = tuple(int(s) for s in np.atleast_1d(shape)) if shape != () else ()
        value = _init(initializer, shape)
        self.weights.append((name, value))
        return value

    def build(self, input_shape):
        self.built = True

    def __call__(self, inputs, *args, **kwargs):
        if not self.built:
            shape = [np.shape(t) if not isinstance(t, SparseTensorValue) else tuple(t.dense_shape) for t in inputs] \
                if isinstance(inputs, (list, tuple)) else np.shape(inputs)
            self.build(shape)
            self.built = True
        return self.call(inputs, *args, **kwargs)


_ACT = {None: lambda x: x, "linear": lambda x: x, "relu": lambda x: np.maximum(x, F32(0)),
        "sigmoid": lambda x: (F32(1) / (F32(1) + np.exp(-x, dtype=F32))).astype(F32), "tanh": lambda x: np.tanh(x, dtype=F32)}


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform", bias_initializer="zeros",
                 **kwargs):
        super().__init__(**kwargs)
        self.units, self.use_bias = int(units), use_bias
        self.activation = activation if callable(activation) else _ACT[activation]
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer

    def build(self, input_shape):
        self.kernel = self.add_weight("kernel", (int(input_shape[-1]), self.units), self.kernel_initializer)
        self.bias = self.add_weight("bias", (self.units,), self.bias_initializer) if self.use_bias else None
        self.built = True

    def call(self, inputs, **kwargs):
        out = (np.asarray(inputs, F32) @ self.kernel).astype(F32)
        if self.bias is not None:
            out = (out + self.bias).astype(F32)
        return T(self.activation(out))


class BatchNormalization(Layer):
    """keras BatchNormalization called without ``training=`` under the reference trainer: learning phase 0, i.e. the
    moving statistics (initial mean 0 / variance 1) are used."""

    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, trainable=True, name=None, **kwargs):
        super().__init__(name=name, trainable=trainable)
        self.epsilon = epsilon

    def build(self, input_shape):
        f = int(input_shape[-1])
        self.gamma, self.beta = self.add_weight("gamma", (f,), "ones"), self.add_weight("beta", (f,), "zeros")
        self.moving_mean, self.moving_variance = T(np.zeros(f)), T(np.ones(f))
        self.built = True

    def call(self, inputs, **kwargs):
        x = np.asarray(inputs, np.float64)
        y = (x - self.moving_mean) / np.sqrt(np.asarray(self.moving_variance, np.float64) + self.epsilon) * self.gamma + self.beta
        return T(y)


# Variables for the next tf.layers.batch_normalization calls, consumed in call order: dicts with gamma / beta /
# moving_mean / moving_variance (what a restored checkpoint provides; oracle/make_ckpt_golden.py).  Empty = fresh
# variables (gamma 1, beta 0, moving mean 0 / variance 1).
BN_VARIABLES = []


def _batch_normalization(inputs, training=False, name=None, epsilon=1e-3, **kwargs):
    """tf.layers.batch_normalization: training=True -> batch statistics, else the moving ones; y = x_hat * gamma + beta."""
    x = np.asarray(inputs, np.float64)
    f = x.shape[-1]
    v = BN_VARIABLES.pop(0) if BN_VARIABLES else {}
    if training:
        mean, var = x.mean(0), x.var(0)
    else:
        mean, var = np.asarray(v.get("moving_mean", np.zeros(f)), np.float64), np.asarray(v.get("moving_variance", np.ones(f)), np.float64)
    y = (x - mean) / np.sqrt(var + epsilon)
    if v:
        y = y * np.asarray(v["gamma"], np.float64) + np.asarray(v["beta"], np.float64)
    return T(y)


def install():
    """Put the stand-in on sys.modules (``tensorflow`` and the sub-modules kgcn/layers.py imports) and return it."""
    tf = types.ModuleType("tensorflow")
    tf.__version__ = "1.15.0"
    tf._kgcn_b200_numpy_tf = True
    tf.SparseTensorValue = tf.SparseTensor = SparseTensorValue
    tf.float32, tf.int32, tf.int64 = np.float32, np.int32, np.int64
    tf.matmul = lambda a, b, **k: T(np.asarray(a, F32) @ np.asarray(b, F32))
    tf.add = lambda a, b: T(np.asarray(a, F32) + np.asarray(b, F32))
    tf.sparse_tensor_dense_matmul = sparse_tensor_dense_matmul
    tf.sparse_tensor_to_dense = sparse_tensor_to_dense

    def add_n(xs):
        acc = np.asarray(xs[0], F32)
        for x in xs[1:]:
            acc = (acc + np.asarray(x, F32)).astype(F32)      # left to right, like AddN's CPU kernel
        return T(acc)

    tf.add_n = add_n
    tf.stack = lambda xs, axis=0: T(np.stack([np.asarray(x, F32) for x in xs], axis))
    tf.unstack = lambda x, axis=0: [T(v) if np.asarray(v).dtype.kind == "f" else v for v in np.moveaxis(np.asarray(x), axis, 0)]
    tf.concat = lambda xs, axis: T(np.concatenate([np.asarray(x, F32) for x in xs], axis))
    tf.split = lambda x, sizes, axis=0: [T(v) for v in np.split(np.asarray(x, F32), np.cumsum(np.asarray(sizes))[:-1], axis)]
    tf.pad = lambda x, paddings: T(np.pad(np.asarray(x, F32), [(int(a), int(b)) for a, b in paddings]))
    tf.reshape = lambda x, shape: T(np.reshape(np.asarray(x, F32), [int(s) for s in shape]))
    tf.shape = lambda x, out_type=None: np.asarray(np.shape(x))
    tf.reduce_sum = lambda x, axis=None, keepdims=False: T(np.sum(np.asarray(x, F32), axis=axis, keepdims=keepdims, dtype=F32))
    tf.reduce_max = lambda x, axis=None: T(np.max(np.asarray(x, F32), axis=axis))
    tf.expand_dims = lambda x, axis: T(np.expand_dims(np.asarray(x, F32), axis))
    nn = types.ModuleType("tensorflow.nn")
    nn.relu = lambda x: T(np.maximum(np.asarray(x, F32), F32(0)))
    nn.bias_add = lambda x, b: T(np.asarray(x, F32) + np.asarray(b, F32))
    tf.nn = nn
    layers = types.ModuleType("tensorflow.layers")
    layers.batch_normalization = _batch_normalization
    tf.layers = layers
    klayers = types.ModuleType("tensorflow.python.keras.layers")
    klayers.Layer, klayers.Dense, klayers.BatchNormalization = Layer, Dense, BatchNormalization
    keras = types.ModuleType("tensorflow.keras")
    keras.layers = klayers
    tf.keras = keras
    python = types.ModuleType("tensorflow.python")
    pykeras = types.ModuleType("tensorflow.python.keras")
    pykeras.layers = klayers
    python.keras = pykeras
    tf.python = python
    mods = {"tensorflow": tf, "tensorflow.nn": nn, "tensorflow.layers": layers, "tensorflow.keras": keras,
            "tensorflow.keras.layers": klayers, "tensorflow.python": python, "tensorflow.python.keras": pykeras,
            "tensorflow.python.keras.layers": klayers}
    sys.modules.update(mods)
    return tf


class _Placeholder:
    """tf.placeholder / tf.sparse_placeholder: an identity the reference's construct_feed keys its feed_dict with."""

    def __init__(self, dtype, shape=None, name=None, sparse=False):
        self.dtype, self.shape, self.name, self.sparse = dtype, shape, name, sparse

    def __repr__(self):
        return "<placeholder %s>" % self.name


class Dropout(Layer):
    """keras Dropout called without ``training=`` under the reference trainer: learning phase 0 -> identity
    (SURVEY.md App. A.10)."""

    def __init__(self, rate=0.0, **kwargs):
        super().__init__(**kwargs)
        self.rate = rate

    def call(self, inputs, **kwargs):
        return inputs


def _softmax(x, axis=-1, name=None):
    x = np.asarray(x, F32)
    e = np.exp(x - x.max(axis=axis, keepdims=True), dtype=F32)
    return T(e / e.sum(axis=axis, keepdims=True, dtype=F32))


def _softmax_xent(labels=None, logits=None, **kwargs):
    """tf.nn.softmax_cross_entropy_with_logits_v2: -sum(labels * log_softmax(logits)) over the last axis."""
    z = np.asarray(logits, F32)
    z = z - z.max(axis=-1, keepdims=True)
    logp = z - np.log(np.exp(z, dtype=F32).sum(axis=-1, keepdims=True, dtype=F32), dtype=F32)
    return T(-(np.asarray(labels, F32) * logp).sum(axis=-1, dtype=F32))


def install_model_level():
    """install() plus the model-file symbols.  Returns the stand-in module."""
    tf = install()
    tf.bool = np.bool_
    tf.placeholder = lambda dtype, shape=None, name=None: _Placeholder(dtype, shape, name)
    tf.sparse_placeholder = lambda dtype, shape=None, name=None: _Placeholder(dtype, shape, name, sparse=True)
    tf.disable_v2_behavior = lambda: None
    tf.sigmoid = lambda x, name=None: T(_ACT["sigmoid"](np.asarray(x, F32)))
    tf.tanh = lambda x, name=None: T(_ACT["tanh"](np.asarray(x, F32)))
    tf.reduce_mean = lambda x, axis=None, keepdims=False: T(np.mean(np.asarray(x, F32), axis=axis, keepdims=keepdims, dtype=F32))
    tf.cast = lambda x, dtype: np.asarray(x).astype(dtype)
    tf.equal = lambda a, b: np.asarray(a) == np.asarray(b)
    tf.argmax = lambda x, axis=None, **k: np.argmax(np.asarray(x), axis=axis)
    tf.nn.softmax = _softmax
    tf.nn.sigmoid = tf.sigmoid
    tf.nn.tanh = tf.tanh
    tf.nn.softmax_cross_entropy_with_logits_v2 = _softmax_xent
    tf.nn.softmax_cross_entropy_with_logits = _softmax_xent
    klayers = sys.modules["tensorflow.python.keras.layers"]
    klayers.Dropout = Dropout
    contrib = types.ModuleType("tensorflow.contrib")
    ckeras = types.ModuleType("tensorflow.contrib.keras")
    ckeras.layers = klayers
    contrib.keras = ckeras
    tf.contrib = contrib
    sys.modules.update({"tensorflow.contrib": contrib, "tensorflow.contrib.keras": ckeras})
    return tf
