import contextlib
import importlib
import re
import sys
import types

import numpy as np
import torch

from .. import default_model as _default_model
from .. import layers as _layers
from ..feed import SparseTensorValue


# ------------------------------------------------------------------------------------------------
# variable store: TF-style names -> torch Parameters, reused across eager re-executions
# ------------------------------------------------------------------------------------------------
class VariableStore:
    def __init__(self):
        self.params = {}          # "graph_conv_1/kernel0" -> Parameter (trainable or not)
        self.initial_values = {}  # name -> numpy array waiting for its variable (restore before the first run)
        self._counts = {}
        self._scope = []

    def begin_pass(self):
        self._counts = {}

    def scoped(self, name):
        return "/".join(self._scope + [name])

    def layer_name(self, cls_name):
        # Keras' to_snake_case (two passes): GraphConv -> graph_conv, GINAggregate -> gin_aggregate
        base = re.sub(r"([a-z])([A-Z])", r"\1_\2", re.sub(r"(.)([A-Z][a-z0-9]+)", r"\1_\2", cls_name)).lower()
        k = self._counts.get(base, 0)
        self._counts[base] = k + 1
        return self.scoped(base if k == 0 else "%s_%d" % (base, k))

    def get(self, key, make):
        if key not in self.params:
            self.params[key] = make()
            if key in self.initial_values:
                self.assign(key, self.initial_values.pop(key))
        return self.params[key]

    def assign(self, key, value):
        param = self.params[key]
        value = np.asarray(value)
        if tuple(value.shape) != tuple(param.shape):
            raise ValueError("variable %s has shape %r, the checkpoint holds %r" % (key, tuple(param.shape), tuple(value.shape)))
        with torch.no_grad():
            param.copy_(torch.as_tensor(value, dtype=param.dtype))


_STORE = None


def active_store():
    return _STORE


@contextlib.contextmanager
def _use_store(store):
    global _STORE
    prev, _STORE = _STORE, store
    store.begin_pass()
    try:
        yield store
    finally:
        _STORE = prev


# ------------------------------------------------------------------------------------------------
# the tensorflow-named module
# ------------------------------------------------------------------------------------------------
class _DType:
    def __init__(self, name, torch_dtype):
        self.name, self.torch = name, torch_dtype

    def __repr__(self):
        return "tf." + self.name


class Placeholder:
    def __init__(self, dtype, shape=None, name=None, sparse=False):
        self.dtype, self.shape, self.name, self.sparse = dtype, shape, name, sparse

    def __repr__(self):
        return "<%s %s %r>" % ("sparse_placeholder" if self.sparse else "placeholder", self.name, self.shape)


def _t(x):
    return x if torch.is_tensor(x) else torch.as_tensor(x)


def _xent(labels, logits):
    labels = _t(labels).to(logits.dtype)
    return -(labels * torch.log_softmax(logits, dim=-1)).sum(dim=-1)


def _make_tf():
    tf = types.ModuleType("tensorflow")
    tf.__version__ = "1.15.0"
    tf._kgcn_b200_facade = True
    for n, d in (("float32", torch.float32), ("float64", torch.float64), ("int32", torch.int32), ("int64", torch.int64),
                 ("bool", torch.bool)):
        setattr(tf, n, _DType(n, d))
    tf.SparseTensorValue = SparseTensorValue
    tf.placeholder = lambda dtype, shape=None, name=None: Placeholder(dtype, shape, name)
    tf.sparse_placeholder = lambda dtype, shape=None, name=None: Placeholder(dtype, shape, name, sparse=True)
    tf.disable_v2_behavior = lambda: None
    tf.sigmoid = torch.sigmoid
    tf.tanh = torch.tanh
    tf.reshape = lambda x, shape: x.reshape(tuple(int(s) for s in shape))
    tf.matmul = lambda a, b: a @ b
    tf.transpose = lambda x, perm=None: x.permute(*perm) if perm is not None else x.t()
    tf.reduce_mean = lambda x, axis=None, **kw: x.mean() if axis is None else x.mean(dim=axis)
    tf.reduce_sum = lambda x, axis=None, **kw: x.sum() if axis is None else x.sum(dim=axis)
    tf.reduce_all = lambda x, axis=None, **kw: x.all() if axis is None else x.all(dim=axis)
    tf.cast = lambda x, dtype: _t(x).to(dtype.torch if isinstance(dtype, _DType) else dtype)
    tf.equal = lambda a, b: _t(a) == _t(b)
    tf.less = lambda a, b: _t(a) < _t(b)
    tf.argmax = lambda x, axis=None, **kw: _t(x).argmax(dim=axis)
    tf.stop_gradient = lambda x: x.detach()
    tf.concat = lambda values, axis=0, **kw: torch.cat(list(values), dim=axis)
    tf.stack = lambda values, axis=0, **kw: torch.stack(list(values), dim=axis)
    tf.expand_dims = lambda x, axis=-1, **kw: x.unsqueeze(axis)
    tf.shape = lambda input=None, **kw: torch.tensor(tuple((input if input is not None else kw["x"]).shape))
    tf.ones = lambda shape, dtype=None, **kw: torch.ones(tuple(int(s) for s in shape), device="cuda")
    tf.zeros = lambda shape, dtype=None, **kw: torch.zeros(tuple(int(s) for s in shape), device="cuda")
    tf.where = lambda cond, a, b: torch.where(cond, a, b)

    @contextlib.contextmanager
    def variable_scope(name, *a, **kw):
        st = active_store()
        if st is not None:
            st._scope.append(str(name))
        try:
            yield
        finally:
            if st is not None:
                st._scope.pop()
    tf.variable_scope = variable_scope

    nn = types.ModuleType("tensorflow.nn")
    nn.relu, nn.tanh, nn.sigmoid = torch.relu, torch.tanh, torch.sigmoid
    nn.softmax = lambda logits, axis=-1, name=None, **kw: torch.softmax(logits, dim=axis)
    nn.softmax_cross_entropy_with_logits_v2 = lambda labels=None, logits=None, **kw: _xent(labels, logits)
    nn.softmax_cross_entropy_with_logits = lambda labels=None, logits=None, **kw: _xent(labels, logits)
    nn.sigmoid_cross_entropy_with_logits = lambda labels=None, logits=None, **kw: \
        torch.nn.functional.binary_cross_entropy_with_logits(logits, _t(labels).to(logits.dtype), reduction="none")

    def weighted_xent(targets=None, logits=None, pos_weight=None, labels=None, **kw):
        y = _t(targets if targets is not None else labels).to(logits.dtype)
        pw = _t(pos_weight).to(logits.dtype).to(logits.device) if not isinstance(pos_weight, (int, float)) else pos_weight
        return -(pw * y * torch.nn.functional.logsigmoid(logits) + (1 - y) * torch.nn.functional.logsigmoid(-logits))
    nn.weighted_cross_entropy_with_logits = weighted_xent
    tf.nn = nn

    # ---- Keras: the reference subclasses Layer / Dense and instantiates K.layers.Dense / Dropout ----
    klayers = types.ModuleType("tensorflow.keras.layers")
    klayers.Layer = _layers.Layer

    class Dense(_layers.Layer):
        """Keras Dense on [..., F] inputs through the library's GEMM (glorot_uniform kernel, zero bias)."""

        def __init__(self, units, activation=None, use_bias=True, **kw):
            super().__init__(**kw)
            self.units, self.activation, self.use_bias = int(units), activation, use_bias

        def build(self, input_shape):
            f_in = int(input_shape[-1])
            self.kernel = self.add_weight("kernel", (f_in, self.units), "glorot_uniform", fan_in=f_in, fan_out=self.units,
                                          device=self._build_device)
            self.bias = self.add_weight("bias", (self.units,), "zeros", device=self._build_device) if self.use_bias else None

        def call(self, inputs, **kw):
            from .. import ops
            x = inputs.reshape(1, -1, inputs.shape[-1]).contiguous()           # [1, rows, F]: one "graph" of `rows` nodes
            fused = self.activation if (self.activation is None or isinstance(self.activation, str)) else None
            out = ops.GraphDenseFunction.apply(x, self.kernel, self.bias, ops.act_id(fused), None)
            out = out.reshape(tuple(inputs.shape[:-1]) + (self.units,))
            return self.activation(out) if callable(self.activation) else out

    class Dropout(_layers.Layer):
        """Identity: under the reference trainer Keras' learning phase is never fed, so Dropout runs
        in inference mode during training and evaluation alike."""

        def __init__(self, rate=0.0, **kw):
            super().__init__(**kw)
            self.rate = rate

        def call(self, inputs, **kw):
            return inputs

    klayers.Dense, klayers.Dropout = Dense, Dropout
    keras = types.ModuleType("tensorflow.keras")
    keras.layers = klayers
    tf.keras = keras
    contrib = types.ModuleType("tensorflow.contrib")
    contrib.keras = keras
    tf.contrib = contrib
    python = types.ModuleType("tensorflow.python")
    pykeras = types.ModuleType("tensorflow.python.keras")
    pykeras.layers = klayers
    python.keras = pykeras
    tf.python = python
    compat = types.ModuleType("tensorflow.compat")
    compat.v1 = tf
    tf.compat = compat
    mods = {"tensorflow": tf, "tensorflow.nn": nn, "tensorflow.keras": keras, "tensorflow.keras.layers": klayers,
            "tensorflow.contrib": contrib, "tensorflow.contrib.keras": keras, "tensorflow.python": python,
            "tensorflow.python.keras": pykeras, "tensorflow.python.keras.layers": klayers, "tensorflow.compat": compat,
            "tensorflow.compat.v1": tf}
    return mods


def _make_kgcn_alias():
    kgcn = types.ModuleType("kgcn")
    kgcn.__path__ = []          # mark as a package
    kgcn._kgcn_b200_facade = True
    legacy = types.ModuleType("kgcn.legacy")
    legacy.__path__ = []
    # kgcn/legacy/layers.py: same layers, but GraphBatchNormalization is tf.layers.batch_normalization(training=True),
    # i.e. batch statistics (legacy/layers.py:202,213)
    legacy_layers = types.ModuleType("kgcn.legacy.layers")
    legacy_layers.__dict__.update({k: v for k, v in vars(_layers).items() if not k.startswith("__")})

    class GraphBatchNormalization(_layers.GraphBatchNormalization):
        def __init__(self, bn_name=None, **kwargs):
            kwargs.setdefault("batch_statistics", True)
            super().__init__(bn_name=bn_name, **kwargs)

    legacy_layers.GraphBatchNormalization = GraphBatchNormalization
    legacy.layers = legacy_layers
    kgcn.layers, kgcn.default_model, kgcn.legacy = _layers, _default_model, legacy
    return {"kgcn": kgcn, "kgcn.layers": _layers, "kgcn.default_model": _default_model, "kgcn.legacy": legacy,
            "kgcn.legacy.layers": legacy_layers}


_SAVED = {}


def install():
    """Install the façade modules (idempotent).  Refuses to shadow a real TensorFlow."""
    existing = sys.modules.get("tensorflow")
    if existing is not None and not getattr(existing, "_kgcn_b200_facade", False) and not getattr(existing, "_kgcn_b200_stub", False):
        raise RuntimeError("a real `tensorflow` module is already imported; the kgcn_b200 façade will not shadow it")
    mods = dict(_make_tf(), **_make_kgcn_alias())
    for k, m in mods.items():
        if k not in _SAVED:
            _SAVED[k] = sys.modules.get(k)
        sys.modules[k] = m
    return mods["tensorflow"]


def uninstall():
    for k, m in _SAVED.items():
        if m is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = m
    _SAVED.clear()


# ------------------------------------------------------------------------------------------------
# runner: the eager stand-in for CoreModel.build + sess.run (kgcn/core.py:138-166, 267-269)
# ------------------------------------------------------------------------------------------------
class ModelRunner:
    """``spec`` is the reference's ``"model.py"`` config value: ``pkg.module`` or ``pkg.module:Class``
    (gcn.py:135-151).  ``search_path`` is prepended to ``sys.path`` while importing."""

    def __init__(self, spec, info, config, batch_size, search_path=None, device="cuda"):
        install()
        self.info, self.config, self.batch_size, self.device = info, config, int(batch_size), torch.device(device)
        mod_name, _, cls_name = spec.partition(":")
        if search_path:
            sys.path.insert(0, search_path)
        try:
            module = importlib.import_module(mod_name)
        finally:
            if search_path:
                sys.path.remove(search_path)
        self.model = getattr(module, cls_name)() if cls_name else module
        self.store = VariableStore()
        self._restore_strict, self._restored_names = None, set()
        self.placeholders = self.model.build_placeholders(info, config, self.batch_size)

    def parameters(self):
        """Trainable variables (tf.trainable_variables())."""
        return [p for p in self.store.params.values() if p.requires_grad]

    def named_parameters(self):
        return {k: p for k, p in self.store.params.items() if p.requires_grad}

    def named_variables(self):
        """All variables incl. the non-trainable batch-normalisation statistics (tf.global_variables())."""
        return dict(self.store.params)

    def restore(self, prefix, strict=True):
        """``saver.restore(sess, prefix)`` (kgcn/core.py) from a TensorFlow V2 checkpoint, by variable name.

        Variables that already exist are assigned now; the rest are assigned when the first ``run`` creates them
        (eager mode has no build step).  Optimizer slots in the file (``.../Adam``, ``beta1_power``) are ignored.
        ``strict``: a variable of the model that the checkpoint lacks raises KeyError on creation / now
        (TF: NotFoundError).  Returns the names found in the file."""
        from ..tf_checkpoint import load_checkpoint
        values = load_checkpoint(prefix).tensors(skip_slots=True)
        missing = [k for k in self.store.params if k not in values]
        if strict and missing:
            raise KeyError("%s: no value for variable(s) %s" % (prefix, ", ".join(sorted(missing))))
        for k in self.store.params:
            if k in values:
                self.store.assign(k, values[k])
        self.store.initial_values = {k: v for k, v in values.items() if k not in self.store.params}
        self._restore_strict, self._restored_names = (prefix if strict else None), set(values)
        return sorted(values)

    def save(self, prefix):
        """``saver.save(sess, prefix)``: all variables to ``prefix.index`` + ``prefix.data-00000-of-00001``."""
        from ..tf_checkpoint import save_checkpoint
        save_checkpoint(prefix, {k: p.detach().cpu().numpy() for k, p in self.store.params.items()})

    def _bind(self, feed):
        bound = {}
        for key, ph in self.placeholders.items():
            v = feed.get(key)
            if key == "adjs":
                bound[key] = v                      # list[B][C] of triples or a BatchedCSR: GraphConv takes both
            elif v is None or isinstance(v, (bool, float, int)):
                bound[key] = v
            elif torch.is_tensor(v):
                bound[key] = v.to(self.device)
            else:
                a = np.asarray(v)
                t = torch.as_tensor(a)
                if a.dtype == np.float64:
                    t = t.float()
                bound[key] = t.to(self.device)
        return bound

    def run(self, feed):
        """One eager execution of ``build_model`` on this step's feed.  Returns a dict with the five
        values of the model-module protocol: model, prediction, cost_opt, cost_sum, metrics."""
        before = set(self.store.params)
        with _use_store(self.store):
            out = self.model.build_model(self._bind(feed), self.info, self.config, self.batch_size)
        if self._restore_strict:
            fresh = sorted(k for k in set(self.store.params) - before if k not in self._restored_names)
            if fresh:
                raise KeyError("%s: no value for variable(s) %s" % (self._restore_strict, ", ".join(fresh)))
        model, prediction, cost_opt, cost_sum, metrics = out[:5]
        return {"model": model, "prediction": prediction, "cost_opt": cost_opt, "cost_sum": cost_sum, "metrics": metrics}

    def train_step(self, feed, learning_rate=None):
        """``sess.run([train_step, cost_sum, metrics], feed_dict)`` of the fit loop (kgcn/core.py:267-269) with the
        optimizer ``build_optimizer`` creates (core.py:121-127): ``tf.train.AdamOptimizer(learning_rate).minimize(cost_opt)``
        with TensorFlow's defaults (beta1 0.9, beta2 0.999, epsilon 1e-8) and TensorFlow's update formula -- the library's
        ``kgcn_adam_f32`` over each trainable variable with its own ``<var>/Adam`` / ``<var>/Adam_1`` slots.  Returns
        the dict of :meth:`run` (values of the step BEFORE the update, like the fetches of one ``sess.run``)."""
        from .._lib import check, lib, ptr
        out = self.run(feed)
        lr = float(self.config.get("learning_rate", 0.01) if learning_rate is None else learning_rate)
        names = [k for k, p in self.store.params.items() if p.requires_grad]
        params = [self.store.params[k] for k in names]
        grads = torch.autograd.grad(out["cost_opt"], params, allow_unused=True)
        self.adam_step = getattr(self, "adam_step", 0) + 1
        slots = self.__dict__.setdefault("adam_slots", {})
        stream = torch.cuda.current_stream().cuda_stream
        with torch.no_grad():
            for k, p, g in zip(names, params, grads):
                if g is None:                       # TF: variables without a gradient are skipped by minimize()
                    continue
                if k not in slots:
                    slots[k] = (torch.zeros_like(p), torch.zeros_like(p))
                m, v = slots[k]
                check(lib.kgcn_adam_f32(ptr(p), ptr(g.contiguous()), ptr(m), ptr(v), p.numel(), lr, 0.9, 0.999, 1e-8,
                                        self.adam_step, 1.0, None, stream))
        return out
