"""kGCN graph layers on the B200 library -- the host-side mirror of ``kgcn/layers.py``.

Same class names, constructor arguments, call signatures and weight names as the reference
(clinfo/kGCN @ 32328d5) so that model definitions written against ``kgcn.layers`` read the same:

    GraphConv(output_dim, adj_channel_num, initializer='glorot_uniform')(x, adj=adjs)   layers.py:32-119
    GraphDense(output_dim, **dense_kw)(x, enabled_node_nums=None, shape=None, max_node_num=None)  :223-265
    GraphGather()(x)                                                                   :156-167
    GINAggregate(adj_channel_num, initializer='zeros')(x, adj=adjs)                    :400-475
    BatchGraphConv(output_dim, ...)([x, adj])                                          :363-398
    load_bspmm(args)                                                                   :19-29

Tensors are CUDA ``torch.Tensor``s; ``adj`` is either the reference's ``list[B][C]`` of
``(indices, values, dense_shape)`` triples (what kgcn/feed.py:112-126 feeds) or a pre-packed
:class:`kgcn_b200.csr.BatchedCSR`.  All arithmetic happens in ``libkgcn_b200.so``; there is no
CPU fallback (a missing library is an ImportError, a CPU tensor is a KgcnError).
"""
import math
import sys

import torch

from . import _lib, ops
from .csr import BatchedCSR, as_batched_csr

enabled_batched = False
enabled_bspmm = False
enabled_bconv = False


def _active_store():
    mod = sys.modules.get("kgcn_b200.compat.facade")
    return mod.active_store() if mod is not None else None


def load_bspmm(args):
    """Mode switch of kgcn/layers.py:19-29.  The reference enables a plugin branch only when its
    flag is set AND ``./<name>.so`` exists; here all three ops live in ``libkgcn_b200.so``, which
    is always present (importing this package fails otherwise), so the flag alone decides --
    first of batched / bspmm / bconv, as in the reference."""
    global enabled_batched, enabled_bspmm, enabled_bconv
    enabled_batched = enabled_bspmm = enabled_bconv = False
    if getattr(args, "batched", False):
        enabled_batched = True
    elif getattr(args, "bspmm", False):
        enabled_bspmm = True
    elif getattr(args, "bconv", False):
        enabled_bconv = True


def _init_tensor(initializer, shape, fan_in, fan_out, device):
    if callable(initializer):
        t = torch.as_tensor(initializer(shape), dtype=torch.float32)
        return t.reshape(shape).to(device)
    if initializer in ("glorot_uniform", None):
        lim = math.sqrt(6.0 / (fan_in + fan_out))  # Keras glorot_uniform (layers.py:54-57)
        return torch.empty(shape, dtype=torch.float32, device=device).uniform_(-lim, lim)
    if initializer == "zeros":
        return torch.zeros(shape, dtype=torch.float32, device=device)
    if initializer == "ones":
        return torch.ones(shape, dtype=torch.float32, device=device)
    raise ValueError("unsupported initializer %r" % (initializer,))


class PendingActivation(torch.Tensor):
    """Output of ``GraphConv`` / ``GraphDense`` under the ``tensorflow``-named façade while the activation is still open.

    The shipped models write ``layer = GraphConv(...)(x, adj=adjs); layer = tf.sigmoid(layer)`` (example_model/model.py:
    41-46).  Run eagerly that would be one library launch plus a separate elementwise pass.  Under the façade the layer
    returns this placeholder tensor instead (shape / dtype / device are real, no storage); ``torch.sigmoid`` /
    ``relu`` / ``tanh`` applied to it launch the layer ONCE with that activation fused in the kernel's epilogue, and any
    other use launches it with no activation and proceeds on the result.  Outside the façade layers return plain tensors."""

    @staticmethod
    def __new__(cls, shape, like, thunk):
        t = torch.Tensor._make_wrapper_subclass(cls, tuple(int(v) for v in shape), dtype=like.dtype, device=like.device)
        t._thunk, t._values = thunk, {}
        return t

    def materialize(self, act=None):
        """The layer output with ``act`` applied: one fused launch per distinct activation asked for (normally exactly one)."""
        if act not in self._values:
            self._values[act] = _apply_act_torch(self._values[None], act) if None in self._values else self._thunk(act)
        return self._values[act]

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in _PENDING_METADATA:
            return super().__torch_function__(func, types, args, kwargs)
        act = _PENDING_FUSABLE.get(func)
        if act is not None and len(args) == 1 and not any(kwargs.values()) and isinstance(args[0], cls):
            return args[0].materialize(act)
        args, kwargs = torch.utils._pytree.tree_map(_materialized, (args, kwargs))
        return func(*args, **kwargs)


PendingActivation.__torch_dispatch__ = classmethod(   # below the Python API (not reached from the façade): same rule
    lambda cls, func, types, args=(), kwargs=None: func(*torch.utils._pytree.tree_map(_materialized, args),
                                                        **torch.utils._pytree.tree_map(_materialized, kwargs or {})))


def _materialized(t):
    return t.materialize() if isinstance(t, PendingActivation) else t


_PENDING_FUSABLE = {torch.sigmoid: "sigmoid", torch.Tensor.sigmoid: "sigmoid", torch.nn.functional.sigmoid: "sigmoid",
                    torch.relu: "relu", torch.Tensor.relu: "relu", torch.nn.functional.relu: "relu",
                    torch.tanh: "tanh", torch.Tensor.tanh: "tanh", torch.nn.functional.tanh: "tanh"}
_PENDING_METADATA = {torch.Tensor.shape.__get__, torch.Tensor.device.__get__, torch.Tensor.dtype.__get__,
                     torch.Tensor.ndim.__get__, torch.Tensor.is_cuda.__get__, torch.Tensor.dim, torch.Tensor.size,
                     torch.Tensor.numel}


def _shape_of(t):
    """Static shape of one element of a list input: a tensor, a packed batch, or a SparseTensorValue-like triple
    (the block-diagonal adjacency of BatchGraphConv, kgcn/layers.py:388-390)."""
    if isinstance(t, BatchedCSR):
        return (t.n_graphs * t.n_rows, t.n_graphs * t.n_cols)
    if hasattr(t, "shape"):
        return tuple(t.shape)
    dense_shape = getattr(t, "dense_shape", None)
    if dense_shape is None and isinstance(t, (list, tuple)) and len(t) == 3:
        dense_shape = t[2]
    return tuple(int(v) for v in dense_shape)


class Layer(torch.nn.Module):
    """Keras-style lazily built layer: weights are created on the first call from the input shape."""

    def __init__(self, name=None, trainable=True, dtype=None, **kwargs):
        super().__init__()
        self._layer_name = name
        self._store_name = None
        self.trainable = trainable
        self.built = False

    def add_weight(self, name, shape, initializer, trainable=True, fan_in=1, fan_out=1, device="cuda"):
        def make():
            value = _init_tensor(initializer, tuple(shape), fan_in, fan_out, device)
            return torch.nn.Parameter(value, requires_grad=bool(trainable and self.trainable))

        store = _active_store()
        if store is not None:   # eager re-execution of a TF-style build_model: reuse variables by TF-style name
            if self._store_name is None:
                self._store_name = store.scoped(self._layer_name) if self._layer_name else \
                    store.layer_name(getattr(self, "TF_SCOPE_CLASS", None) or type(self).__name__)
            param = store.get(self._store_name + "/" + name, make)
            if tuple(param.shape) != tuple(shape):
                raise ValueError("variable %s/%s has shape %r, layer wants %r" % (self._store_name, name, tuple(param.shape), tuple(shape)))
        else:
            param = make()
        self.register_parameter(name, param)
        return param

    def build(self, input_shape):
        self.built = True

    def forward(self, inputs, *args, **kwargs):
        inputs = [_materialized(t) for t in inputs] if isinstance(inputs, (list, tuple)) else _materialized(inputs)
        if not self.built:
            first = inputs[0] if isinstance(inputs, (list, tuple)) else inputs
            self._build_device = first.device
            self.build([_shape_of(t) for t in inputs] if isinstance(inputs, (list, tuple)) else tuple(inputs.shape))
            self.built = True
        return self.call(inputs, *args, **kwargs)


class GraphConv(Layer):
    """``out[b] = sum_c A[b][c] . (x[b] . kernel_c + bias_c)`` (kgcn/layers.py:105-116).

    ``activation`` (extension, default None = reference behaviour) fuses the ``tf.sigmoid`` /
    ``tf.nn.relu`` that the shipped models apply right after the layer into the same kernel."""

    def __init__(self, output_dim, adj_channel_num, initializer="glorot_uniform", activation=None,
                 reference_order=False, **kwargs):
        super().__init__(**kwargs)
        self.output_dim = int(output_dim)
        self.adj_channel_num = int(adj_channel_num)
        self.initializer = initializer
        self.activation = activation
        self.reference_order = reference_order
        if enabled_bspmm:
            from . import bspmm_call as bspmm
            self.bspmm_obj = bspmm.BatchedSpMM()
        if enabled_bconv:
            from . import bconv_call as bconv
            self.bconv_obj = bconv.BatchedConv()
        if enabled_batched:
            from . import batched_call as batched
            self.bspmdt_obj = batched.BatchedSpMDT()

    def build(self, input_shape):  # input: batch_size x node_num x #inputs
        f_in = int(input_shape[2])
        self.w, self.bias = [], []
        for i in range(self.adj_channel_num):
            self.w.append(self.add_weight("kernel" + str(i), (f_in, self.output_dim), self.initializer,
                                          fan_in=f_in, fan_out=self.output_dim, device=self._build_device))
            self.bias.append(self.add_weight("bias" + str(i), (1, self.output_dim), "zeros", device=self._build_device))

    def _stacked(self):
        if self.adj_channel_num == 1:
            return self.w[0].unsqueeze(0), self.bias[0]
        return torch.stack(list(self.w)), torch.cat(list(self.bias), 0)

    def call(self, inputs, adj=None):
        if adj is None:
            raise ValueError("GraphConv needs adj=")
        flat = None
        if (enabled_bconv or enabled_bspmm or enabled_batched) and _values_require_grad(adj):
            # adjacency attributions (visualization.py): the registered d-values gradient of the plugin ops needs the
            # value tensors on the tape (kgcn/bspmm_call.py:49-54), so this batch is packed with its permutation
            from . import _plugin
            csr, flat, _ = _plugin.pack_sparse_list(adj, inputs.device, nested=True)
        else:
            csr = as_batched_csr(adj, inputs.device)
        act = ops.act_id(self.activation)
        B, N, _ = inputs.shape
        if enabled_bconv or enabled_bspmm or enabled_batched:
            # plugin branches (layers.py:68-104): dense X.W_c + b_c for the whole batch, then ONE
            # batched sparse op.  fw[b][c] of the reference is fw[b, c] here.
            fw = torch.stack([ops.GraphDenseFunction.apply(inputs, self.w[c], self.bias[c].reshape(-1), 0, None)
                              for c in range(self.adj_channel_num)], dim=1)          # [B, C, N, F_out]
            if enabled_bconv:
                out = self.bconv_obj.call_packed(csr, fw, flat)
            elif enabled_bspmm:
                out = self.bspmm_obj.call_packed(csr, fw, flat).sum(dim=1) if self.adj_channel_num > 1 else \
                    self.bspmm_obj.call_packed(csr, fw, flat)[:, 0]
            else:
                out = self.bspmdt_obj.call_packed(csr, fw, flat)
            return _apply_act_torch(out, self.activation)
        w, bias = self._stacked()
        flags = _lib.FLAG_REFERENCE_ORDER if self.reference_order else _lib.FLAG_DEFAULT
        if self.activation is None and _active_store() is not None:   # façade: the model applies tf.sigmoid next
            return PendingActivation((B, N, self.output_dim), inputs,
                                     lambda a: ops.GraphConvFunction.apply(inputs, w, bias, csr, ops.act_id(a), flags))
        return ops.GraphConvFunction.apply(inputs, w, bias, csr, act, flags)

    def compute_output_shape(self, input_shape):
        return input_shape[0], input_shape[1], self.output_dim


def _values_require_grad(adj):
    """True when ``adj`` is the reference's list[B][C] of triples and any ``values`` entry is a tensor on the autograd tape."""
    if isinstance(adj, BatchedCSR):
        return False
    for row in adj:
        for sp in (row if isinstance(row, (list, tuple)) and not hasattr(row, "indices") else [row]):
            v = sp.values if hasattr(sp, "values") and hasattr(sp, "dense_shape") else sp[1]
            if torch.is_tensor(v) and v.requires_grad:
                return True
    return False


def _apply_act_torch(x, activation):
    if activation in (None, "none", "linear"):
        return x
    return {"relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh}[activation](x)


class GraphGather(Layer):
    """``reduce_sum(inputs, axis=1)`` over ALL node rows, padding included (layers.py:163-164)."""

    def call(self, inputs, **kwargs):
        return ops.GraphGatherFunction.apply(inputs)

    def compute_output_shape(self, input_shape):
        return input_shape[0], input_shape[2]


class GraphDense(Layer):
    """Keras ``Dense`` applied to every node row (layers.py:223-265).  Accepted Dense kwargs:
    ``activation`` (None / 'relu' / 'sigmoid' / 'tanh' fused; any other callable is applied after),
    ``use_bias``, ``kernel_initializer``, ``bias_initializer``."""

    def __init__(self, output_dim, activation=None, use_bias=True, kernel_initializer="glorot_uniform",
                 bias_initializer="zeros", **kwargs):
        super().__init__(**kwargs)
        self.output_dim = self.units = int(output_dim)
        self.activation = activation
        self.use_bias = use_bias
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer

    def build(self, input_shape):  # input: batch_size x node_num x #inputs
        self.data_shape = input_shape
        f_in = int(input_shape[-1])
        self.kernel = self.add_weight("kernel", (f_in, self.units), self.kernel_initializer, fan_in=f_in,
                                      fan_out=self.units, device=self._build_device)
        self.bias = self.add_weight("bias", (self.units,), self.bias_initializer, device=self._build_device) \
            if self.use_bias else None

    def call(self, inputs, enabled_node_nums=None, shape=None, max_node_num=None, **kwargs):
        fused = self.activation if (self.activation is None or isinstance(self.activation, str)) else None
        x = inputs
        if shape is not None:
            x = x.reshape(int(shape[0]), int(shape[1]), int(shape[2]))
        elif x.dim() == 2:  # already flattened [B*N, F]
            x = x.unsqueeze(0)
        if enabled_node_nums is not None and not torch.is_tensor(enabled_node_nums):
            enabled_node_nums = torch.as_tensor(enabled_node_nums, dtype=torch.int32, device=x.device)
        if self.activation is None and enabled_node_nums is None and shape is None and inputs.dim() == 3 \
                and _active_store() is not None:                          # façade: GraphDense(50)(h); tf.sigmoid(...)
            kernel, bias = self.kernel, self.bias
            return PendingActivation(tuple(x.shape[:2]) + (self.units,), x,
                                     lambda a: ops.GraphDenseFunction.apply(x, kernel, bias, ops.act_id(a), None))
        out = ops.GraphDenseFunction.apply(x, self.kernel, self.bias, ops.act_id(fused), enabled_node_nums)
        if callable(self.activation):
            out = self.activation(out)
            if enabled_node_nums is not None:  # padded rows stay exactly zero (layers.py:249-253)
                keep = torch.arange(x.shape[1], device=x.device)[None, :] < enabled_node_nums[:, None]
                out = out * keep[:, :, None]
        return out

    def compute_output_shape(self, input_shape):
        return input_shape[0], input_shape[1], self.output_dim


class GraphBatchNormalization(Layer):
    """Batch normalisation over node rows (kgcn/layers.py:170-220), kernels ``kgcn_graph_bn_{fwd,bwd}_f32``.

    [TF-semantics, SURVEY.md Appendix A.10] The reference's Keras layer is called without ``training=`` in graph
    mode, so it follows Keras' learning phase, which kgcn/core.py never feeds: it normalises with the MOVING
    statistics (initial mean 0 / variance 1, never updated) in training and inference alike, i.e.
    ``y = gamma * x / sqrt(1 + 1e-3) + beta`` with trainable ``gamma`` (ones) and ``beta`` (zeros).  That is the
    default here (``batch_statistics=False``).  ``batch_statistics=True`` gives the behaviour of
    ``kgcn/legacy/layers.py:170-218`` (``tf.layers.batch_normalization(training=True)``): statistics of the
    current batch over the enabled rows, moving averages updated with momentum 0.99.  With ``enabled_node_nums``
    only the first ``n_b`` rows of graph ``b`` are normalised and the rest are exact zeros (the extract / split /
    pad sequence of layers.py:202-214)."""

    EPS = 1e-3
    MOMENTUM = 0.99
    # The reference layer owns no variables itself: they belong to the tf.keras BatchNormalization /
    # tf.layers.batch_normalization it instantiates inside call() (layers.py:205,215; legacy/layers.py:202,213), so
    # TensorFlow names them ``[bn_name or batch_normalization(_k)]/{gamma,beta,moving_mean,moving_variance}`` -- the
    # names in the shipped checkpoint (model/reaction/model.best.ckpt.index: rollout/batch_normalization_1/gamma ...).
    TF_SCOPE_CLASS = "BatchNormalization"

    def __init__(self, bn_name=None, batch_statistics=False, **kwargs):
        super().__init__(**kwargs)
        self.bn_name = bn_name
        if bn_name and not self._layer_name:
            self._layer_name = bn_name
        self.batch_statistics = batch_statistics

    def build(self, input_shape):
        f = int(input_shape[-1])
        self.data_shape = input_shape
        self.gamma = self.add_weight("gamma", (f,), "ones", device=self._build_device)
        self.beta = self.add_weight("beta", (f,), "zeros", device=self._build_device)
        # non-trainable variables (saved and restored with the rest, never touched by the optimizer)
        self.moving_mean = self.add_weight("moving_mean", (f,), "zeros", trainable=False, device=self._build_device)
        self.moving_variance = self.add_weight("moving_variance", (f,), "ones", trainable=False, device=self._build_device)

    def call(self, inputs, enabled_node_nums=None, shape=None, max_node_num=None, training=True):
        enabled = None
        if enabled_node_nums is not None:
            enabled = torch.as_tensor(enabled_node_nums, device=inputs.device).reshape(-1).to(torch.int32).contiguous()
        use_batch = bool(self.batch_statistics and training)
        y, mean, var = ops.GraphBatchNormFunction.apply(inputs, self.gamma, self.beta, self.moving_mean, self.moving_variance,
                                                        enabled, self.EPS, 1 if use_batch else 0)
        if use_batch:
            with torch.no_grad():
                self.moving_mean.mul_(self.MOMENTUM).add_(mean, alpha=1.0 - self.MOMENTUM)
                self.moving_variance.mul_(self.MOMENTUM).add_(var, alpha=1.0 - self.MOMENTUM)
        return y

    def compute_output_shape(self, input_shape):
        return input_shape


class GraphMaxPooling(Layer):
    """``y[b,i,:] = sum_c max_j A[b][c][i,j] * x[b,j,:]`` over the densified rows (kgcn/layers.py:122-153): one
    launch of ``kgcn_maxpool_fwd_f32`` instead of B * C * F TensorFlow op chains."""

    def __init__(self, adj_channel_num, **kwargs):
        super().__init__(**kwargs)
        self.adj_channel_num = adj_channel_num

    def call(self, inputs, adj=None):
        csr = as_batched_csr(adj, inputs.device)
        if csr.channels != self.adj_channel_num:
            raise ValueError("GraphMaxPooling built for %d adjacency channels, got %d" % (self.adj_channel_num, csr.channels))
        return ops.GraphMaxPoolFunction.apply(inputs, csr)

    def compute_output_shape(self, input_shape):
        return input_shape


class GINAggregate(Layer):
    """``sum_c (epsilon_c * x[b] + A[b][c] . x[b])`` (layers.py:459-471); no weights besides epsilon."""

    def __init__(self, adj_channel_num, initializer="zeros", **kwargs):
        super().__init__(**kwargs)
        self.adj_channel_num = int(adj_channel_num)
        self.initializer = initializer

    def build(self, input_shape):
        self.epsilon = [self.add_weight("epsilon" + str(i), (), self.initializer, device=self._build_device)
                        for i in range(self.adj_channel_num)]

    def call(self, inputs, adj=None):
        csr = as_batched_csr(adj, inputs.device)
        if csr.channels != self.adj_channel_num:
            raise ValueError("GINAggregate built for %d adjacency channels, got %d" % (self.adj_channel_num, csr.channels))
        return ops.GinAggregateFunction.apply(inputs, torch.stack(list(self.epsilon)).reshape(-1), csr)

    def compute_output_shape(self, input_shape):
        return input_shape


class BatchGraphConv(Layer):
    """Block-diagonal form ``relu(A . (x . W + b))`` with inputs ``[x [sumN, F_in], A]``
    (layers.py:363-398; used by example_model/sparse.py with batch dimension 1)."""

    def __init__(self, output_dim, adj_channel_num=1, initializer="glorot_uniform", input_dim=None, **kwargs):
        super().__init__(**kwargs)
        self.output_dim, self.adj_channel_num = int(output_dim), int(adj_channel_num)
        self.initializer, self.input_dim = initializer, input_dim

    def build(self, input_shape):
        f_in = int(input_shape[0][1]) if self.input_dim is None else int(self.input_dim)
        # the reference loop overwrites self.w / self.bias per channel (layers.py:376-386): one pair survives
        last = self.adj_channel_num - 1
        self.w = self.add_weight("kernel" + str(last), (f_in, self.output_dim), self.initializer, fan_in=f_in,
                                 fan_out=self.output_dim, device=self._build_device)
        self.bias = self.add_weight("bias" + str(last), (self.output_dim,), "zeros", device=self._build_device)

    def call(self, inputs, **kwargs):
        net, adj = inputs[0], inputs[1]
        csr = adj if isinstance(adj, BatchedCSR) else as_batched_csr([[adj]], net.device)
        out = ops.GraphConvFunction.apply(net.unsqueeze(0), self.w.unsqueeze(0), self.bias.reshape(1, -1), csr,
                                          ops.act_id("relu"), _lib.FLAG_DEFAULT)
        return out[0]

    def compute_output_shape(self, input_shape):
        return input_shape[0][0], self.output_dim
