"""Minimal step loop around the graph layers: the part of ``kgcn/core.py`` (CoreModel.fit,
lines 247-281: ``construct_feed`` -> ``sess.run([train_step, cost_sum, metrics])``) that is needed
to measure molecules/sec, for the classifier family of ``example_model/model.py:41-69``:

    L x [GraphConv(d_l) -> act] (-> GraphDense(d) -> act) -> GraphGather -> Dense(label_dim)
    -> softmax cross-entropy * mask -> reduce_mean      (Adam, TF defaults, core.py:121-127)

Everything on the data path is a C-ABI call into ``libkgcn_b200.so`` on preallocated buffers (no
torch ops, no autograd), so one step is a fixed launch sequence that is captured once into a CUDA
graph per resident batch and replayed.  Parameters, gradients and Adam moments live in ONE flat
buffer each, so data-parallel training over molecules needs exactly one all-reduce (NCCL, sum) of
the flat gradient buffer per step; the loss scale ``1/B_global`` is folded
into the head's gradient so the sum over ranks is the global ``reduce_mean`` gradient.
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr
from .csr import BatchedCSR
from .ops import act_id


class NetSpec:
    def __init__(self, feature_dim, conv_dims, n_nodes, channels=1, label_dim=2, dense_dim=None, act="sigmoid"):
        self.feature_dim, self.conv_dims, self.n_nodes = int(feature_dim), [int(d) for d in conv_dims], int(n_nodes)
        self.channels, self.label_dim, self.dense_dim, self.act = int(channels), int(label_dim), dense_dim, act

    def param_shapes(self):
        shapes, f = [], self.feature_dim
        for i, d in enumerate(self.conv_dims):
            shapes.append(("conv%d/kernel" % i, (self.channels, f, d)))
            shapes.append(("conv%d/bias" % i, (self.channels, d)))
            f = d
        if self.dense_dim:
            shapes.append(("graph_dense/kernel", (f, int(self.dense_dim))))
            shapes.append(("graph_dense/bias", (int(self.dense_dim),)))
            f = int(self.dense_dim)
        shapes.append(("dense/kernel", (f, self.label_dim)))
        shapes.append(("dense/bias", (self.label_dim,)))
        return shapes


def pad_features(features, width):
    """[B, N, F] -> [B, N, width] with zero columns (no copy when F == width)."""
    features = np.asarray(features, np.float32)
    if width is None or features.shape[-1] == width:
        return features
    out = np.zeros(features.shape[:-1] + (int(width),), np.float32)
    out[..., :features.shape[-1]] = features
    return out


class PeerExchange:
    """Peer-mapped mailboxes of the ranks of one node (cudaIpc through kgcn_p2p_*): rank r allocates
    ``ll[2][world][n_pad]`` 8-byte slots and every rank maps all of them; the 64-byte handles travel through
    ``torch.distributed.all_gather`` (plumbing).  Consumed by kgcn_reduce_adam_f32."""

    def __init__(self, n_params, rank, world, pg, device):
        import torch.distributed as dist
        self.rank, self.world = rank, world
        n_pad = (n_params + 127) // 128 * 128
        p_own, h = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        check(lib.kgcn_p2p_alloc(2 * world * n_pad * 8, ctypes.byref(p_own), h))
        self.own = p_own.value
        mine = torch.tensor(list(bytes(h)), dtype=torch.uint8, device=device)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine, group=pg)
        self.mapped = []
        g = _lib.P2PGroup()
        g.rank, g.world, g.n_pad = rank, world, n_pad
        for r in range(world):
            if r == rank:
                g.mailbox[r] = self.own
                continue
            hr = (ctypes.c_ubyte * 64).from_buffer_copy(bytes(every[r].cpu().tolist()))
            q = ctypes.c_void_p()
            check(lib.kgcn_p2p_open(hr, ctypes.byref(q)))
            g.mailbox[r] = q.value
            self.mapped.append(q.value)
        self.error = torch.zeros(1, dtype=torch.int32, device=device)
        g.error_flag = self.error.data_ptr()
        self.group = g
        dist.barrier(group=pg)

    def close(self):
        for q in self.mapped:
            lib.kgcn_p2p_close(q)
        self.mapped = []
        if self.own is not None:
            lib.kgcn_p2p_free(self.own)
            self.own = None


class DeviceBatch:
    """Static device buffers for one batch (CUDA-graph friendly: pointers never change)."""

    def __init__(self, csr, features, labels, mask):
        self.csr, self.features, self.labels, self.mask = csr, features, labels, mask

    @classmethod
    def from_host(cls, counts, indices, values, features, labels, n_nodes, mask=None, device="cuda", pad_to=None):
        """``pad_to``: store the features zero-padded to this width (``Trainer.dims[0]``, see Trainer ``pad_features``)."""
        csr = BatchedCSR.from_flat(counts, indices, values, n_nodes, n_nodes, device=device)
        B = features.shape[0]
        features = pad_features(features, pad_to)
        mask = np.ones(B, np.float32) if mask is None else mask
        up = lambda a: torch.as_tensor(np.ascontiguousarray(a, np.float32)).to(device)
        return cls(csr, up(features), up(labels), up(mask))


def _pad32(d):
    return (int(d) + 31) // 32 * 32


class Trainer:
    """``pad_features`` (default on): feature widths that are not multiples of 32 (Tox21: 75 -> 50 -> 50 -> 50) are stored
    zero-padded to the next multiple of 32 -- features ``[B, N, 96]``, kernels ``[C, 96, 64]`` with zero rows / columns,
    padded activations kept at exact zero by the layer kernel -- so every GraphConv runs on the fused tcgen05 kernels.
    ``views`` / ``gviews`` are the LOGICAL tensors (strided views into the padded storage): checkpoints, tests and the
    oracle never see the padding.  ``p2p`` (data parallel only, default on): gradients are all-reduced inside the
    optimizer launch over NVLink peer memory (kgcn_reduce_adam_f32); ``p2p=False`` keeps the NCCL all-reduce between
    two graph replays as the cross-check."""

    def __init__(self, spec, batch_size, device="cuda", lr=0.01, world_size=1, seed=1234, flags=_lib.FLAG_DEFAULT,
                 process_group=None, pad_features=True, p2p=None, rank=0):
        self.spec, self.B, self.device, self.lr = spec, int(batch_size), torch.device(device), float(lr)
        self.world_size, self.flags, self.pg, self.rank = int(world_size), flags, process_group, int(rank)
        self.act = act_id(spec.act)
        s, B, N, C = spec, self.B, spec.n_nodes, spec.channels
        # ---- padded widths: only when the fused layer kernel then takes every GraphConv of the network ----
        ldims = [s.feature_dim] + s.conv_dims
        pdims = [_pad32(d) for d in ldims]
        ref_order = bool(flags & _lib.FLAG_REFERENCE_ORDER)
        fused_ok = lambda dims: all(lib.kgcn_graphconv_fwd_fused(B, C, N, dims[i], dims[i + 1]) for i in range(len(s.conv_dims)))
        self.padded = bool(pad_features) and not ref_order and pdims != ldims and fused_ok(pdims)
        self.dims = pdims if self.padded else ldims          # what the kernels see
        self.ldims = ldims
        lshapes = spec.param_shapes()
        pshape = {}
        for i in range(len(s.conv_dims)):
            pshape["conv%d/kernel" % i] = (C, self.dims[i], self.dims[i + 1])
            pshape["conv%d/bias" % i] = (C, self.dims[i + 1])
        f_last = self.dims[-1]
        if s.dense_dim:
            pshape["graph_dense/kernel"] = (f_last, int(s.dense_dim))
            f_last = int(s.dense_dim)
        pshape["dense/kernel"] = (f_last, s.label_dim)
        self.f_head = f_last
        # 16-byte align every tensor inside the flat buffers
        offs, total = {}, 0
        for name, shape in lshapes:
            shape = pshape.get(name, shape)
            offs[name] = total
            total += (int(np.prod(shape)) + 3) // 4 * 4
        self.n_params, self.offs = total, offs
        f32 = dict(dtype=torch.float32, device=self.device)
        self.params = torch.zeros(total, **f32)
        self.grads = torch.zeros(total, **f32)
        self.adam_m = torch.zeros(total, **f32)
        self.adam_v = torch.zeros(total, **f32)
        self.step_state = torch.zeros(2, dtype=torch.int32, device=self.device)
        self.views, self.gviews, self.mviews, self.vviews = {}, {}, {}, {}   # logical tensors
        self.pviews, self.pgviews = {}, {}                                   # padded storage (what the C ABI is given)
        for name, lshape in lshapes:
            shape = pshape.get(name, lshape)
            off, n = offs[name], int(np.prod(shape))
            sl = tuple(slice(0, d) for d in lshape)
            for flat, logical, padded in ((self.params, self.views, self.pviews), (self.grads, self.gviews, self.pgviews),
                                          (self.adam_m, self.mviews, None), (self.adam_v, self.vviews, None)):
                full = flat[off:off + n].view(shape)
                logical[name] = full[sl]
                if padded is not None:
                    padded[name] = full
        self.init_params(np.random.default_rng(seed))

        dims = self.dims + ([int(s.dense_dim)] if s.dense_dim else [])
        self.acts = [None] + [torch.empty(B, N, d, **f32) for d in dims[1:]]
        self.dact = [torch.empty(B, N, max(dims), **f32) for _ in range(2)]
        self.gathered = torch.empty(B, dims[-1], **f32)
        self.dgathered = torch.empty(B, dims[-1], **f32)
        self.logits = torch.empty(B, s.label_dim, **f32)
        self.prediction = torch.empty(B, s.label_dim, **f32)
        self.dlogits = torch.empty(B, s.label_dim, **f32)
        self.stats = torch.zeros(4, **f32)   # [cost_sum, correct_count, block ticket, reserved]
        ws = 0
        for i in range(len(s.conv_dims)):
            ws = max(ws, int(lib.kgcn_graphconv_workspace_bytes(B, C, N, self.dims[i], self.dims[i + 1])))
        if s.dense_dim:
            ws = max(ws, int(lib.kgcn_graphdense_workspace_bytes(B, N, self.dims[-1], int(s.dense_dim))))
        ws = max(ws, int(lib.kgcn_readout_workspace_bytes(B, dims[-1], s.label_dim)))
        self.ws = torch.empty(max(ws, 16), dtype=torch.uint8, device=self.device)
        # ---- fused step: head emits the last layer's dU, each layer's backward leaves weight-gradient partials, one
        # tail launch reduces them, all-reduces across ranks and applies Adam ----
        self.splits = [int(lib.kgcn_graphconv_bwd_splits(B, C, N, self.dims[i], self.dims[i + 1], 1 if i > 0 else 0))
                       for i in range(len(s.conv_dims))]
        self.fused_step = (not ref_order and not s.dense_dim and all(n > 0 for n in self.splits) and fused_ok(self.dims) and
                           os.environ.get("KGCN_FUSED_STEP", "1") != "0")
        self.partials, self._segments = [], None
        L = len(s.conv_dims)
        self._dims_c = (ctypes.c_int32 * (L + 1))(*self.dims)
        self._ldims_c = (ctypes.c_int32 * (L + 1))(*self.ldims)
        # chained launches: all forward layers in one launch, all dx layers in one launch (CTA-local layer-to-layer hand-off)
        self.chain = (self.fused_step and L <= 4 and os.environ.get("KGCN_CHAIN", "1") != "0" and
                      bool(lib.kgcn_graphconv_chain_supported(B, C, N, L, self._dims_c)))
        self.du = [torch.empty(B, N, self.dims[i + 1], **f32) for i in range(L)] if self.chain else None
        if self.chain:
            arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() if t is not None else None for t in ts])
            self._w_ptrs = arr([self.pviews["conv%d/kernel" % i] for i in range(L)])
            self._b_ptrs = arr([self.pviews["conv%d/bias" % i] for i in range(L)])
            self._y_ptrs = arr(self.acts[1:L + 1])
            self._du_ptrs = arr(self.du)
        # the graph-local part of the step (forward layers + fused readout head + dx chain) as ONE launch
        self.step_grid = 0
        want_p2p = p2p if p2p is not None else os.environ.get("KGCN_P2P", "1") != "0"
        if self.chain and os.environ.get("KGCN_STEP_CHAIN", "1") != "0" and (self.world_size == 1 or want_p2p):
            self.step_grid = int(lib.kgcn_gcn_step_chain_grid(B, C, N, L, self._dims_c, s.label_dim))
        self.step_chain = self.step_grid > 0
        # the dx jobs of the step launch also store what they aggregate, G_l = A^T . dU_l, and the weight-gradient launch reads it
        # instead of gathering it a second time (bit-identical; KGCN_GSAVE=0 is the A-B knob).  One adjacency channel only by default:
        # with C channels G is C times as wide as dU, the copying jobs need their own (larger) stage layout and the dx jobs C times
        # the stores -- measured slower on config 4 (216 vs 204 us per step); KGCN_GSAVE=2 forces it for any C
        self.g_save = None
        g_mode = os.environ.get("KGCN_GSAVE", "1")
        if (self.step_chain and g_mode != "0" and (C == 1 or g_mode == "2") and
                bool(lib.kgcn_gcn_step_chain_g_supported(B, C, N, L, self._dims_c))):
            self.g_save = [None] + [torch.empty(B, N, C * self.dims[i + 1], **f32) for i in range(1, L)]
            self._g_ptrs = arr(self.g_save)
        if self.fused_step:
            n_seg = len(s.conv_dims) + (2 if self.step_chain else 0)
            segs = (_lib.GradSegment * n_seg)()
            for i, n in enumerate(self.splits):
                buf = torch.empty(n * (self.dims[i] + 1) * C * self.dims[i + 1], **f32)
                self.partials.append(buf)
                segs[i] = _lib.GradSegment(offs["conv%d/kernel" % i], offs["conv%d/bias" % i], buf.data_ptr(), n, self.dims[i],
                                           self.dims[i + 1], C, 0)
            if self.step_chain:
                # per-CTA head partials {dW_dense [F * L] | db_dense (4) | cost_sum, correct_count, 0, 0}
                FL = self.f_head * s.label_dim
                self.head_partial = torch.zeros(self.step_grid * (FL + 8), **f32)
                hp = self.head_partial.data_ptr()
                segs[L] = _lib.GradSegment(0, offs["dense/kernel"], hp, self.step_grid, 0, FL, 1, FL + 8)
                segs[L + 1] = _lib.GradSegment(0, offs["dense/bias"], hp + 4 * FL, self.step_grid, 0, 4, 1, FL + 8)
                self._stats_partial = hp + 4 * (FL + 4)
                self._stats_stride = FL + 8
            self._segments = segs
            self._part_ptrs = (ctypes.c_void_p * L)(*[t.data_ptr() for t in self.partials])
            self._part_bytes = (ctypes.c_size_t * L)(*[t.numel() * 4 for t in self.partials])
        self.p2p = None
        if self.world_size > 1 and (p2p if p2p is not None else os.environ.get("KGCN_P2P", "1") != "0"):
            self.p2p = PeerExchange(self.n_params, self.rank, self.world_size, self.pg, self.device)
        self.graphs = {}
        self.launches_per_step = None

    # -- parameters -----------------------------------------------------------------------------
    def init_params(self, rng):
        """glorot_uniform kernels, zero biases (kgcn/layers.py:54-61; Keras Dense defaults)."""
        for name, view in self.views.items():
            if name.endswith("kernel"):
                fan_in, fan_out = view.shape[-2], view.shape[-1]
                lim = math.sqrt(6.0 / (fan_in + fan_out))
                view.copy_(torch.as_tensor(rng.uniform(-lim, lim, size=tuple(view.shape)).astype(np.float32)))
            else:
                view.zero_()

    def load_oracle_params(self, p):
        """Load the dict layout of oracle.ref_layers.init_network (tests / smoke only)."""
        t = lambda a: torch.as_tensor(np.asarray(a, np.float32)).to(self.device)
        for i in range(len(self.spec.conv_dims)):
            self.views["conv%d/kernel" % i].copy_(t(np.stack(p["conv_w"][i])))
            self.views["conv%d/bias" % i].copy_(t(np.concatenate(p["conv_b"][i], 0)))
        if self.spec.dense_dim:
            self.views["graph_dense/kernel"].copy_(t(p["gd_w"]))
            self.views["graph_dense/bias"].copy_(t(p["gd_b"]))
        self.views["dense/kernel"].copy_(t(p["out_w"]))
        self.views["dense/bias"].copy_(t(p["out_b"]))

    # -- checkpoints (tf.train.Saver, kgcn/core.py) ------------------------------------------------
    BETA1, BETA2 = 0.9, 0.999

    def _tf_names(self, scope=""):
        """(our name, channel or None, TensorFlow variable name) for every variable, as TensorFlow names them when
        example_model/model.py builds this network: ``graph_conv[_i]/kernel{c}`` ``[F_in, F_out]``, ``bias{c}``
        ``[1, F_out]`` (kgcn/layers.py:54-61), ``graph_dense/{kernel,bias}``, ``dense/{kernel,bias}``."""
        pre = scope.rstrip("/") + "/" if scope else ""
        out = []
        for i in range(len(self.spec.conv_dims)):
            layer = pre + ("graph_conv" if i == 0 else "graph_conv_%d" % i)
            for c in range(self.spec.channels):
                out.append(("conv%d/kernel" % i, c, "%s/kernel%d" % (layer, c)))
                out.append(("conv%d/bias" % i, c, "%s/bias%d" % (layer, c)))
        if self.spec.dense_dim:
            out += [("graph_dense/kernel", None, pre + "graph_dense/kernel"), ("graph_dense/bias", None, pre + "graph_dense/bias")]
        out += [("dense/kernel", None, pre + "dense/kernel"), ("dense/bias", None, pre + "dense/bias")]
        return out

    def steps_done(self):
        return int(self.step_state[0].item())

    def save_checkpoint(self, prefix, scope="", with_slots=True):
        """``saver.save(sess, prefix)``: a TensorFlow V2 checkpoint (kgcn_b200/tf_checkpoint.py) with TensorFlow's
        variable names.  ``with_slots`` adds what tf.train.AdamOptimizer keeps -- ``<var>/Adam`` (m), ``<var>/Adam_1``
        (v), ``beta1_power`` / ``beta2_power`` (= beta^(t+1) after t steps) -- plus ``global_step``, so training resumes
        exactly where it stopped."""
        from .tf_checkpoint import save_checkpoint
        tensors = {}
        for ours, c, tf_name in self._tf_names(scope):
            for views, suffix in ((self.views, ""),) + (((self.mviews, "/Adam"), (self.vviews, "/Adam_1")) if with_slots else ()):
                a = views[ours].detach().cpu().numpy()
                if c is not None:
                    a = a[c].reshape(1, -1) if ours.endswith("bias") else a[c]
                tensors[tf_name + suffix] = np.ascontiguousarray(a)
        if with_slots:
            t = self.steps_done()
            tensors["beta1_power"] = np.array(np.float32(self.BETA1) ** np.float32(t + 1), np.float32)
            tensors["beta2_power"] = np.array(np.float32(self.BETA2) ** np.float32(t + 1), np.float32)
            tensors["global_step"] = np.array(t, np.int64)
        save_checkpoint(prefix, tensors)

    def load_checkpoint(self, prefix, scope="", with_slots=True):
        """``saver.restore(sess, prefix)``.  A missing variable is a KeyError (TF: NotFoundError).  Optimizer state is
        restored when the file has it; the step count comes from ``global_step``, else from ``beta2_power``."""
        from .tf_checkpoint import load_checkpoint
        reader = load_checkpoint(prefix)
        have_slots = with_slots
        for ours, c, tf_name in self._tf_names(scope):
            for views, suffix, required in ((self.views, "", True), (self.mviews, "/Adam", False), (self.vviews, "/Adam_1", False)):
                if suffix and not with_slots:
                    continue
                if not reader.has_tensor(tf_name + suffix):
                    if required:
                        raise KeyError("%s: no variable named %r" % (prefix, tf_name + suffix))
                    have_slots = False
                    continue
                a = torch.as_tensor(np.ascontiguousarray(reader.get_tensor(tf_name + suffix), np.float32))
                dst = views[ours] if c is None else views[ours][c]
                if a.numel() != dst.numel():
                    raise ValueError("%s: %s has shape %r, the network needs %r" % (prefix, tf_name + suffix, tuple(a.shape), tuple(dst.shape)))
                dst.copy_(a.reshape(dst.shape))
        t = 0
        if with_slots and have_slots:
            if reader.has_tensor("global_step"):
                t = int(reader.get_tensor("global_step"))
            elif reader.has_tensor("beta2_power"):
                t = max(0, int(round(math.log(float(reader.get_tensor("beta2_power"))) / math.log(self.BETA2))) - 1)
        else:
            self.adam_m.zero_()
            self.adam_v.zero_()
        self.step_state.copy_(torch.tensor([t, 0], dtype=torch.int32))
        return t

    # -- one step, eager launch sequence --------------------------------------------------------
    def _forward(self, batch, st):
        s, B, N, C = self.spec, self.B, self.spec.n_nodes, self.spec.channels
        csr = batch.csr
        x = batch.features
        if x.shape[-1] != self.dims[0]:
            raise ValueError("batch features are %d wide, the trainer stores %d (DeviceBatch.from_host(..., pad_to=trainer.dims[0]))"
                             % (x.shape[-1], self.dims[0]))
        self.acts[0] = x
        if self.chain:
            L = len(s.conv_dims)
            check(lib.kgcn_graphconv_chain_fwd_f32(ptr(csr.rowptr), ptr(csr.col), ptr(csr.val), B, C, N, L, self._dims_c, self._ldims_c,
                                                   ptr(x), self._w_ptrs, self._b_ptrs, self._y_ptrs, self.act, st))
            self._last_nodes = self.acts[L]
            return self.dims[-1], 0
        for i in range(len(s.conv_dims)):
            y, f, d = self.acts[i + 1], self.dims[i], self.dims[i + 1]
            w, b = self.pviews["conv%d/kernel" % i], self.pviews["conv%d/bias" % i]
            if self.padded:
                check(lib.kgcn_graphconv_fwd_padded_f32(ptr(csr.rowptr), ptr(csr.col), ptr(csr.val), B, C, N, ptr(x), f, ptr(w), ptr(b),
                                                        d, self.ldims[i + 1], self.act, ptr(y), st))
            else:
                check(lib.kgcn_graphconv_fwd_f32(ptr(csr.rowptr), ptr(csr.col), ptr(csr.val), B, C, N, ptr(x), f, ptr(w), ptr(b),
                                                 d, self.act, ptr(y), self.flags, ptr(self.ws), self.ws.numel(), st))
            x = y
        f = self.dims[-1]
        if s.dense_dim:
            y = self.acts[len(s.conv_dims) + 1]
            check(lib.kgcn_graphdense_fwd_f32(ptr(x), B, N, f, ptr(self.pviews["graph_dense/kernel"]),
                                              ptr(self.pviews["graph_dense/bias"]), int(s.dense_dim), self.act, None,
                                              ptr(y), st))
            x, f = y, int(s.dense_dim)
        self._last_nodes = x      # GraphGather is fused into the readout head (kgcn_gather_readout_xent_f32)
        return f, 0

    def _head(self, batch, f, st, train):
        s, B = self.spec, self.B
        inv_batch = 1.0 / (B * self.world_size)
        args = (ptr(self._last_nodes), B, s.n_nodes, f, ptr(self.gathered), ptr(self.pviews["dense/kernel"]),
                ptr(self.pviews["dense/bias"]), s.label_dim, ptr(batch.labels), ptr(batch.mask), inv_batch, ptr(self.logits),
                ptr(self.prediction), ptr(self.stats), ptr(self.dlogits) if train else None, ptr(self.dgathered) if train else None,
                ptr(self.pgviews["dense/kernel"]) if train else None, ptr(self.pgviews["dense/bias"]) if train else None)
        if train and self.fused_step:   # + dU of the last graph layer (dgathered broadcast over the nodes, times act')
            du_top = self.du[-1] if self.chain else self.dact[0]
            check(lib.kgcn_gather_readout_xent_du_f32(*args, self.act, ptr(du_top), ptr(self.ws), self.ws.numel(), st))
        else:
            check(lib.kgcn_gather_readout_xent_f32(*args, ptr(self.ws), self.ws.numel(), st))

    def _backward_fused(self, batch, st):
        """dU chain: the head wrote dU of the last layer into dact[0]; every layer's dx launch multiplies by the activation
        gradient of the layer below, so dact[cur] always holds a ready dU; weight gradients stay as per-CTA partials."""
        s, B, N, C = self.spec, self.B, self.spec.n_nodes, self.spec.channels
        csr = batch.csr
        if self.chain:
            L = len(s.conv_dims)
            if L > 1:   # du[L-1] (head) -> du[L-2] -> .. -> du[0], one launch
                x_ptrs = (ctypes.c_void_p * L)(*[a.data_ptr() for a in self.acts[:L]])
                check(lib.kgcn_graphconv_chain_dx_f32(ptr(csr.rowptr_t), ptr(csr.col_t), ptr(csr.val_t), B, C, N, L, self._dims_c,
                                                      x_ptrs, self._w_ptrs, self._du_ptrs, self.act, st))
            x_ptrs = (ctypes.c_void_p * L)(*[a.data_ptr() for a in self.acts[:L]])
            check(lib.kgcn_graphconv_chain_dw_f32(ptr(csr.rowptr_t), ptr(csr.col_t), ptr(csr.val_t), B, C, N, L, self._dims_c, x_ptrs,
                                                  self._du_ptrs, self._part_ptrs, self._part_bytes, st))
            return
        cur = 0
        for i in range(len(s.conv_dims) - 1, -1, -1):
            fin, fout = self.dims[i], self.dims[i + 1]
            du = self.dact[cur].view(-1)[:B * N * fout]
            dx = None
            if i > 0:
                cur ^= 1
                dx = self.dact[cur].view(-1)[:B * N * fin]
            part = self.partials[i]
            check(lib.kgcn_graphconv_bwd_partial_f32(ptr(csr.rowptr_t), ptr(csr.col_t), ptr(csr.val_t), B, C, N, ptr(self.acts[i]), fin,
                                                     ptr(self.pviews["conv%d/kernel" % i]), fout, ptr(du), ptr(dx), self.act,
                                                     ptr(part), part.numel() * 4, st))

    def _backward(self, batch, f, st):
        s, B, N, C = self.spec, self.B, self.spec.n_nodes, self.spec.channels
        csr = batch.csr
        n_conv = len(s.conv_dims)
        k = n_conv + (1 if s.dense_dim else 0)
        cur = 0
        bcast = 0
        if s.dense_dim:
            dy = self.dact[0].view(-1)[:B * N * f].view(B, N, f)
            check(lib.kgcn_gather_bwd_f32(ptr(self.dgathered), B, N, f, ptr(dy), st))
            fin = self.dims[-1]
            dx = self.dact[1].view(-1)[:B * N * fin].view(B, N, fin)
            check(lib.kgcn_graphdense_bwd_f32(ptr(self.acts[k - 1]), B, N, fin, ptr(self.pviews["graph_dense/kernel"]),
                                              int(s.dense_dim), self.act, None, ptr(self.acts[k]), ptr(dy), ptr(dx),
                                              ptr(self.pgviews["graph_dense/kernel"]), ptr(self.pgviews["graph_dense/bias"]),
                                              ptr(self.ws), self.ws.numel(), st))
            dy, cur, k, f = dx, 1, k - 1, fin
        else:
            # the GraphGather gradient ([B, F] broadcast over nodes) is consumed directly by the last conv layer
            dy, bcast = self.dgathered, _lib.FLAG_DY_BROADCAST
        for i in range(n_conv - 1, -1, -1):
            fin = self.dims[i]
            dx = None
            if i > 0:
                cur ^= 1
                dx = self.dact[cur].view(-1)[:B * N * fin].view(B, N, fin)
            check(lib.kgcn_graphconv_bwd_f32(ptr(csr.rowptr_t), ptr(csr.col_t), ptr(csr.val_t), B, C, N,
                                             ptr(self.acts[i]), fin, ptr(self.pviews["conv%d/kernel" % i]), f, self.act,
                                             ptr(self.acts[i + 1]), ptr(dy), ptr(dx), ptr(self.pgviews["conv%d/kernel" % i]),
                                             ptr(self.pgviews["conv%d/bias" % i]), self.flags | bcast, ptr(self.ws),
                                             self.ws.numel(), st))
            dy, f, bcast = dx, fin, 0

    def _optimizer(self, st):
        """The step's tail.  Fused step / peer exchange: ONE launch reduces the weight-gradient partials, all-reduces over
        NVLink peer memory and applies Adam; otherwise plain Adam on the (already NCCL-reduced) flat gradient buffer."""
        if self.fused_step or self.p2p is not None:
            n_conv = len(self.spec.conv_dims)
            use_head = self.step_chain and self._head_in_chain
            segs, n_seg = (self._segments, n_conv + (2 if use_head else 0)) if self.fused_step else (None, 0)
            group = ctypes.byref(self.p2p.group) if self.p2p is not None else None
            check(lib.kgcn_reduce_adam_f32(ptr(self.params), ptr(self.grads), ptr(self.adam_m), ptr(self.adam_v), self.n_params,
                                           segs, n_seg, self.lr, 0.9, 0.999, 1e-8, 1.0, ptr(self.step_state), group,
                                           self._stats_partial if use_head else None, self.step_grid if use_head else 0,
                                           self._stats_stride if use_head else 0, ptr(self.stats) if use_head else None, st))
        else:
            check(lib.kgcn_adam_f32(ptr(self.params), ptr(self.grads), ptr(self.adam_m), ptr(self.adam_v), self.n_params,
                                    self.lr, 0.9, 0.999, 1e-8, 1, 1.0, ptr(self.step_state), st))

    def _reduce_only(self, st):
        """Fused step without the update (tests / gradient checks): reduce the partials into ``grads``."""
        s, C = self.spec, self.spec.channels
        for i, n in enumerate(self.splits):
            f, d = self.dims[i], self.dims[i + 1]
            # X^T.[G_0 | G_1 | ..] partial blocks -> channel-major kernel gradient + bias gradient
            check(lib.kgcn_reduce_partials_f32(ptr(self.partials[i]), n, f, d, C, ptr(self.pgviews["conv%d/kernel" % i]),
                                               ptr(self.pgviews["conv%d/bias" % i]), st))

    def _allreduce(self):
        if self.world_size > 1 and self.p2p is None:
            import torch.distributed as dist
            dist.all_reduce(self.grads, op=dist.ReduceOp.SUM, group=self.pg)

    @property
    def single_graph(self):
        """True when the whole step (collective included) is one capturable launch sequence."""
        return self.world_size == 1 or self.p2p is not None

    def forward_eager(self, batch):
        st = torch.cuda.current_stream().cuda_stream
        f, _ = self._forward(batch, st)
        self._head(batch, f, st, train=False)

    def _step_chain(self, batch, st, stable=False):
        """Forward layers + fused readout head + dx chain: ONE launch; then the weight-gradient jobs.  ``stable``: the
        batch was complete before the previous kernel on this stream was launched (KGCN_FLAG_INPUTS_STABLE)."""
        s, B, N, C = self.spec, self.B, self.spec.n_nodes, self.spec.channels
        csr, L = batch.csr, len(self.spec.conv_dims)
        x = batch.features
        if x.shape[-1] != self.dims[0]:
            raise ValueError("batch features are %d wide, the trainer stores %d" % (x.shape[-1], self.dims[0]))
        self.acts[0] = x
        self._launch_step_chain(batch, st, stable)
        self._launch_dw_chain(batch, st)

    def _launch_step_chain(self, batch, st, stable=False):
        s, B, N, C = self.spec, self.B, self.spec.n_nodes, self.spec.channels
        csr, L, x = batch.csr, len(self.spec.conv_dims), batch.features
        check(lib.kgcn_gcn_step_chain_g_f32(ptr(csr.rowptr), ptr(csr.col), ptr(csr.val), ptr(csr.rowptr_t), ptr(csr.col_t), ptr(csr.val_t),
                                            B, C, N, L, self._dims_c, self._ldims_c, ptr(x), self._w_ptrs, self._b_ptrs, self._y_ptrs,
                                            self._du_ptrs, self._g_ptrs if self.g_save is not None else None, self.act,
                                            ptr(self.pviews["dense/kernel"]), ptr(self.pviews["dense/bias"]),
                                            s.label_dim, ptr(batch.labels), ptr(batch.mask), 1.0 / (B * self.world_size), ptr(self.logits),
                                            ptr(self.prediction), ptr(self.gathered), ptr(self.head_partial),
                                            _lib.FLAG_INPUTS_STABLE if stable else _lib.FLAG_DEFAULT, st))

    def _launch_dw_chain(self, batch, st):
        B, N, C = self.B, self.spec.n_nodes, self.spec.channels
        csr, L = batch.csr, len(self.spec.conv_dims)
        self.acts[0] = batch.features
        x_ptrs = (ctypes.c_void_p * L)(*[a.data_ptr() for a in self.acts[:L]])
        # G of the layers above the first comes from the step launch's dx jobs (like du, written by _launch_step_chain for this batch)
        g = self._g_ptrs if self.g_save is not None else None
        check(lib.kgcn_graphconv_chain_dw_g_f32(ptr(csr.rowptr_t), ptr(csr.col_t), ptr(csr.val_t), B, C, N, L, self._dims_c, x_ptrs,
                                                self._du_ptrs, g, self._part_ptrs, self._part_bytes, st))

    _head_in_chain = False

    def _fwd_bwd(self, batch, allow_step_chain=True, stable=False):
        st = torch.cuda.current_stream().cuda_stream
        self._head_in_chain = bool(self.step_chain and allow_step_chain)
        if self._head_in_chain:
            self._step_chain(batch, st, stable)
            return
        f, _ = self._forward(batch, st)
        self._head(batch, f, st, train=True)
        if self.fused_step:
            self._backward_fused(batch, st)
            if not self.single_graph:      # NCCL cross-check path: the all-reduce needs the reduced gradients
                self._reduce_only(st)
        else:
            self._backward(batch, f, st)

    def step_eager(self, batch, apply_update=True, stable=False):
        self._fwd_bwd(batch, allow_step_chain=apply_update, stable=stable)   # without the tail launch the head's gradients need the head kernel
        st = torch.cuda.current_stream().cuda_stream
        if not apply_update:
            if self.fused_step and self.single_graph:
                self._reduce_only(st)
            self._allreduce()
            return
        if self.single_graph:
            self._optimizer(st)
        else:
            self._allreduce()
            if self.fused_step:
                check(lib.kgcn_adam_f32(ptr(self.params), ptr(self.grads), ptr(self.adam_m), ptr(self.adam_v), self.n_params,
                                        self.lr, 0.9, 0.999, 1e-8, 1, 1.0, ptr(self.step_state), st))
            else:
                self._optimizer(st)

    # -- CUDA-graph replay ------------------------------------------------------------------------
    def _capture_fn(self, fn, warm_fn="same"):
        """Warm-up outside capture (lazy cudaFuncSetAttribute calls) runs ``warm_fn`` -- by default ``fn`` itself, for a
        training step the forward + backward WITHOUT the optimizer launch: warm-up must neither move the parameters (the
        number of Adam updates would then depend on how many graphs were captured) nor wait for other ranks."""
        warm_fn = fn if warm_fn == "same" else warm_fn
        if warm_fn is not None:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    warm_fn()
            torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g

    def capture(self, key, batch, train=True):
        """Capture the step for ``batch``'s static buffers; replay with :meth:`replay`.

        The whole step (forward, head, backward, reduce + all-reduce + Adam) is ONE graph -- also data parallel, where the
        gradient exchange is peer-memory loads inside the optimizer launch.  Only the NCCL cross-check path (``p2p=False``)
        splits it: forward+backward graph, eager ``dist.all_reduce``, optimizer graph."""
        if not train:
            self.graphs[key] = (self._capture_fn(lambda: self.forward_eager(batch)),)
        elif self.single_graph:
            self.graphs[key] = (self._capture_fn(lambda: self.step_eager(batch), warm_fn=lambda: self._fwd_bwd(batch)),)
        else:
            if getattr(self, "_opt_graph", None) is None:
                self._opt_graph = self._capture_fn(lambda: self._plain_adam(), warm_fn=None)
            self.graphs[key] = (self._capture_fn(lambda: self._fwd_bwd(batch)), "allreduce", self._opt_graph)
        return self.graphs[key]

    def capture_many(self, key, batches):
        """Several consecutive training steps (one per batch, in order) as ONE graph -- the loop over batches that are
        resident in HBM (a whole epoch fits).  Inside the graph the steps are linked by programmatic
        dependent launch, so step k + 1's first kernel is resident, has loaded and aggregated its first tiles while step
        k's reduce + all-reduce + Adam launch is still running (every step after the first is launched with
        KGCN_FLAG_INPUTS_STABLE), and the per-graph launch gap is paid once per replay instead of once per step."""
        if not self.single_graph:
            raise ValueError("capture_many needs the single-graph step (p2p gradient exchange or one GPU)")
        batches = list(batches)

        def run():
            for k, batch in enumerate(batches):
                self.step_eager(batch, stable=k > 0)

        self.graphs[key] = (self._capture_fn(run, warm_fn=lambda: self._fwd_bwd(batches[0])),)
        return self.graphs[key]

    def _plain_adam(self):
        check(lib.kgcn_adam_f32(ptr(self.params), ptr(self.grads), ptr(self.adam_m), ptr(self.adam_v), self.n_params,
                                self.lr, 0.9, 0.999, 1e-8, 1, 1.0, ptr(self.step_state), torch.cuda.current_stream().cuda_stream))

    def replay(self, key):
        for g in self.graphs[key]:
            if g == "allreduce":
                self._allreduce()
            else:
                g.replay()

    def read_stats(self):
        """(cost_sum, correct_count) of the last step -- a device->host read."""
        s = self.stats.cpu()
        return float(s[0]), float(s[1])


def shard_range(n_items, rank, world_size):
    """Contiguous shard [lo, hi) of rank ``rank`` (global batch split into G contiguous shards)."""
    per = n_items // world_size
    rem = n_items % world_size
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


class _Slot:
    """Device staging + CSR buffers + captured graph for one in-flight host-fed step."""

    def __init__(self, trainer, layout, max_nnz):
        t = trainer
        s, B, N, C = t.spec, t.B, t.spec.n_nodes, t.spec.channels
        dev = t.device
        i32 = dict(dtype=torch.int32, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        # ONE device block per slot: [packed COO / labels / mask | features] -- the step's inputs arrive in a single H2D copy
        self.d_all = torch.zeros(layout["bytes_all"], dtype=torch.uint8, device=dev)
        self.d_packed = self.d_all[:layout["bytes"]]
        v = lambda name, dt: self.d_packed[layout[name][0]:layout[name][1]].view(dt)
        self.d_off, self.d_idx, self.d_val = v("off", torch.int64), v("idx", torch.int32), v("val", torch.float32)
        labels, mask = v("labels", torch.float32).view(B, s.label_dim), v("mask", torch.float32)
        self.d_flag = torch.zeros(1, **i32)
        mk = lambda: (torch.zeros(B * C * N + 1, **i32), torch.zeros(max_nnz, **i32), torch.zeros(max_nnz, **f32))
        rp, col, val = mk()
        rpt, colt, valt = mk()
        feats = self.d_all[layout["feat_off"]:layout["bytes_all"]].view(torch.float32).view(B, N, t.dims[0])
        self.batch = DeviceBatch(BatchedCSR(B, C, N, N, rp, col, val, rpt, colt, valt), feats, labels, mask)
        self.h_stats = torch.zeros(2, dtype=torch.float32).pin_memory()
        self.copied, self.done = torch.cuda.Event(), torch.cuda.Event()
        self.graph = None
        self.busy = False


class HostFedPipeline:
    """End-to-end step from HOST buffers, the call a kGCN user makes per step (feed -> sess.run,
    kgcn/core.py:267-269): pinned host COO + features + labels are copied to static device buffers
    (ONE cudaMemcpyAsync per step: the packed COO / label / mask block and the feature block are one buffer), packed to CSR
    (+ transposed CSR) ON THE DEVICE (kgcn_pack_coo_device), the training step runs, and ``cost_sum`` /
    ``correct_count`` come back to the host.  The device part (2 pack launches + the step) is one CUDA
    graph per slot.  With ``depth`` = 2 slots the copies of step i+1 run on a copy stream while step i
    computes (``submit`` / ``collect``); ``run`` is the simple synchronous form."""

    def __init__(self, trainer, max_nnz, train=True, depth=2):
        t = self.trainer = trainer
        s, B, C = t.spec, t.B, t.spec.channels
        self.max_nnz = int(max_nnz)
        # static layout of the packed block (16-byte aligned sections)
        sizes = [("off", 8 * (B * C + 1)), ("idx", 8 * self.max_nnz), ("val", 4 * self.max_nnz),
                 ("labels", 4 * B * s.label_dim), ("mask", 4 * B)]
        self.layout, pos = {}, 0
        for name, n in sizes:
            self.layout[name] = (pos, pos + n)
            pos = (pos + n + 15) // 16 * 16
        self.layout["bytes"] = pos
        self.layout["feat_off"] = (pos + 255) // 256 * 256
        self.layout["bytes_all"] = self.layout["feat_off"] + 4 * B * s.n_nodes * t.dims[0]
        self.train = train
        self.slots = [_Slot(t, self.layout, self.max_nnz) for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(device=t.device)
        self.compute_stream = torch.cuda.Stream(device=t.device)
        self.n_submitted = self.n_collected = 0

    # -- device work of one slot ----------------------------------------------------------------
    def _pack(self, slot):
        t, b = self.trainer, slot.batch
        st = torch.cuda.current_stream().cuda_stream
        B, N, C = t.B, t.spec.n_nodes, t.spec.channels
        for tr, (rp, col, val) in ((0, (b.csr.rowptr, b.csr.col, b.csr.val)), (1, (b.csr.rowptr_t, b.csr.col_t, b.csr.val_t))):
            check(lib.kgcn_pack_coo_device(B * C, N, N, ptr(slot.d_off), ptr(slot.d_idx), ptr(slot.d_val), tr, ptr(rp),
                                           ptr(col), ptr(val), None, ptr(slot.d_flag), st))

    def _device_part(self, slot):
        self._pack(slot)
        if self.train:
            self.trainer.step_eager(slot.batch)
        else:
            self.trainer.forward_eager(slot.batch)

    def capture(self):
        t = self.trainer
        for slot in self.slots:
            def fb(slot=slot):
                self._pack(slot)
                t._fwd_bwd(slot.batch)
            if self.train and not t.single_graph:   # NCCL cross-check path: the collective stays outside graph capture
                if getattr(t, "_opt_graph", None) is None:
                    t._opt_graph = t._capture_fn(lambda: t._plain_adam(), warm_fn=None)
                slot.graph = (t._capture_fn(fb), "allreduce", t._opt_graph)
            else:
                slot.graph = (t._capture_fn(lambda slot=slot: self._device_part(slot), warm_fn=fb if self.train else "same"),)

    # -- host side ------------------------------------------------------------------------------
    def pin_host_batch(self, counts, indices, values, features, labels, mask=None):
        """Host-side batch in pinned memory: the flat COO exactly as kgcn/feed.py emits it per graph
        and channel (concatenated), labels [B,L] and mask [B] in ONE packed block + fp32 features."""
        counts = np.asarray(counts, np.int64)
        nnz = int(counts.sum())
        if nnz > self.max_nnz:
            raise _lib.KgcnError(1, "batch has %d nnz, pipeline capacity is %d" % (nnz, self.max_nnz))
        B = features.shape[0]
        # range check on the host (TF: InvalidArgumentError at the sparse op; the host packer: KgcnIndexError).  The
        # device packer only raises a flag, so a bad index must never reach it silently.
        idx_chk = np.asarray(indices)
        if idx_chk.size and (int(idx_chk.min()) < 0 or int(idx_chk.max()) >= self.trainer.spec.n_nodes):
            raise _lib.KgcnIndexError(3, "adjacency index out of range [0, %d): min %d, max %d"
                                      % (self.trainer.spec.n_nodes, int(idx_chk.min()), int(idx_chk.max())))
        blob = np.zeros(self.layout["bytes_all"], np.uint8)   # [packed block | features]: one pinned buffer, one copy per step
        packed = blob[:self.layout["bytes"]]
        sec = lambda name, dt: packed[self.layout[name][0]:self.layout[name][1]].view(dt)
        off = sec("off", np.int64)
        off[0] = 0
        np.cumsum(counts.reshape(-1), out=off[1:])
        sec("idx", np.int32)[:2 * nnz] = np.ascontiguousarray(indices, np.int32).reshape(-1)
        sec("val", np.float32)[:nnz] = np.asarray(values, np.float32)
        sec("labels", np.float32)[:] = np.asarray(labels, np.float32).reshape(-1)
        sec("mask", np.float32)[:] = np.ones(B, np.float32) if mask is None else np.asarray(mask, np.float32)
        used = self.layout["idx"][0] + 8 * nnz   # the idx section is copied only up to the entries in use
        feats = blob[self.layout["feat_off"]:].view(np.float32).reshape(B, self.trainer.spec.n_nodes, self.trainer.dims[0])
        feats[...] = pad_features(features, self.trainer.dims[0])                                # padded on the host
        # pinned staging on the GPU's own NUMA node (hostmem: the copies of all ranks of a node otherwise share one socket)
        from . import hostmem
        with hostmem.numa_preferred(hostmem.gpu_numa_node(self.trainer.device.index or 0)):
            pinned = torch.from_numpy(blob).pin_memory()
        return {"blob": pinned, "packed": pinned[:self.layout["bytes"]], "nnz": nnz, "idx_used_end": used,
                "features": pinned[self.layout["feat_off"]:].view(torch.float32).view(feats.shape)}

    def h2d_bytes(self, host):
        return int(host["blob"].numel())

    def submit(self, host):
        """Enqueue one step: H2D on the copy stream, graph replay + D2H of the stats on the compute stream."""
        slot = self.slots[self.n_submitted % len(self.slots)]
        if slot.busy:
            raise RuntimeError("pipeline full: collect() a result before submitting more than %d steps" % len(self.slots))
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(slot.done)          # the slot's previous step has finished with these buffers
            slot.d_all.copy_(host["blob"], non_blocking=True)
            slot.copied.record(self.copy_stream)
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(slot.copied)
            if slot.graph is not None:
                for g in slot.graph:
                    if g == "allreduce":
                        self.trainer._allreduce()
                    else:
                        g.replay()
            else:
                self._device_part(slot)
            slot.h_stats.copy_(self.trainer.stats[:2], non_blocking=True)
            slot.done.record(self.compute_stream)
        slot.busy = True
        self.n_submitted += 1

    def collect(self):
        """Wait for the oldest submitted step; returns its (cost_sum, correct_count)."""
        slot = self.slots[self.n_collected % len(self.slots)]
        slot.done.synchronize()
        slot.busy = False
        self.n_collected += 1
        return float(slot.h_stats[0]), float(slot.h_stats[1])

    def run(self, host):
        """One synchronous end-to-end step."""
        self.submit(host)
        return self.collect()

    def run_many(self, hosts):
        """Pipelined loop over host batches; yields (cost_sum, correct_count) per step, in order."""
        depth = len(self.slots)
        for i, h in enumerate(hosts):
            if i >= depth:
                yield self.collect()
            self.submit(h)
        while self.n_collected < self.n_submitted:
            yield self.collect()
