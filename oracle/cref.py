"""ctypes wrapper of oracle/libkgcn_ref.so (graphconv_ref.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libkgcn_ref.so")
if not os.path.exists(_PATH):
    raise ImportError("oracle/libkgcn_ref.so missing: run `make -C oracle` (or __graft_entry__.build())")
_lib = ctypes.CDLL(_PATH)
_vp, _i64, _i32, _f = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float
_lib.kgcn_ref_graphconv_fwd.argtypes = [_i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32]
_lib.kgcn_ref_train_step.argtypes = [_i64, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp,
                                     _vp, _vp, _i32, _f, _i32, _i32, _vp, _vp, _i32]
_lib.kgcn_ref_max_threads.restype = ctypes.c_int


def max_threads():
    return int(_lib.kgcn_ref_max_threads())


def _off(counts):
    off = np.zeros(counts.size + 1, np.int64)
    np.cumsum(np.asarray(counts, np.int64).reshape(-1), out=off[1:])
    return off


def graphconv_fwd(counts, indices, values, x, w, bias, act=0, n_threads=0):
    """counts [B,C]; indices [nnz,2] int32; x [B,N,Fi]; w [C,Fi,Fo]; bias [C,Fo] -> y [B,N,Fo]."""
    B, N, Fi = x.shape
    C, _, Fo = w.shape
    off = _off(counts)
    idx = np.ascontiguousarray(indices, np.int32)
    val = np.ascontiguousarray(values, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    bias = np.ascontiguousarray(bias, np.float32)
    y = np.empty((B, N, Fo), np.float32)
    rc = _lib.kgcn_ref_graphconv_fwd(B, C, N, Fi, Fo, off.ctypes.data, idx.ctypes.data, val.ctypes.data, x.ctypes.data,
                                     w.ctypes.data, bias.ctypes.data, act, y.ctypes.data, n_threads)
    assert rc == 0
    return y


class RefNet:
    """Flat-buffer network state with the layout of kgcn_b200/trainer.py (tensors padded to 4 floats)."""

    def __init__(self, feature_dim, conv_dims, channels, n_labels, act=2):
        self.dims = np.asarray([feature_dim] + list(conv_dims), np.int32)
        self.C, self.n_labels, self.act = channels, n_labels, act
        pad4 = lambda n: (n + 3) // 4 * 4
        self.offsets, n = {}, 0
        for l in range(len(conv_dims)):
            self.offsets["conv%d/kernel" % l] = (n, (channels, int(self.dims[l]), int(self.dims[l + 1])))
            n += pad4(channels * int(self.dims[l]) * int(self.dims[l + 1]))
            self.offsets["conv%d/bias" % l] = (n, (channels, int(self.dims[l + 1])))
            n += pad4(channels * int(self.dims[l + 1]))
        self.offsets["dense/kernel"] = (n, (int(self.dims[-1]), n_labels))
        n += pad4(int(self.dims[-1]) * n_labels)
        self.offsets["dense/bias"] = (n, (n_labels,))
        n += pad4(n_labels)
        self.params = np.zeros(n, np.float32)
        self.grads = np.zeros(n, np.float32)
        self.m = np.zeros(n, np.float32)
        self.v = np.zeros(n, np.float32)
        self.step = 0

    def view(self, buf, name):
        off, shape = self.offsets[name]
        return buf[off:off + int(np.prod(shape))].reshape(shape)

    def load_oracle_params(self, p):
        for l in range(len(self.dims) - 1):
            self.view(self.params, "conv%d/kernel" % l)[...] = np.stack(p["conv_w"][l])
            self.view(self.params, "conv%d/bias" % l)[...] = np.concatenate(p["conv_b"][l], 0)
        self.view(self.params, "dense/kernel")[...] = p["out_w"]
        self.view(self.params, "dense/bias")[...] = p["out_b"]

    def train_step(self, counts, indices, values, x, labels, mask, n_nodes, lr=0.01, apply_update=True,
                   want_grads=True, inv_batch=None, n_threads=0):
        B = x.shape[0]
        off = _off(counts)
        idx = np.ascontiguousarray(indices, np.int32)
        val = np.ascontiguousarray(values, np.float32)
        x = np.ascontiguousarray(x, np.float32)
        labels = np.ascontiguousarray(labels, np.float32)
        mask = np.ascontiguousarray(mask, np.float32)
        stats = np.zeros(2, np.float32)
        logits = np.zeros((B, self.n_labels), np.float32)
        if apply_update and want_grads:
            self.step += 1
        rc = _lib.kgcn_ref_train_step(B, n_nodes, self.C, len(self.dims) - 1, self.dims.ctypes.data, self.n_labels, self.act,
                                      off.ctypes.data, idx.ctypes.data, val.ctypes.data, x.ctypes.data, labels.ctypes.data,
                                      mask.ctypes.data, 1.0 / B if inv_batch is None else inv_batch, self.params.ctypes.data,
                                      self.grads.ctypes.data, self.m.ctypes.data, self.v.ctypes.data, max(self.step, 1), lr,
                                      int(apply_update), int(want_grads), stats.ctypes.data, logits.ctypes.data, n_threads)
        assert rc == 0
        return stats, logits
