"""Torch-tensor wrappers over the C ABI, with autograd.

PyTorch only supplies device memory, the current stream and the autograd tape; every computation
is a call into ``libkgcn_b200.so`` with raw device pointers (no torch ops on the data path, no
CPU fallback: CPU tensors are rejected).
"""
import torch

from . import _lib
from ._lib import ACT_IDS, FLAG_DEFAULT, check, lib, ptr

_WORKSPACES = {}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(name, t, dtype=torch.float32):
    if not torch.is_tensor(t) or not t.is_cuda:
        raise _lib.KgcnError(6, "%s must be a CUDA tensor: kgcn_b200 has no CPU path" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t.contiguous()


def workspace(nbytes, device):
    """Per-(device, stream) scratch buffer, grown on demand; stream order makes reuse safe."""
    key = (device.index, _stream())
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _WORKSPACES[key] = buf
    return buf


def act_id(act):
    if isinstance(act, int):
        return act
    if act not in ACT_IDS:
        raise ValueError("unsupported fused activation %r (supported: %s)" % (act, sorted(k for k in ACT_IDS if k)))
    return ACT_IDS[act]


# ------------------------------------------------------------------------------------------------
# raw launches
# ------------------------------------------------------------------------------------------------
def bspmm_raw(csr, rhs, rs_g, rs_c, out, os_g, os_c, feat, self_scale=None):
    check(lib.kgcn_bspmm_f32(ptr(csr.rowptr), ptr(csr.col), ptr(csr.val), csr.n_graphs, csr.channels, csr.n_rows,
                             csr.n_cols, feat, ptr(rhs), rs_g, rs_c, ptr(out), os_g, os_c, ptr(self_scale), _stream()))
    return out


LAYOUTS = ("per_matrix", "sum", "shared_sum")


def _strides(layout, C, rows, feat):
    """(stride_g, stride_c) of a [B, C, rows, F] (per-channel) or [B, rows, F] (shared) operand."""
    if layout == "per_channel":
        return C * rows * feat, rows * feat
    return rows * feat, 0


def bspmm(csr, rhs, layout, self_scale=None):
    """layout 'per_matrix': rhs [B,C,K,F] -> [B,C,R,F]; 'sum': rhs [B,C,K,F] -> [B,R,F];
    'shared_sum': rhs [B,K,F] -> [B,R,F] (channels summed, one right-hand side per graph)."""
    rhs = _need_cuda("rhs", rhs)
    B, C, R, K = csr.n_graphs, csr.channels, csr.n_rows, csr.n_cols
    F = rhs.shape[-1]
    if layout == "shared_sum":
        if tuple(rhs.shape) != (B, K, F):
            raise ValueError("rhs must be [%d, %d, F], got %r" % (B, K, tuple(rhs.shape)))
        rs = _strides("shared", C, K, F)
    else:
        if tuple(rhs.shape) != (B, C, K, F):
            raise ValueError("rhs must be [%d, %d, %d, F], got %r" % (B, C, K, tuple(rhs.shape)))
        rs = _strides("per_channel", C, K, F)
    if layout == "per_matrix":
        out = torch.empty((B, C, R, F), dtype=torch.float32, device=rhs.device)
        os_ = _strides("per_channel", C, R, F)
    else:
        out = torch.empty((B, R, F), dtype=torch.float32, device=rhs.device)
        os_ = _strides("shared", C, R, F)
    return bspmm_raw(csr, rhs, rs[0], rs[1], out, os_[0], os_[1], F, self_scale)


class BspmmFunction(torch.autograd.Function):
    """out = A . rhs with the gradients the reference registers (kgcn/bspmm_call.py:21-57,
    bconv_call.py:28-70, batched_call.py:32-75): d rhs = A^T . d out, d values = gather-dot."""

    @staticmethod
    def forward(ctx, rhs, values, csr, layout):
        ctx.csr, ctx.layout = csr, layout
        ctx.save_for_backward(rhs)
        ctx.want_dvalues = values is not None and values.requires_grad
        return bspmm(csr, rhs, layout)

    @staticmethod
    def backward(ctx, dout):
        (rhs,) = ctx.saved_tensors
        csr, layout = ctx.csr, ctx.layout
        dout = _need_cuda("dout", dout)
        B, C, R, K = csr.n_graphs, csr.channels, csr.n_rows, csr.n_cols
        F = rhs.shape[-1]
        t = csr.transposed()
        d_rhs = None
        if ctx.needs_input_grad[0]:
            if layout == "per_matrix":
                d_rhs = bspmm(t, dout, "per_matrix")
            elif layout == "shared_sum":
                d_rhs = bspmm(t, dout, "shared_sum")
            else:  # 'sum': every channel sees the same d out (bconv_call.py:46), outputs stay per channel
                d_rhs = torch.empty((B, C, K, F), dtype=torch.float32, device=rhs.device)
                bspmm_raw(t, dout, R * F, 0, d_rhs, C * K * F, K * F, F)
        d_values = None
        if ctx.want_dvalues:
            if csr.perm is None:
                raise _lib.KgcnError(6, "gradient w.r.t. sparse values needs a BatchedCSR packed with want_perm=True")
            d_values = torch.empty((csr.nnz,), dtype=torch.float32, device=rhs.device)
            ds = _strides("per_channel" if layout == "per_matrix" else "shared", C, R, F)
            rs = _strides("shared" if layout == "shared_sum" else "per_channel", C, K, F)
            check(lib.kgcn_bspmm_dvalues_f32(ptr(csr.rowptr), ptr(csr.col), ptr(csr.perm), B, C, R, K, F, ptr(dout),
                                             ds[0], ds[1], ptr(rhs), rs[0], rs[1], ptr(d_values), _stream()))
        return d_rhs, d_values, None, None


_IDENTITY_CSR = {}


def _identity_rows(B, N, device):
    """rowptr / col of B identity matrices [N, N] (one stored entry per row), cached per shape and device."""
    key = (B, N, device.index)
    if key not in _IDENTITY_CSR:
        rowptr = torch.arange(B * N + 1, dtype=torch.int32, device=device)
        col = torch.arange(N, dtype=torch.int32, device=device).repeat(B).contiguous()
        _IDENTITY_CSR[key] = (rowptr, col)
    return _IDENTITY_CSR[key]


def dot_all(a, b):
    """sum(a * b) over [B, N, F] operands through the library, deterministic: per-row dot products with the
    gather-dot kernel of the sparse-values gradient (kgcn_bspmm_dvalues_f32 on an identity pattern), then node sums
    (kgcn_gather_fwd_f32) folded 32 at a time.  Returns a 0-d device tensor."""
    a, b = _need_cuda("a", a), _need_cuda("b", b)
    B, N, F = a.shape
    rowptr, col = _identity_rows(B, N, a.device)
    rows = torch.empty(B * N, dtype=torch.float32, device=a.device)
    check(lib.kgcn_bspmm_dvalues_f32(ptr(rowptr), ptr(col), None, B, 1, N, N, F, ptr(a), N * F, 0, ptr(b), N * F, 0,
                                     ptr(rows), _stream()))
    cur, n, width = rows, B * N, N
    while n > 1:
        groups = (n + width - 1) // width
        if groups * width != n:                       # zero padding does not change the sum
            padded = torch.zeros(groups * width, dtype=torch.float32, device=a.device)
            padded[:n].copy_(cur[:n])
            cur = padded
        out = torch.empty(groups, dtype=torch.float32, device=a.device)
        check(lib.kgcn_gather_fwd_f32(ptr(cur), groups, width, 1, ptr(out), _stream()))
        cur, n, width = out, groups, 32
    return cur.reshape(())


class GinAggregateFunction(torch.autograd.Function):
    """``y[b] = sum_c (eps_c * x[b] + A[b][c] . x[b])`` (kgcn/layers.py:459-471) in ONE launch: the epsilon term is the
    ``self_scale`` argument of kgcn_bspmm_f32.  Backward: ``dx = sum_c (eps_c * dy + A_c^T . dy)`` is the same launch on
    the transposed batch; ``d eps_c = <x, dy>`` for every channel."""

    @staticmethod
    def forward(ctx, x, eps, csr):
        x, eps = _need_cuda("inputs", x), _need_cuda("epsilon", eps)
        if eps.numel() != csr.channels:
            raise ValueError("GINAggregate has %d epsilons, the adjacency batch %d channels" % (eps.numel(), csr.channels))
        ctx.csr = csr
        ctx.save_for_backward(x, eps)
        return bspmm(csr, x, "shared_sum", self_scale=eps)

    @staticmethod
    def backward(ctx, dout):
        x, eps = ctx.saved_tensors
        dout = _need_cuda("dout", dout)
        d_x = bspmm(ctx.csr.transposed(), dout, "shared_sum", self_scale=eps) if ctx.needs_input_grad[0] else None
        d_eps = dot_all(x, dout).expand(eps.numel()) if ctx.needs_input_grad[1] else None
        return d_x, d_eps, None


def graphconv_workspace_bytes(B, C, N, f_in, f_out):
    return int(lib.kgcn_graphconv_workspace_bytes(B, C, N, f_in, f_out))


def graphconv_fwd(csr, x, w, bias, act=0, flags=FLAG_DEFAULT, out=None):
    B, N, f_in = x.shape
    C, _, f_out = w.shape
    if csr.n_graphs != B or csr.channels != C or csr.n_rows != N or csr.n_cols != N:
        raise ValueError("adjacency batch is %d graphs x %d channels of [%d,%d]; features are %r, weights %r"
                         % (csr.n_graphs, csr.channels, csr.n_rows, csr.n_cols, tuple(x.shape), tuple(w.shape)))
    y = out if out is not None else torch.empty((B, N, f_out), dtype=torch.float32, device=x.device)
    nbytes = graphconv_workspace_bytes(B, C, N, f_in, f_out)
    ws = workspace(nbytes, x.device)
    check(lib.kgcn_graphconv_fwd_f32(ptr(csr.rowptr), ptr(csr.col), ptr(csr.val), B, C, N, ptr(x), f_in, ptr(w),
                                     ptr(bias), f_out, act, ptr(y), flags, ptr(ws), ws.numel(), _stream()))
    return y


def graphconv_bwd(csr, x, w, act, y, dy, need_dx=True, flags=FLAG_DEFAULT):
    B, N, f_in = x.shape
    C, _, f_out = w.shape
    dx = torch.empty_like(x) if need_dx else None
    dw = torch.empty_like(w)
    db = torch.empty((C, f_out), dtype=torch.float32, device=x.device)
    nbytes = graphconv_workspace_bytes(B, C, N, f_in, f_out)
    ws = workspace(nbytes, x.device)
    check(lib.kgcn_graphconv_bwd_f32(ptr(csr.rowptr_t), ptr(csr.col_t), ptr(csr.val_t), B, C, N, ptr(x), f_in, ptr(w),
                                     f_out, act, ptr(y), ptr(dy), ptr(dx), ptr(dw), ptr(db), flags, ptr(ws),
                                     ws.numel(), _stream()))
    return dx, dw, db


class GraphConvFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, csr, act, flags):
        x, w = _need_cuda("inputs", x), _need_cuda("kernel", w)
        bias = _need_cuda("bias", bias) if bias is not None else None
        y = graphconv_fwd(csr, x, w, bias, act, flags)
        ctx.csr, ctx.act, ctx.flags, ctx.has_bias = csr, act, flags, bias is not None
        ctx.save_for_backward(x, w, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = _need_cuda("dy", dy)
        dx, dw, db = graphconv_bwd(ctx.csr, x, w, ctx.act, y, dy, need_dx=ctx.needs_input_grad[0], flags=ctx.flags)
        return dx, dw, (db if ctx.has_bias else None), None, None, None


def graphdense_fwd(x, kernel, bias, act=0, enabled=None):
    B, N, f_in = x.shape
    f_out = kernel.shape[1]
    y = torch.empty((B, N, f_out), dtype=torch.float32, device=x.device)
    check(lib.kgcn_graphdense_fwd_f32(ptr(x), B, N, f_in, ptr(kernel), ptr(bias), f_out, act, ptr(enabled), ptr(y),
                                      _stream()))
    return y


class GraphDenseFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kernel, bias, act, enabled):
        x, kernel = _need_cuda("inputs", x), _need_cuda("kernel", kernel)
        bias = _need_cuda("bias", bias) if bias is not None else None
        if enabled is not None:
            enabled = _need_cuda("enabled_node_nums", enabled, torch.int32)
        y = graphdense_fwd(x, kernel, bias, act, enabled)
        ctx.act, ctx.enabled, ctx.has_bias = act, enabled, bias is not None
        ctx.save_for_backward(x, kernel, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, kernel, y = ctx.saved_tensors
        dy = _need_cuda("dy", dy)
        B, N, f_in = x.shape
        f_out = kernel.shape[1]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dk = torch.empty_like(kernel)
        db = torch.empty((f_out,), dtype=torch.float32, device=x.device)
        nbytes = int(lib.kgcn_graphdense_workspace_bytes(B, N, f_in, f_out))
        ws = workspace(nbytes, x.device)
        check(lib.kgcn_graphdense_bwd_f32(ptr(x), B, N, f_in, ptr(kernel), f_out, ctx.act, ptr(ctx.enabled), ptr(y),
                                          ptr(dy), ptr(dx), ptr(dk), ptr(db), ptr(ws), ws.numel(), _stream()))
        return dx, dk, (db if ctx.has_bias else None), None, None


class GraphGatherFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _need_cuda("inputs", x)
        B, N, F = x.shape
        ctx.shape = (B, N, F)
        out = torch.empty((B, F), dtype=torch.float32, device=x.device)
        check(lib.kgcn_gather_fwd_f32(ptr(x), B, N, F, ptr(out), _stream()))
        return out

    @staticmethod
    def backward(ctx, dout):
        B, N, F = ctx.shape
        dout = _need_cuda("dout", dout)
        dx = torch.empty((B, N, F), dtype=torch.float32, device=dout.device)
        check(lib.kgcn_gather_bwd_f32(ptr(dout), B, N, F, ptr(dx), _stream()))
        return dx


class GraphMaxPoolFunction(torch.autograd.Function):
    """GraphMaxPooling (kgcn/layers.py:122-153) through kgcn_maxpool_{fwd,bwd}_f32."""

    @staticmethod
    def forward(ctx, x, csr):
        x = _need_cuda("inputs", x)
        B, N, F = x.shape
        if csr.n_graphs != B or csr.n_rows != N or csr.n_cols != N:
            raise ValueError("GraphMaxPooling: adjacency batch %s does not match inputs %s" % ((csr.n_graphs, csr.n_rows, csr.n_cols), (B, N, F)))
        y = torch.empty_like(x)
        ws = None
        if ctx.needs_input_grad[0]:   # not x.requires_grad: .contiguous() inside forward returns a no-grad copy
            ws = torch.empty(int(lib.kgcn_maxpool_workspace_bytes(B, csr.channels, N, F)), dtype=torch.uint8, device=x.device)
        check(lib.kgcn_maxpool_fwd_f32(ptr(csr.rowptr), ptr(csr.col), ptr(csr.val), B, csr.channels, N, ptr(x), F, ptr(y),
                                       ptr(ws), 0 if ws is None else ws.numel(), _stream()))
        ctx.csr, ctx.ws = csr, ws
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        csr, ws = ctx.csr, ctx.ws
        if ws is None:
            raise RuntimeError("GraphMaxPooling backward: the forward pass kept no workspace (input did not require grad)")
        B, N, F = x.shape
        dy = _need_cuda("dy", dy)
        dx = torch.empty_like(x)
        check(lib.kgcn_maxpool_bwd_f32(ptr(csr.rowptr_t), ptr(csr.col_t), ptr(csr.val_t), B, csr.channels, N, ptr(x), F, ptr(dy),
                                       ptr(ws), ws.numel(), ptr(dx), _stream()))
        return dx, None


class SegmentSumFunction(torch.autograd.Function):
    """Per-molecule row-range sums of the block-diagonal model (example_model/sparse.py:79-90)."""

    @staticmethod
    def forward(ctx, x, start, size):
        x = _need_cuda("inputs", x)
        start = _need_cuda("start", start, torch.int64)
        size = _need_cuda("size", size, torch.int64)
        rows, F = x.shape
        out = torch.empty((start.numel(), F), dtype=torch.float32, device=x.device)
        check(lib.kgcn_segment_sum_fwd_f32(ptr(x), ptr(start), ptr(size), start.numel(), F, ptr(out), _stream()))
        ctx.rows = rows
        ctx.save_for_backward(start, size)
        return out

    @staticmethod
    def backward(ctx, dout):
        start, size = ctx.saved_tensors
        dout = _need_cuda("dout", dout)
        dx = torch.zeros((ctx.rows, dout.shape[1]), dtype=torch.float32, device=dout.device)
        check(lib.kgcn_segment_sum_bwd_f32(ptr(dout), ptr(start), ptr(size), start.numel(), dout.shape[1], ptr(dx), _stream()))
        return dx, None, None


def segment_sum(x, sizes):
    """out[m] = sum of the rows of molecule m; ``sizes`` = atoms per molecule (any integer sequence / tensor)."""
    size = torch.as_tensor(sizes, dtype=torch.int64, device=x.device).reshape(-1).contiguous()
    start = (torch.cumsum(size, 0) - size).contiguous()
    return SegmentSumFunction.apply(x, start, size)


class GraphBatchNormFunction(torch.autograd.Function):
    """GraphBatchNormalization through kgcn_graph_bn_{fwd,bwd}_f32 (mode 0: given statistics, 1: batch statistics)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, mean, var, enabled, eps, mode):
        x = _need_cuda("inputs", x)
        B, N, F = x.shape
        y = torch.empty_like(x)
        if mode == 1:
            mean = torch.empty(F, dtype=torch.float32, device=x.device)
            var = torch.empty(F, dtype=torch.float32, device=x.device)
        ws = workspace(int(lib.kgcn_graph_bn_workspace_bytes(B, F)) + 8 * F, x.device)
        check(lib.kgcn_graph_bn_fwd_f32(ptr(x), B, N, F, ptr(enabled), ptr(gamma), ptr(beta), ptr(mean), ptr(var), eps, mode,
                                        ptr(y), ptr(ws), ws.numel(), _stream()))
        ctx.save_for_backward(x, gamma, mean, var)
        ctx.enabled, ctx.eps, ctx.mode, ctx.has_beta = enabled, eps, mode, beta is not None
        ctx.mark_non_differentiable(mean, var)
        return y, mean, var

    @staticmethod
    def backward(ctx, dy, _dmean, _dvar):
        x, gamma, mean, var = ctx.saved_tensors
        B, N, F = x.shape
        dy = _need_cuda("dy", dy)
        dx = torch.empty_like(x)
        dgamma = torch.empty(F, dtype=torch.float32, device=x.device)
        dbeta = torch.empty(F, dtype=torch.float32, device=x.device)
        ws = workspace(int(lib.kgcn_graph_bn_workspace_bytes(B, F)) + 8 * F, x.device)
        check(lib.kgcn_graph_bn_bwd_f32(ptr(x), ptr(dy), B, N, F, ptr(ctx.enabled), ptr(gamma), ptr(mean), ptr(var), ctx.eps,
                                        ctx.mode, ptr(dx), ptr(dgamma), ptr(dbeta), ptr(ws), ws.numel(), _stream()))
        return dx, (dgamma if gamma is not None else None), (dbeta if ctx.has_beta else None), None, None, None, None, None
