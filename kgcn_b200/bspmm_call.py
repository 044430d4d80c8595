"""``Bspmm`` -- N independent sparse x dense products (mirror of kgcn/bspmm_call.py).

    BatchedSpMM().call(sp_matrices, dense_matrices, adjoint_a=False, adjoint_b=False) -> list[N]

``sp_matrices``: list[N] of ``(indices[nnz,2], values[nnz], dense_shape[2])`` (SparseTensorValue-like);
``dense_matrices``: list[N] of ``[rows, cols]``.  The registered gradient of the reference
(bspmm_call.py:21-57: ``d rhs = Bspmm(A, dY, adjoint_a=True)``, ``d values = gather-dot``) is
provided through autograd.  The op itself is one launch of ``kgcn_bspmm_f32``.
"""
import torch

from . import _plugin


class BatchedSpMM:
    def __init__(self):
        from . import _lib  # noqa: F401  (loads libkgcn_b200.so or raises -- there is no fallback)

    def call(self, sp_matrices, dense_matrices, adjoint_a=False, adjoint_b=False):
        device = _plugin.default_device(dense_matrices)
        csr, flat, _ = _plugin.pack_sparse_list(sp_matrices, device, nested=False)
        if adjoint_a:
            csr = csr.transposed()
        dense = [_plugin.to_device_f32(d, device) for d in dense_matrices]
        if adjoint_b:
            dense = [d.t() for d in dense]
        rhs = torch.stack([d.contiguous() for d in dense]).unsqueeze(1)          # [N, 1, K, F]
        out = _plugin.run(csr, flat, rhs, "per_matrix")                           # [N, 1, R, F]
        return list(out[:, 0].unbind(0))

    def call_packed(self, csr, rhs, flat_values=None):
        """rhs [B, C, K, F] -> [B, C, R, F]: the B*C independent products in one launch."""
        return _plugin.run(csr, flat_values, rhs.contiguous(), "per_matrix")
