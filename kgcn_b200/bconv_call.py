"""``Bconv`` -- batched sparse x dense with the channel sum fused (mirror of kgcn/bconv_call.py).

    BatchedConv().call(sp_matrices, dense_matrices, adjoint_a=False, adjoint_b=False) -> list[B]

``sp_matrices`` and ``dense_matrices`` are ``list[B][C]``; output ``b`` is
``sum_c A[b][c] . D[b][c]`` (bconv_call.py:10-21, flatten order batch-major / channel-minor).
Gradient as registered by the reference (bconv_call.py:28-70) through autograd.
"""
import torch

from . import _plugin


class BatchedConv:
    def __init__(self):
        from . import _lib  # noqa: F401

    def call(self, sp_matrices, dense_matrices, adjoint_a=False, adjoint_b=False):
        device = _plugin.default_device(dense_matrices)
        csr, flat, _ = _plugin.pack_sparse_list(sp_matrices, device, nested=True)
        if adjoint_a:
            csr = csr.transposed()
        rows = []
        for dms in dense_matrices:
            ds = [_plugin.to_device_f32(d, device) for d in dms]
            if adjoint_b:
                ds = [d.t() for d in ds]
            rows.append(torch.stack([d.contiguous() for d in ds]))
        rhs = torch.stack(rows)                                                   # [B, C, K, F]
        out = _plugin.run(csr, flat, rhs, "sum")                                  # [B, R, F]
        return list(out.unbind(0))

    def call_packed(self, csr, rhs, flat_values=None):
        """rhs [B, C, K, F] -> [B, R, F]."""
        return _plugin.run(csr, flat_values, rhs.contiguous(), "sum")
