"""``Bspmm`` + ``Bspmdt`` of ``batched.so`` (mirror of kgcn/batched_call.py).

    BatchedSpMM().call(sp_matrices, dense_matrices, ...) -> list[N]            batched_call.py:6-14
    BatchedSpMDT().call(sp_matrices, dense, ...)         -> list[N]            batched_call.py:17-26

``Bspmdt`` multiplies N sparse matrices with ONE stacked dense matrix ``[N*rows, cols]`` (the
``tf.reshape(inputs, [batch_size*input_row, input_col])`` of kgcn/layers.py:99); that stacked
form is exactly the contiguous ``[B, N, F]`` layout the kernel reads, so no copy is made.
"""
from . import _plugin
from .bspmm_call import BatchedSpMM  # noqa: F401  (same op, same contract)


class BatchedSpMDT:
    def __init__(self):
        from . import _lib  # noqa: F401

    def call(self, sp_matrices, dense_matrices, adjoint_a=False, adjoint_b=False):
        device = _plugin.default_device(dense_matrices)
        csr, flat, _ = _plugin.pack_sparse_list(sp_matrices, device, nested=False)
        if adjoint_a:
            csr = csr.transposed()
        dense = _plugin.to_device_f32(dense_matrices, device)
        n = csr.n_graphs
        if dense.shape[0] % n != 0:
            raise ValueError("stacked dense has %d rows, not a multiple of %d matrices" % (dense.shape[0], n))
        rhs = dense.reshape(n, 1, dense.shape[0] // n, dense.shape[1])
        if adjoint_b:
            rhs = rhs.transpose(2, 3)
        out = _plugin.run(csr, flat, rhs.contiguous(), "per_matrix")
        return list(out[:, 0].unbind(0))

    def call_packed(self, csr, rhs, flat_values=None):
        """rhs [B, C, K, F] -> [B, R, F] (channel sum = the tf.reduce_sum of layers.py:103)."""
        return _plugin.run(csr, flat_values, rhs.contiguous(), "sum")
