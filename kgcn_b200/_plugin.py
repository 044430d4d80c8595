"""Shared plumbing of the three plugin-op wrappers (bspmm_call / bconv_call / batched_call)."""
import numpy as np
import torch

from . import ops
from .csr import BatchedCSR, flatten_coo, _triple


def pack_sparse_list(sp_matrices, device, nested):
    """list (or list-of-lists) of SparseTensorValue-like triples -> (BatchedCSR, flat values tensor
    or None).  If any ``values`` entry is a tensor that requires grad, the CSR value arrays are
    gathered from the concatenated values ON DEVICE so autograd can return per-matrix gradients
    (kgcn/bspmm_call.py:49-54)."""
    rows = sp_matrices if nested else [[m] for m in sp_matrices]
    want_grad = any(torch.is_tensor(_triple(m)[1]) and _triple(m)[1].requires_grad for row in rows for m in row)
    counts, indices, values, shape = flatten_coo(rows)
    csr = BatchedCSR.from_flat(counts, indices, values, shape[0], shape[1], device=device, want_perm=want_grad)
    flat = None
    if want_grad:
        flat = torch.cat([_triple(m)[1].reshape(-1).to(device=device, dtype=torch.float32) for row in rows for m in row])
        csr = csr.with_values(flat[csr.perm.long()], flat[csr.perm_t.long()])
    return csr, flat, counts


def to_device_f32(t, device):
    if torch.is_tensor(t):
        return t.to(device=device, dtype=torch.float32)
    return torch.as_tensor(np.asarray(t, np.float32), device=device)


def default_device(dense):
    first = dense
    while isinstance(first, (list, tuple)):
        first = first[0]
    if torch.is_tensor(first) and first.is_cuda:
        return first.device
    return torch.device("cuda", torch.cuda.current_device())


def run(csr, flat_values, rhs, layout):
    return ops.BspmmFunction.apply(rhs, flat_values, csr, layout)
