"""Config 5's data, generated ON THE DEVICE per rank (1 048 576 molecules of 64 atoms and 128
features are 32 GB -- host RAM cannot hold them comfortably, one B200 can): torch's device RNG (plumbing, outside any
timed region) draws the graphs as flat COO, the library's device packer (kgcn_pack_coo_device) turns them into the
BatchedCSR + its transpose.

Molecule model (the N = 64 case of synth.random_molecule_coo, every molecule fully populated): a random spanning tree
(atom i >= 1 bonds to a uniformly random earlier atom), ``extra`` ring-closure bonds between random atom pairs, symmetric,
diagonal 1 -- ``N + 2 (N - 1) + 2 extra`` = 202 entries per molecule at N = 64, extra = 6 (duplicate pairs are kept and
accumulate, exactly like duplicate COO entries fed to the reference).  Features N(0, 1), labels = parity of the number of
ring closures that landed on distinct atoms (learnable from the structure, not used by the throughput numbers)."""
import torch

from ._lib import check, lib, ptr
from .csr import BatchedCSR


def molecule_coo(gen, B, N, extra=6, device="cuda"):
    """-> (indices [B, E, 2] int32, values [B, E] f32) with E = N + 2 (N - 1) + 2 extra, unsorted inside a molecule."""
    dev = torch.device(device)
    node = torch.arange(1, N, device=dev)
    parent = (torch.rand(B, N - 1, generator=gen, device=dev) * node).long()
    parent = torch.minimum(parent, node - 1)
    child = node.expand(B, N - 1)
    u = torch.randint(0, N, (B, extra), generator=gen, device=dev)
    v = torch.randint(0, N, (B, extra), generator=gen, device=dev)
    diag = torch.arange(N, device=dev).expand(B, N)
    rows = torch.cat([diag, child, parent, u, v], 1)
    cols = torch.cat([diag, parent, child, v, u], 1)
    idx = torch.stack([rows, cols], 2).to(torch.int32).contiguous()
    vals = torch.ones(idx.shape[:2], dtype=torch.float32, device=dev)
    labels_int = ((u != v).sum(1) % 2).long()
    return idx, vals, labels_int


def pack_device(idx, vals, B, C, N, stream=None):
    """Flat device COO (indices [B*C, E, 2] int32) -> BatchedCSR (+ transpose) through kgcn_pack_coo_device."""
    dev = idx.device
    E = idx.shape[1]
    nnz = B * C * E
    off = torch.arange(0, B * C + 1, device=dev, dtype=torch.int64) * E
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream if stream is None else stream
    out = []
    for tr in (0, 1):
        rp = torch.empty(B * C * N + 1, dtype=torch.int32, device=dev)
        col = torch.empty(nnz, dtype=torch.int32, device=dev)
        val = torch.empty(nnz, dtype=torch.float32, device=dev)
        check(lib.kgcn_pack_coo_device(B * C, N, N, ptr(off), ptr(idx), ptr(vals), tr, ptr(rp), ptr(col), ptr(val), None, ptr(flag), st))
        out += [rp, col, val]
    if int(flag.item()) != 0:
        raise RuntimeError("device packer flagged an out-of-range index")
    return BatchedCSR(B, C, N, N, *out)


def device_batches(seed, n_batches, B, N, F, extra=6, device="cuda"):
    """``n_batches`` resident batches of ``B`` molecules: list of dicts (csr, features [B, N, F], labels [B, 2], mask [B], and
    the flat COO ``idx`` [B, E, 2] / ``vals`` [B, E] they were packed from)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    out = []
    for _ in range(n_batches):
        idx, vals, labels_int = molecule_coo(gen, B, N, extra, device)
        csr = pack_device(idx, vals, B, 1, N)
        feats = torch.randn(B, N, F, generator=gen, device=device, dtype=torch.float32)
        labels = torch.nn.functional.one_hot(labels_int, 2).to(torch.float32)
        out.append({"csr": csr, "features": feats, "labels": labels, "mask": torch.ones(B, device=device), "idx": idx, "vals": vals})
    return out
