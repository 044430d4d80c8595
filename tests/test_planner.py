"""Host-side planners of the library (pure host code behind the C ABI: callable without a GPU): every BASELINE config
(C2 ring graphs, C3 Tox21 scale, C4 three bond types, C5's per-GPU shard) must be taken by the fused
tcgen05 kernels at its full batch -- forward layer, chained step launch, weight-gradient kernel, stored-G variant -- on the
STORED widths (75 -> 96, 50 -> 64: Trainer.pad_features); the unpadded widths of C3 / C4 are not eligible, which is what the
padding is for."""
import ctypes

import pytest

from kgcn_b200._lib import lib

CONFIGS = {
    "c2": (1024, 1, 32, [64, 64, 64]),
    "c3": (512, 1, 50, [96, 64, 64, 64]),
    "c4": (512, 3, 50, [96, 64, 64, 64]),
    "c5": (512, 1, 64, [128, 128, 128]),
}


def _dims(d):
    return (ctypes.c_int32 * len(d))(*d)


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_baseline_configs_run_on_the_fused_kernels(name):
    B, C, N, d = CONFIGS[name]
    L = len(d) - 1
    assert all(lib.kgcn_graphconv_fwd_fused(B, C, N, d[i], d[i + 1]) for i in range(L))
    assert lib.kgcn_graphconv_chain_supported(B, C, N, L, _dims(d))
    grid = lib.kgcn_gcn_step_chain_grid(B, C, N, L, _dims(d), 2)
    assert 0 < grid <= 148                                           # one CTA per SM, contiguous graph ranges
    assert grid * -(-B // grid) >= B
    splits = [lib.kgcn_graphconv_bwd_splits(B, C, N, d[i], d[i + 1], 1 if i > 0 else 0) for i in range(L)]
    assert all(s == grid for s in splits)                            # the weight-gradient CTAs own the same graph ranges
    assert lib.kgcn_gcn_step_chain_g_supported(B, C, N, L, _dims(d))


def test_unpadded_widths_and_odd_shapes_are_not_eligible():
    B, C, N = 512, 1, 50
    d = [75, 50, 50, 50]
    assert not any(lib.kgcn_graphconv_fwd_fused(B, C, N, d[i], d[i + 1]) for i in range(3))
    assert lib.kgcn_gcn_step_chain_grid(B, C, N, 3, _dims(d), 2) == 0
    assert lib.kgcn_gcn_step_chain_g_supported(B, C, N, 3, _dims(d)) == 0
    assert lib.kgcn_gcn_step_chain_g_supported(1024, 1, 32, 1, _dims([64, 64])) == 0      # one layer: there is no dx job
    assert lib.kgcn_gcn_step_chain_grid(1024, 1, 32, 2, _dims([64, 64, 64]), 5) == 0      # head: at most 4 labels
    assert lib.kgcn_gcn_step_chain_grid(0, 1, 32, 2, _dims([64, 64, 64]), 2) == 0
