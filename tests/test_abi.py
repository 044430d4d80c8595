"""CPU: the C-ABI library loads without a GPU and exports every symbol include/kgcn_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "kgcn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kgcn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from kgcn_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 14
    raw = ctypes.CDLL(_lib.loaded_from)
    for n in names:
        assert hasattr(raw, n), "header declares %s but the library does not export it" % n
        assert n in _lib.SIGNATURES, "%s has no ctypes signature in kgcn_b200/_lib.py" % n
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_error_string():
    from kgcn_b200 import _lib
    assert _lib.lib.kgcn_abi_version() == 1
    # argument validation happens before any CUDA call, so it works without a device
    rc = _lib.lib.kgcn_gather_fwd_f32(None, 1, 1, 1, None, None)
    assert rc == 7 and "NULL" in _lib.last_error()
    rc = _lib.lib.kgcn_bspmm_f32(1, 1, 1, 1, 0, 1, 1, 1, 1, 0, 0, 1, 0, 0, None, None)
    assert rc == 1 and "bad shape" in _lib.last_error()


def test_no_cpu_fallback():
    import pytest
    import torch
    from kgcn_b200 import KgcnError, layers
    with pytest.raises(KgcnError):
        layers.GraphGather()(torch.zeros(2, 3, 4))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "kgcn_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_hostmem_numa_helpers_degrade_gracefully():
    """kgcn_b200.hostmem (NUMA placement of the pinned staging buffers): unknown nodes and disallowed nodes are no-ops,
    an allowed node sets and restores the thread's memory policy, and the page query answers for a touched buffer."""
    import numpy as np
    import torch
    from kgcn_b200 import hostmem
    with hostmem.numa_preferred(-1) as ok:
        assert ok is False
    with hostmem.numa_preferred(None) as ok:
        assert ok is False
    with hostmem.numa_preferred(4095) as ok:          # no such node / not in Mems_allowed: preference silently skipped
        assert ok is False
    allowed = hostmem.mems_allowed()
    if allowed is not None:
        first = int(allowed.split(",")[0].split("-")[0])
        with hostmem.numa_preferred(first) as ok:
            t = torch.from_numpy(np.ones(1 << 16, np.uint8))
            assert ok in (True, False)
        assert hostmem.node_of_buffer(t) in (-1, first) or not ok
    assert hostmem.gpu_numa_node(0) >= -1
