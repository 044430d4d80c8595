import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def unflatten_adjs(rec, prefix):
    """counts/indices/values/shapes -> list[G][C] of (indices, values, shape)."""
    counts, idx, val, shapes = rec[prefix + "counts"], rec[prefix + "indices"], rec[prefix + "values"], rec[prefix + "shapes"]
    out, pos = [], 0
    for g in range(counts.shape[0]):
        row = []
        for c in range(counts.shape[1]):
            n = int(counts[g, c])
            row.append((idx[pos:pos + n].astype(np.int32), val[pos:pos + n].astype(np.float32), [int(shapes[g, c, 0]), int(shapes[g, c, 1])]))
            pos += n
        out.append(row)
    return out


@pytest.fixture(scope="session")
def golden():
    return load_golden
