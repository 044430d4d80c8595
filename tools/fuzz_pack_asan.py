#!/usr/bin/env python
"""AddressSanitizer + UBSan fuzz of the HOST packer ``kgcn_pack_coo_host`` (csrc/pack.cu, multi-threaded): random COO batches
(empty matrices, duplicates, unsorted entries, int32 / int64 indices, both orientations, out-of-range indices) in exact-size
buffers, every result compared with an independent numpy stable sort.  CPU only (the CUDA kernels of pack.cu are compiled but
never launched).  usage: python tools/fuzz_pack_asan.py [trials]"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = "/tmp/libpack_asan.so"

if os.environ.get("KGCN_ASAN_CHILD") != "1":
    src = [os.path.join(ROOT, "kgcn_b200", "csrc", f) for f in ("pack.cu", "abi.cu")]
    subprocess.check_call(["nvcc", "-O1", "-g", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler",
                           "-fPIC,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer", "--expt-relaxed-constexpr",
                           "-shared"] + src + ["-o", SO, "-lcudart"])
    pre = ":".join(subprocess.check_output(["gcc", "-print-file-name=" + n], text=True).strip() for n in ("libasan.so", "libubsan.so"))
    env = dict(os.environ, LD_PRELOAD=pre, ASAN_OPTIONS="detect_leaks=0:protect_shadow_gap=0", KGCN_ASAN_CHILD="1")
    sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env))

import numpy as np  # noqa: E402

lib = ctypes.CDLL(SO)
vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
lib.kgcn_pack_coo_host.argtypes = [i64, i32, i32, vp, vp, i32, vp, i32, vp, vp, vp, vp]
rng = np.random.default_rng(0)
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
n_ok = n_range = 0
for trial in range(trials):
    big = trial % 250 == 249          # nnz > 65536 and >= 64 matrices: the multi-threaded path of the packer
    n_mat = int(rng.integers(1500, 3000)) if big else int(rng.integers(0, 40))
    R, K = int(rng.integers(1, 20)), int(rng.integers(1, 20))
    counts = rng.integers(20, 60, size=n_mat) if big else rng.integers(0, 30, size=n_mat)
    if trial % 5 == 0:
        counts[rng.random(n_mat) < 0.5] = 0
    off = np.zeros(n_mat + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    nnz = int(off[-1])
    rows, cols = rng.integers(0, R, nnz), rng.integers(0, K, nnz)
    bad = trial % 7 == 3 and nnz > 0
    if bad:
        j = int(rng.integers(0, nnz))
        if rng.random() < 0.5:
            rows[j] = R + int(rng.integers(0, 3)) if rng.random() < 0.5 else -1 - int(rng.integers(0, 3))
        else:
            cols[j] = K + int(rng.integers(0, 3)) if rng.random() < 0.5 else -1 - int(rng.integers(0, 3))
    is64 = trial % 2
    idx = np.ascontiguousarray(np.stack([rows, cols], 1).astype(np.int64 if is64 else np.int32).reshape(-1, 2))
    val = rng.standard_normal(nnz).astype(np.float32)
    for transpose in (0, 1):
        out_rows = K if transpose else R
        rowptr = np.full(n_mat * out_rows + 1, -7, np.int32)
        col, v, perm = np.full(max(nnz, 1), -7, np.int32)[:nnz], np.zeros(max(nnz, 1), np.float32)[:nnz], np.full(max(nnz, 1), -7, np.int32)[:nnz]
        want_perm = trial % 3 != 0
        rc = lib.kgcn_pack_coo_host(n_mat, R, K, off.ctypes.data, idx.ctypes.data if nnz else None, is64, val.ctypes.data if nnz else None,
                                    transpose, rowptr.ctypes.data, col.ctypes.data if nnz else None, v.ctypes.data if nnz else None,
                                    perm.ctypes.data if (want_perm and nnz) else None)
        if bad:
            assert rc == 3, (trial, rc)
            n_range += 1
            continue
        assert rc == 0, (trial, rc)
        key_r, key_c = (cols, rows) if transpose else (rows, cols)
        mat = np.repeat(np.arange(n_mat), counts)
        order = np.argsort(mat * out_rows + key_r, kind="stable")
        assert np.array_equal(col, key_c[order]) and np.array_equal(v, val[order])
        if want_perm and nnz:
            assert np.array_equal(perm, order)
        want_ptr = np.zeros(n_mat * out_rows + 1, np.int64)
        np.add.at(want_ptr, mat * out_rows + key_r + 1, 1)
        assert np.array_equal(rowptr, np.cumsum(want_ptr))
        n_ok += 1
print("no sanitizer report over %d trials: %d packs equal to the numpy stable sort, %d out-of-range batches rejected" % (trials, n_ok, n_range))
