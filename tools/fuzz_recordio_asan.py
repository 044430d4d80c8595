#!/usr/bin/env python
"""AddressSanitizer + UBSan fuzz of the host record-IO entry points (csrc/recordio.cu): truncated, mutated and random
tensorflow.Example records and TFRecord files, each in an exact-size heap buffer so that any read past the record is
caught.  CPU only.  usage: python tools/fuzz_recordio_asan.py [trials]   (builds /tmp/librec_asan.so with g++, then
re-executes itself with libasan preloaded)."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = "/tmp/librec_asan.so"

if os.environ.get("KGCN_ASAN_CHILD") != "1":
    src = [os.path.join(ROOT, "kgcn_b200", "csrc", f) for f in ("recordio.cu", "abi.cu")]
    subprocess.check_call(["g++", "-x", "c++", "-std=c++17", "-g", "-O1", "-include", "cmath", "-fsanitize=address,undefined",
                           "-fno-omit-frame-pointer", "-shared", "-fPIC"] + src + ["-I/usr/local/cuda/include", "-o", SO])
    pre = ":".join(subprocess.check_output(["gcc", "-print-file-name=" + n], text=True).strip() for n in ("libasan.so", "libubsan.so"))
    env = dict(os.environ, LD_PRELOAD=pre, ASAN_OPTIONS="detect_leaks=0", KGCN_ASAN_CHILD="1")
    sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env))

import numpy as np  # noqa: E402

lib = ctypes.CDLL(SO)
vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
lib.kgcn_tfrecord_scan.argtypes = [vp, ctypes.c_size_t, i32, vp, vp, i64, vp]
lib.kgcn_tfexample_gather.argtypes = [vp, vp, vp, i64, ctypes.c_char_p, i32, vp, i64, vp, vp]
lib.kgcn_crc32c.argtypes = [vp, ctypes.c_size_t]
rng = np.random.default_rng(0)


def varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def lf(f, p):
    return varint((f << 3) | 2) + varint(len(p)) + p


def example():
    ints = lf(3, lf(1, b"".join(varint(int(x)) for x in rng.integers(-5, 1000, 20))))
    flts = lf(2, lf(1, rng.standard_normal(10).astype("<f4").tobytes()))
    return lf(1, lf(1, lf(1, b"adj_row") + lf(2, ints)) + lf(1, lf(1, b"adj_values") + lf(2, flts)))


trials = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
n_ok = n_err = 0
for trial in range(trials):
    good = example()
    mode = trial % 4
    if mode == 0:
        rec = good[:int(rng.integers(0, len(good) + 1))]
    elif mode == 1:
        b = bytearray(good)
        for _ in range(int(rng.integers(1, 5))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        rec = bytes(b)
    elif mode == 2:
        rec = rng.integers(0, 256, int(rng.integers(0, 120)), dtype=np.uint8).tobytes()
    else:
        rec = good
    buf = ctypes.create_string_buffer(rec, max(len(rec), 1)) if rec else ctypes.create_string_buffer(1)
    heap = (ctypes.c_ubyte * max(len(rec), 1)).from_buffer_copy(buf.raw[:max(len(rec), 1)])
    off, ln = np.array([0], np.int64), np.array([len(rec)], np.int64)
    for key, kind, dt in ((b"adj_row", 2, np.int64), (b"adj_values", 1, np.float32), (b"x", 2, np.int64)):
        cap = int(rng.integers(0, 40))
        vals, counts, total = np.empty(max(cap, 1), dt), np.zeros(1, np.int64), i64(0)
        rc = lib.kgcn_tfexample_gather(ctypes.addressof(heap), off.ctypes.data, ln.ctypes.data, 1, key, kind, vals.ctypes.data, cap,
                                       counts.ctypes.data, ctypes.byref(total))
        n_ok += rc == 0
        n_err += rc != 0
    nrec = i64(0)
    lib.kgcn_tfrecord_scan(ctypes.addressof(heap), len(rec), trial & 1, None, None, 0, ctypes.byref(nrec))
    lib.kgcn_crc32c(ctypes.addressof(heap), len(rec))
print("no sanitizer report over %d trials: %d calls parsed, %d rejected" % (trials, n_ok, n_err))
