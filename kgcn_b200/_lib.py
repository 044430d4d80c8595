"""ctypes binding of ``libkgcn_b200.so`` (the C ABI declared in ``include/kgcn_b200.h``).

There is no CPU fallback: if the shared library is missing, importing this module raises, and
every op in the package therefore fails loudly (BASELINE.json north_star: "no CPU fallback").
Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C kgcn_b200/csrc``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libkgcn_b200.so"
LIB_PATH = os.path.join(_HERE, LIB_NAME)

STATUS_NAMES = {
    0: "KGCN_OK", 1: "KGCN_ERR_BAD_SHAPE", 2: "KGCN_ERR_MISALIGNED", 3: "KGCN_ERR_INDEX_RANGE",
    4: "KGCN_ERR_CUDA", 5: "KGCN_ERR_WORKSPACE", 6: "KGCN_ERR_UNSUPPORTED", 7: "KGCN_ERR_NULL",
}
ACT_IDS = {None: 0, "none": 0, "linear": 0, "relu": 1, "sigmoid": 2, "tanh": 3}
FLAG_DEFAULT = 0
FLAG_REFERENCE_ORDER = 1
FLAG_DY_BROADCAST = 2
FLAG_INPUTS_STABLE = 4


class KgcnError(RuntimeError):
    """A C-ABI entry point returned a non-zero kgcn_status."""

    def __init__(self, status, message):
        self.status = status
        super().__init__("%s: %s" % (STATUS_NAMES.get(status, "status %d" % status), message))


class KgcnIndexError(KgcnError, IndexError):
    """KGCN_ERR_INDEX_RANGE -- the analogue of TF's InvalidArgumentError for a bad sparse index."""


def _candidates():
    # the reference discovers its plugin libraries relative to the CWD (kgcn/layers.py:23-28);
    # keep that rule as a secondary location.
    env = os.environ.get("KGCN_B200_LIB")
    if env:
        yield env
    yield LIB_PATH
    yield os.path.join(os.getcwd(), LIB_NAME)
    yield os.path.join(os.getcwd(), "kgcn_b200.so")


def _load():
    tried = []
    for path in _candidates():
        if os.path.exists(path):
            return ctypes.CDLL(path), path
        tried.append(path)
    raise ImportError(
        "libkgcn_b200.so (the sm_100a CUDA library) was not found; there is no CPU fallback. "
        "Build it with `python -c \"import __graft_entry__ as g; g.build()\"`. Looked in: %s" % ", ".join(tried))


lib, loaded_from = _load()

_i32, _i64, _sz, _vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t, ctypes.c_void_p

# name -> (restype, argtypes); kept in the order of include/kgcn_b200.h
SIGNATURES = {
    "kgcn_abi_version": (ctypes.c_int, []),
    "kgcn_last_error": (ctypes.c_char_p, []),
    "kgcn_launch_count": (ctypes.c_uint64, []),
    "kgcn_pack_coo_host": (ctypes.c_int, [_i64, _i32, _i32, _vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "kgcn_pack_coo_device": (ctypes.c_int, [_i64, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "kgcn_bspmm_f32": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp]),
    "kgcn_bspmm_dvalues_f32": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp]),
    "kgcn_graphconv_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32, _i32]),
    "kgcn_graphconv_fwd_f32": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _i32, _vp, _vp, _i32, _i32, _vp, _i32, _vp, _sz, _vp]),
    "kgcn_graphconv_bwd_f32": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _sz, _vp]),
    "kgcn_graphdense_fwd_f32": (ctypes.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "kgcn_graphdense_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32]),
    "kgcn_graphdense_bwd_f32": (ctypes.c_int, [_vp, _i64, _i32, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "kgcn_gather_fwd_f32": (ctypes.c_int, [_vp, _i64, _i32, _i32, _vp, _vp]),
    "kgcn_gather_bwd_f32": (ctypes.c_int, [_vp, _i64, _i32, _i32, _vp, _vp]),
    "kgcn_maxpool_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32]),
    "kgcn_maxpool_fwd_f32": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _i32, _vp, _vp, _sz, _vp]),
    "kgcn_maxpool_bwd_f32": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _i32, _vp, _vp, _sz, _vp, _vp]),
    "kgcn_segment_sum_fwd_f32": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "kgcn_segment_sum_bwd_f32": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "kgcn_graph_bn_workspace_bytes": (_sz, [_i64, _i32]),
    "kgcn_graph_bn_fwd_f32": (ctypes.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, ctypes.c_float, _i32, _vp, _vp, _sz, _vp]),
    "kgcn_graph_bn_bwd_f32": (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, ctypes.c_float, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "kgcn_readout_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "kgcn_readout_xent_f32": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _i32, _vp, _vp, ctypes.c_float, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "kgcn_gather_readout_xent_f32": (ctypes.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, ctypes.c_float, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "kgcn_graphconv_fwd_fused": (_i32, [_i64, _i32, _i32, _i32, _i32]),
    "kgcn_graphconv_fwd_padded_f32": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "kgcn_gather_readout_xent_du_f32": (ctypes.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, ctypes.c_float, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _sz, _vp]),
    "kgcn_graphconv_bwd_splits": (_i32, [_i64, _i32, _i32, _i32, _i32, _i32]),
    "kgcn_graphconv_bwd_partial_f32": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _i32, _vp, _sz, _vp]),
    "kgcn_graphconv_chain_supported": (_i32, [_i64, _i32, _i32, _i32, _vp]),
    "kgcn_graphconv_chain_fwd_f32": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "kgcn_graphconv_chain_dx_f32": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp]),
    "kgcn_gcn_step_chain_grid": (_i32, [_i64, _i32, _i32, _i32, _vp, _i32]),
    "kgcn_gcn_step_chain_f32": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32,
                                               _vp, _vp, _i32, _vp, _vp, ctypes.c_float, _vp, _vp, _vp, _vp, ctypes.c_uint32, _vp]),
    "kgcn_graphconv_chain_dw_f32": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "kgcn_graphconv_chain_dw_g_f32": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "kgcn_gcn_step_chain_g_supported": (_i32, [_i64, _i32, _i32, _i32, _vp]),
    "kgcn_gcn_step_chain_g_f32": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32,
                                                 _vp, _vp, _i32, _vp, _vp, ctypes.c_float, _vp, _vp, _vp, _vp, ctypes.c_uint32, _vp]),
    "kgcn_reduce_partials_f32": (ctypes.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "kgcn_reduce_adam_f32": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _i32, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp, _i32, _i64, _vp, _vp]),
    "kgcn_p2p_alloc": (ctypes.c_int, [_sz, _vp, _vp]),
    "kgcn_p2p_open": (ctypes.c_int, [_vp, _vp]),
    "kgcn_p2p_close": (ctypes.c_int, [_vp]),
    "kgcn_p2p_free": (ctypes.c_int, [_vp]),
    "kgcn_adam_f32": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, _i64, ctypes.c_float, _vp, _vp]),
    "kgcn_crc32c": (ctypes.c_uint32, [_vp, _sz]),
    "kgcn_crc32c_masked": (ctypes.c_uint32, [_vp, _sz]),
    "kgcn_tfrecord_scan": (ctypes.c_int, [_vp, _sz, _i32, _vp, _vp, _i64, _vp]),
    "kgcn_tfexample_gather": (ctypes.c_int, [_vp, _vp, _vp, _i64, ctypes.c_char_p, _i32, _vp, _i64, _vp, _vp]),
}



class GradSegment(ctypes.Structure):
    """kgcn_grad_segment (include/kgcn_b200.h)."""
    _fields_ = [("kernel_off", _i64), ("bias_off", _i64), ("partial", _vp), ("splits", _i32), ("rows", _i32), ("cols", _i32),
                ("channels", _i32), ("stride", _i64)]


class P2PGroup(ctypes.Structure):
    """kgcn_p2p_group (include/kgcn_b200.h)."""
    _fields_ = [("rank", _i32), ("world", _i32), ("n_pad", _i64), ("mailbox", _vp * 8), ("error_flag", _vp)]


for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header and library out of sync: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args


def last_error():
    return lib.kgcn_last_error().decode("utf-8", "replace")


def check(status):
    if status != 0:
        cls = KgcnIndexError if status == 3 else KgcnError
        raise cls(status, last_error())


def ptr(t):
    """Device / host address of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data
