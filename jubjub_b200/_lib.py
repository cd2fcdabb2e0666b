"""ctypes loader for libjubjub_b200.so (the C ABI in include/jubjub_b200.h).

The library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).  There is no
CPU fallback: if the shared object is missing, or no CUDA device is usable, importing
callers get an exception -- never a silently different code path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# JJ_LIB lets kernel experiments load an alternative build of the SAME library (never a fallback)
LIB_PATH = os.environ.get("JJ_LIB") or os.path.join(_HERE, "libjubjub_b200.so")

# flags (include/jubjub_b200.h)
JJ_MONT = 0
JJ_CANON = 1 << 0
JJ_DEVICE_PTRS = 1 << 1
JJ_ASYNC = 1 << 2
JJ_SUBTRACT = 1 << 3
JJ_SCALAR_MONT = 1 << 4
JJ_OUT_AFFINE = 1 << 5
JJ_OUT_BYTES = 1 << 6
JJ_PRE_ZIP216 = 1 << 7
JJ_CHECK_SUBGROUP = 1 << 8
JJ_TORSION_LADDER = 1 << 9
JJ_CONST_TIME = 1 << 10

ERRORS = {0: "JJ_OK", -1: "JJ_ERR_INVALID_ARG", -2: "JJ_ERR_CUDA", -3: "JJ_ERR_NCCL", -4: "JJ_ERR_OOM",
          -5: "JJ_ERR_NO_DEVICE"}

_vp, _sz, _u32, _i32, _u64 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int32, C.c_uint64

# name -> argtypes (after the leading jj_ctx*); every entry returns int32 unless listed in _SPECIAL
_BINARY = [_vp, _vp, _vp, _sz, _u32]
_UNARY = [_vp, _vp, _sz, _u32]
_WITH_OK = [_vp, _vp, _vp, _sz, _u32]
PROTOTYPES = {
    "jj_destroy": [], "jj_sync": [],
    "jj_device_info": [C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_u64)],
    "jj_set_scalar_mul_variant": [_i32],
    "jj_malloc": [_sz, C.POINTER(_vp)], "jj_free": [_vp],
    "jj_host_alloc": [_sz, C.POINTER(_vp)], "jj_host_free": [_vp],
    "jj_memcpy_h2d": [_vp, _vp, _sz], "jj_memcpy_d2h": [_vp, _vp, _sz],
    "jj_timer_start": [], "jj_timer_stop": [C.POINTER(C.c_float)], "jj_flush_l2": [],
    "jj_measure_imad_peak": [C.POINTER(C.c_double)],
    "jj_graph_begin": [], "jj_graph_end": [C.POINTER(_vp)], "jj_graph_launch": [_vp], "jj_graph_destroy": [_vp],
    "jj_fq_mul": _BINARY, "jj_fr_mul": _BINARY, "jj_fq_add": _BINARY, "jj_fr_add": _BINARY,
    "jj_fq_sub": _BINARY, "jj_fr_sub": _BINARY,
    "jj_fq_square": _UNARY, "jj_fr_square": _UNARY, "jj_fq_neg": _UNARY, "jj_fr_neg": _UNARY,
    "jj_fq_double": _UNARY, "jj_fr_double": _UNARY,
    "jj_fq_invert": _WITH_OK, "jj_fr_invert": _WITH_OK, "jj_fq_sqrt": _WITH_OK, "jj_fr_sqrt": _WITH_OK,
    "jj_fq_to_bytes": _UNARY, "jj_fr_to_bytes": _UNARY,
    "jj_fq_from_bytes": _WITH_OK, "jj_fr_from_bytes": _WITH_OK,
    "jj_fq_from_bytes_wide": _UNARY, "jj_fr_from_bytes_wide": _UNARY,
    "jj_fq_stream": [_u64, _sz, _vp, _sz, _u32], "jj_fr_stream": [_u64, _sz, _vp, _sz, _u32],
    "jj_point_double": _UNARY, "jj_point_add": _BINARY, "jj_point_add_niels": _BINARY,
    "jj_point_add_affine_niels": _BINARY, "jj_point_to_niels": _UNARY, "jj_affine_to_niels": _UNARY,
    "jj_scalar_mul": _BINARY, "jj_scalar_mul_fixed": _BINARY,
    "jj_scalar_mul_encoded": [_vp, _vp, _vp, _vp, _sz, _u32],
    "jj_batch_normalize": _UNARY, "jj_batch_normalize_extended": _UNARY, "jj_affine_to_bytes": _UNARY,
    "jj_mul_by_cofactor": _UNARY, "jj_is_prime_order": _UNARY,
    "jj_point_neg": _UNARY, "jj_point_eq": _BINARY, "jj_affine_to_extended": _UNARY,
    "jj_point_sum": [_vp, _vp, _sz, _sz, _u32], "jj_point_sum_sharded": [_vp, _vp, _sz, _u32],
    "jj_batch_from_bytes": _WITH_OK,
    "jj_is_torsion_free": _UNARY, "jj_is_identity": _UNARY, "jj_is_small_order": _UNARY,
    "jj_comm_init": [_i32, _i32, _vp], "jj_comm_destroy": [],
    "jj_scalar_mul_sharded": [_vp, _vp, _vp, _sz, _u32],
    "jj_scalar_mul_sharded_n": [_vp, _vp, _vp, _vp, _sz, _u32],
    "jj_ipc_export": [_vp, _vp], "jj_ipc_open": [_vp, C.POINTER(_vp)], "jj_ipc_close": [_vp],
    "jj_comm_set_peer_outputs": [C.POINTER(_vp), _i32],
}
# entry points without a leading ctx / with another return type
_SPECIAL = {
    "jj_init": ([C.c_int, C.POINTER(_vp)], _i32),
    "jj_version": ([], C.c_char_p),
    "jj_last_error": ([_vp], C.c_char_p),
    "jj_launch_count": ([_vp], _u64),
    "jj_comm_unique_id": ([_vp], _i32),
}
ALL_SYMBOLS = sorted(list(PROTOTYPES) + list(_SPECIAL))

_lib = None


def load():
    """Load the shared library and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). jubjub_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, args in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = [_vp] + list(args)
        fn.restype = _i32
    for name, (args, res) in _SPECIAL.items():
        fn = getattr(lib, name)
        fn.argtypes = list(args)
        fn.restype = res
    _lib = lib
    return lib
