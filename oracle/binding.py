"""ctypes binding of oracle/libjj_oracle.so for the test-suite and the CPU baseline.

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by jubjub_b200/.

Arrays are numpy uint64 with a trailing limb axis: field elements (n, 4),
extended points (n, 20), affine (n, 8), extended-Niels (n, 16), affine-Niels
(n, 12); byte strings are uint8 (n, 32).  Limbs are Montgomery form unless noted.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libjj_oracle.so")

FQ, FR = 0, 1
OP_MUL, OP_SQUARE, OP_ADD, OP_SUB, OP_NEG, OP_DOUBLE = range(6)


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "jj_oracle.c"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libjj_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.jo_time_fe_mul.restype = C.c_double
        _lib.jo_time_scalar_mul.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u64(a, width):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.ndim == 2 and a.shape[1] == width, (a.shape, width)
    return a


def _u8(a, width=32):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2 and a.shape[1] == width, (a.shape, width)
    return a


# ---- field -------------------------------------------------------------------
def fe_batch(which, op, a, b=None):
    a = _u64(a, 4)
    b = a if b is None else _u64(b, 4)
    out = np.empty_like(a)
    lib().jo_fe_batch(which, op, _p(a), _p(b), _p(out), C.c_size_t(len(a)))
    return out


def fe_invert(which, a):
    a = _u64(a, 4)
    out = np.empty_like(a)
    ok = np.empty(len(a), dtype=np.uint8)
    lib().jo_fe_batch_invert(which, _p(a), _p(out), _p(ok), C.c_size_t(len(a)))
    return out, ok


def fe_to_bytes(which, a):
    a = _u64(a, 4)
    out = np.empty((len(a), 32), dtype=np.uint8)
    lib().jo_fe_batch_to_bytes(which, _p(a), _p(out), C.c_size_t(len(a)))
    return out


def fe_from_bytes(which, b):
    b = _u8(b)
    out = np.empty((len(b), 4), dtype=np.uint64)
    ok = np.empty(len(b), dtype=np.uint8)
    lib().jo_fe_batch_from_bytes(which, _p(b), _p(out), _p(ok), C.c_size_t(len(b)))
    return out, ok


def fe_from_bytes_wide(which, b64):
    b64 = _u8(b64, 64)
    out = np.empty((len(b64), 4), dtype=np.uint64)
    for i in range(len(b64)):
        lib().jo_fe_from_bytes_wide(which, _p(b64[i]), _p(out[i]))
    return out


def fe_from_raw(which, raw):
    raw = _u64(raw, 4)
    out = np.empty_like(raw)
    for i in range(len(raw)):
        lib().jo_fe_from_raw(which, _p(raw[i]), _p(out[i]))
    return out


def fe_sqrt(which, a):
    a = _u64(a, 4)
    out = np.zeros_like(a)
    ok = np.empty(len(a), dtype=np.uint8)
    for i in range(len(a)):
        ok[i] = lib().jo_fe_sqrt(which, _p(a[i]), _p(out[i]))
    return out, ok


def fe_pow_vartime(which, a, e):
    a = _u64(a, 4)
    e = np.ascontiguousarray(e, dtype=np.uint64)
    out = np.empty_like(a)
    for i in range(len(a)):
        lib().jo_fe_pow_vartime(which, _p(a[i]), _p(e), _p(out[i]))
    return out


def fe_one(which):
    out = np.empty((1, 4), dtype=np.uint64)
    lib().jo_fe_one(which, _p(out))
    return out


def fe_stream(which, seed, n, first=0):
    out = np.empty((n, 4), dtype=np.uint64)
    lib().jo_fe_stream(which, C.c_uint64(seed), C.c_size_t(first), C.c_size_t(n), _p(out))
    return out


# ---- points ------------------------------------------------------------------
def generator():
    out = np.empty((1, 8), dtype=np.uint64)
    lib().jo_generator(_p(out))
    return out


def identity(n=1):
    out = np.empty((n, 20), dtype=np.uint64)
    for i in range(n):
        lib().jo_ext_identity(_p(out[i]))
    return out


def affine_to_extended(a):
    a = _u64(a, 8)
    one = fe_one(FQ)[0]
    out = np.empty((len(a), 20), dtype=np.uint64)
    out[:, 0:4] = a[:, 0:4]
    out[:, 4:8] = a[:, 4:8]
    out[:, 8:12] = one
    out[:, 12:16] = a[:, 0:4]
    out[:, 16:20] = a[:, 4:8]
    return out


def _each(fn, width_out, *arrays):
    n = len(arrays[0])
    out = np.empty((n, width_out), dtype=np.uint64)
    for i in range(n):
        fn(*[_p(a[i]) for a in arrays], _p(out[i]))
    return out


def ext_to_affine(p):
    return _each(lib().jo_ext_to_affine, 8, _u64(p, 20))


def ext_to_niels(p):
    return _each(lib().jo_ext_to_niels, 16, _u64(p, 20))


def affine_to_niels(p):
    return _each(lib().jo_affine_to_niels, 12, _u64(p, 8))


def ext_neg(p):
    return _each(lib().jo_ext_neg, 20, _u64(p, 20))


def ext_double(p):
    p = _u64(p, 20)
    out = np.empty_like(p)
    lib().jo_batch_double(_p(p), _p(out), C.c_size_t(len(p)))
    return out


def ext_add(p, q):
    p, q = _u64(p, 20), _u64(q, 20)
    out = np.empty_like(p)
    lib().jo_batch_add(_p(p), _p(q), _p(out), C.c_size_t(len(p)))
    return out


def ext_sub(p, q):
    return _each(lib().jo_ext_sub, 20, _u64(p, 20), _u64(q, 20))


def ext_add_niels(p, n):
    p, n = _u64(p, 20), _u64(n, 16)
    out = np.empty_like(p)
    lib().jo_batch_add_niels(_p(p), _p(n), _p(out), C.c_size_t(len(p)))
    return out


def ext_sub_niels(p, n):
    return _each(lib().jo_ext_sub_niels, 20, _u64(p, 20), _u64(n, 16))


def ext_add_affine_niels(p, n):
    p, n = _u64(p, 20), _u64(n, 12)
    out = np.empty_like(p)
    lib().jo_batch_add_affine_niels(_p(p), _p(n), _p(out), C.c_size_t(len(p)))
    return out


def ext_sub_affine_niels(p, n):
    return _each(lib().jo_ext_sub_affine_niels, 20, _u64(p, 20), _u64(n, 12))


def ext_mul_by_cofactor(p):
    return _each(lib().jo_ext_mul_by_cofactor, 20, _u64(p, 20))


def scalar_mul(points, scalars32, nthreads=0):
    points, scalars32 = _u64(points, 20), _u8(scalars32)
    assert len(points) == len(scalars32)
    out = np.empty_like(points)
    lib().jo_batch_scalar_mul(_p(points), _p(scalars32), _p(out), C.c_size_t(len(points)),
                              nthreads or os.cpu_count())
    return out


def scalar_mul_fixed(base_affine, scalars32, nthreads=0):
    base_affine, scalars32 = _u64(base_affine, 8), _u8(scalars32)
    out = np.empty((len(scalars32), 20), dtype=np.uint64)
    lib().jo_batch_scalar_mul_fixed(_p(base_affine), _p(scalars32), _p(out),
                                    C.c_size_t(len(scalars32)), nthreads or os.cpu_count())
    return out


def affine_niels_mul(n, scalars32):
    n, scalars32 = _u64(n, 12), _u8(scalars32)
    return _each(lib().jo_affine_niels_mul_bits, 20, n, scalars32)


def batch_normalize(p):
    p = _u64(p, 20)
    out = np.empty((len(p), 8), dtype=np.uint64)
    lib().jo_batch_normalize(_p(p), _p(out), C.c_size_t(len(p)))
    return out


def affine_to_bytes(a):
    a = _u64(a, 8)
    out = np.empty((len(a), 32), dtype=np.uint8)
    lib().jo_batch_to_bytes(_p(a), _p(out), C.c_size_t(len(a)))
    return out


def affine_from_bytes(b, zip216=True):
    b = _u8(b)
    out = np.zeros((len(b), 8), dtype=np.uint64)
    ok = np.empty(len(b), dtype=np.uint8)
    for i in range(len(b)):
        ok[i] = lib().jo_affine_from_bytes(_p(b[i]), int(zip216), _p(out[i]))
    return out, ok


def batch_from_bytes(b):
    b = _u8(b)
    out = np.empty((len(b), 8), dtype=np.uint64)
    ok = np.empty(len(b), dtype=np.uint8)
    lib().jo_batch_from_bytes(_p(b), _p(out), _p(ok), C.c_size_t(len(b)))
    return out, ok


def _flags(fn, p):
    p = _u64(p, 20)
    return np.array([fn(_p(p[i])) for i in range(len(p))], dtype=np.uint8)


def is_identity(p):
    return _flags(lib().jo_ext_is_identity, p)


def is_small_order(p):
    return _flags(lib().jo_ext_is_small_order, p)


def is_torsion_free(p, nthreads=0):
    p = _u64(p, 20)
    out = np.empty(len(p), dtype=np.uint8)
    lib().jo_batch_is_torsion_free(_p(p), _p(out), C.c_size_t(len(p)), nthreads or os.cpu_count())
    return out


def ext_eq(p, q):
    p, q = _u64(p, 20), _u64(q, 20)
    return np.array([lib().jo_ext_eq(_p(p[i]), _p(q[i])) for i in range(len(p))], dtype=np.uint8)


def is_on_curve(a):
    a = _u64(a, 8)
    return np.array([lib().jo_affine_is_on_curve(_p(a[i])) for i in range(len(a))], dtype=np.uint8)


def time_fe_mul(which, n, reps=5):
    return lib().jo_time_fe_mul(which, C.c_size_t(n), reps)


def time_scalar_mul(points, scalars32, nthreads, reps=1):
    points, scalars32 = _u64(points, 20), _u8(scalars32)
    out = np.empty_like(points)
    return lib().jo_time_scalar_mul(_p(points), _p(scalars32), _p(out), C.c_size_t(len(points)),
                                    nthreads, reps)
