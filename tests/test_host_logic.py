"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the
product fails loudly without a GPU, and the kernel arithmetic source (compiled for the host with
the PTX primitives emulated, tests/emul) matches the oracle limb for limb."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import model as M
from tests.helpers import edge_field_pairs, edge_field_values, scalar_bytes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FQ, FR = 0, 1


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g

    g.build()
    return g


def test_header_symbols_exported(built):
    from jubjub_b200 import _lib

    header = open(os.path.join(ROOT, "include", "jubjub_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(jj_[a-z0-9_]+)\s*\(", header)) - {"jj_ctx"})
    lib = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_lib.ALL_SYMBOLS) == declared  # the Python binding covers the whole header


def test_rust_shim_binds_the_whole_header():
    """integration/rust (uncompiled reference text) declares every ABI symbol of the header, with the header's
    argument counts, and the same flag / error constants."""
    header = open(os.path.join(ROOT, "include", "jubjub_b200.h")).read()
    rust = open(os.path.join(ROOT, "integration", "rust", "jubjub-b200", "src", "lib.rs")).read()
    decl = {m.group(1): m.group(2) for m in re.finditer(r"\b(jj_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", header, re.S)}
    decl.pop("jj_ctx", None)
    bound = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (jj_[a-z0-9_]+)\(([^;]*?)\)\s*(?:->[^;]*)?;", rust, re.S)}
    extra = {"jj_launch_count"}  # test / bench helper, not part of the documented surface
    assert sorted(set(decl) - extra) == sorted(set(bound) - extra)
    for name, args in decl.items():
        if name in extra:
            continue
        n_c = 0 if args.strip() in ("", "void") else args.count(",") + 1
        n_r = 0 if not bound[name].strip() else bound[name].count(",") + 1
        assert n_c == n_r, (name, n_c, n_r)
    for const, value in re.findall(r"\b(JJ_[A-Z0-9_]+)\s*=\s*([^,/\n]+)", header):
        m = re.search(r"pub const %s: [iu]32 = ([^;]+);" % const, rust)
        assert m, const
        assert eval(m.group(1)) == eval(value.replace("u", "")), const


def test_product_fails_loudly_without_gpu(built):
    import torch

    import jubjub_b200 as jj

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(jj.JubjubError) as e:
        jj.Engine(0)
    assert e.value.code == -5  # JJ_ERR_NO_DEVICE: there is no CPU fallback


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "jubjub_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                # no import, include, dlopen or link of anything under oracle/
                assert not re.search(r"import\s+oracle|from\s+oracle|#include[^\n]*oracle|jj_oracle|libjj_oracle", src), f


@pytest.fixture(scope="module")
def emul(built):
    lib = C.CDLL(os.path.join(ROOT, "tests", "emul", "libjj_emul.so"))

    class E:
        @staticmethod
        def fe(which, op, a, b=None):
            a = np.ascontiguousarray(a, dtype=np.uint64)
            b = a if b is None else np.ascontiguousarray(b, dtype=np.uint64)
            out = np.empty_like(a)
            lib.emul_fe_op(which, op, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                           out.ctypes.data_as(C.c_void_p), C.c_size_t(len(a)))
            return out

        @staticmethod
        def pt(op, p, q, width=20):
            out = np.empty((len(p), width), dtype=np.uint64)
            lib.emul_point_op(op, p.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p),
                              out.ctypes.data_as(C.c_void_p), C.c_size_t(len(p)))
            return out

        @staticmethod
        def smul(p, k):
            out = np.empty_like(p)
            lib.emul_scalar_mul(p.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p),
                                out.ctypes.data_as(C.c_void_p), C.c_size_t(len(p)))
            return out

        @staticmethod
        def smul_ct(p, k):
            out = np.empty_like(p)
            lib.emul_scalar_mul_ct(p.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p),
                                   out.ctypes.data_as(C.c_void_p), C.c_size_t(len(p)))
            return out

        @staticmethod
        def fixed_entries(base, w, first, count):
            tbl = np.zeros(lib.emul_fixed_table_words(w), dtype=np.uint32)
            lib.emul_fixed_table(base.ctypes.data_as(C.c_void_p), tbl.ctypes.data_as(C.c_void_p), w, first, count)
            return tbl.reshape(-1, 24)[first:first + count]

        @staticmethod
        def smul_fixed(tbl, k, w):
            tbl = np.ascontiguousarray(tbl, dtype=np.uint32)
            out = np.empty((len(k), 20), dtype=np.uint64)
            lib.emul_scalar_mul_fixed(tbl.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p),
                                      out.ctypes.data_as(C.c_void_p), C.c_size_t(len(k)), w)
            return out

        @staticmethod
        def table_words(w):
            return lib.emul_fixed_table_words(w)

    return E


@pytest.mark.parametrize("which", [FQ, FR])
def test_emulated_field_arithmetic(emul, oracle, which):
    m = M.Q if which == FQ else M.R_ORDER
    edge = np.array([M.limbs(x) for x in (0, 1, m - 1, m - 2, M.to_mont(1, m), M.to_mont(m - 1, m), (1 << 255) % m,
                                          2**32 - 1, 2**64 - 1, m >> 1, 1 << 32, 1 << 64, 1 << 224)], dtype=np.uint64)
    a = np.concatenate([edge, oracle.fe_stream(which, 1, 20000)])
    b = np.concatenate([edge[::-1], oracle.fe_stream(which, 2, 20000)])
    for op in range(6):
        assert (emul.fe(which, op, a, b) == oracle.fe_batch(which, op, a, b)).all(), op
    ea, eb = edge_field_pairs(m)  # fold boundary, zero / all-ones limbs, neighbours of 0, m/2, m -- all pairs
    for op in range(6):
        assert (emul.fe(which, op, ea, eb) == oracle.fe_batch(which, op, ea, eb)).all(), ("edge", op)
    for op, ref in ((10, 0), (11, 1)):  # the other reduction-row formulation (used by the elementwise kernels)
        assert (emul.fe(which, op, a, b) == oracle.fe_batch(which, ref, a, b)).all(), op
        assert (emul.fe(which, op, ea, eb) == oracle.fe_batch(which, ref, ea, eb)).all(), ("edge", op)
    assert (emul.fe(which, 6, a[:200]) == oracle.fe_invert(which, a[:200])[0]).all()
    assert (emul.fe(which, 7, a) == oracle.fe_to_bytes(which, a).view(np.uint64)).all()
    raw = np.concatenate([a, np.full((3, 4), 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)])
    assert (emul.fe(which, 8, raw) == oracle.fe_from_raw(which, raw)).all()
    assert (emul.fe(which, 9, raw)[:, 0] == oracle.fe_from_bytes(which, raw.view(np.uint8).reshape(-1, 32))[1]).all()


@pytest.mark.parametrize("which", [FQ, FR])
def test_emulated_field_arithmetic_vs_bigint_fuzz(emul, which):
    """Structured fuzzing of the kernel arithmetic source (host emulation) directly against Python big integers --
    independent of the C oracle: values built from a few powers of two +- small offsets, limbs forced to 0 or
    0xffffffff, neighbours of m, m/2, 2^k; every product / square formulation, add, sub, neg, double."""
    from hypothesis import given, settings, strategies as st

    m = M.Q if which == FQ else M.R_ORDER
    rinv = pow(1 << 256, -1, m)
    small = st.integers(-3, 3)
    pw = st.integers(0, 255)
    structured = st.builds(lambda ks, d, flip, base: (((sum(1 << k for k in ks) + d) ^ (flip * ((1 << 256) - 1))) + base) % m,
                           st.lists(pw, min_size=0, max_size=4), small, st.integers(0, 1),
                           st.sampled_from([0, m - 1, m >> 1, (m >> 1) + 1, m >> 225 << 224, (m >> 225 << 224) + (1 << 224)]))
    limbs = st.builds(lambda ws: sum(w << (32 * i) for i, w in enumerate(ws)) % m,
                      st.lists(st.sampled_from([0, 1, 0xFFFFFFFF, 0xFFFFFFFE, 0x80000000, 0x7FFFFFFF, 0x12345678]),
                               min_size=8, max_size=8))
    elem = st.one_of(structured, limbs, st.integers(0, m - 1))

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.tuples(elem, elem), min_size=64, max_size=64))
    def run(pairs):
        a = np.array([M.limbs(x) for x, _ in pairs], dtype=np.uint64)
        b = np.array([M.limbs(y) for _, y in pairs], dtype=np.uint64)
        want = {0: lambda x, y: x * y * rinv % m, 1: lambda x, y: x * x * rinv % m, 2: lambda x, y: (x + y) % m,
                3: lambda x, y: (x - y) % m, 4: lambda x, y: -x % m, 5: lambda x, y: 2 * x % m,
                10: lambda x, y: x * y * rinv % m, 11: lambda x, y: x * x * rinv % m}
        for op, f in want.items():
            got = emul.fe(which, op, a, b)
            for i, (x, y) in enumerate(pairs):
                assert M.from_limbs(got[i]) == f(x, y), (op, hex(x), hex(y))

    run()


def test_emulated_points_and_scalar_mul(emul, oracle):
    n = 48
    g = oracle.affine_to_extended(oracle.generator())
    t = oracle.fe_to_bytes(FR, oracle.fe_stream(FR, M.SEED0 + 3, n))
    p = oracle.scalar_mul(np.repeat(g, n, axis=0), t)
    q = oracle.ext_double(p[::-1].copy())
    assert (emul.pt(0, p, p) == oracle.ext_double(p)).all()
    assert (emul.pt(1, p, q) == oracle.ext_add(p, q)).all()
    assert (emul.pt(2, p, q) == oracle.ext_sub(p, q)).all()
    nq = oracle.ext_to_niels(q)
    assert (emul.pt(7, q, q, 16) == nq).all()
    assert (emul.pt(3, p, nq) == oracle.ext_add_niels(p, nq)).all()
    assert (emul.pt(4, p, nq) == oracle.ext_sub_niels(p, nq)).all()
    aq = oracle.ext_to_affine(q)
    anq = oracle.affine_to_niels(aq)
    assert (emul.pt(8, aq, aq, 12) == anq).all()
    assert (emul.pt(5, p, anq) == oracle.ext_add_affine_niels(p, anq)).all()
    assert (emul.pt(6, p, anq) == oracle.ext_sub_affine_niels(p, anq)).all()
    edge = [0, 1, 2, 7, 8, 9, 15, 16, M.R_ORDER - 1, M.R_ORDER, (1 << 252) - 1, (1 << 256) - 1, int("8" * 64, 16),
            int("7" * 64, 16)]
    k = np.concatenate([scalar_bytes(*edge), oracle.fe_to_bytes(FR, oracle.fe_stream(FR, M.SEED0 + 2, n))])
    pp = np.concatenate([np.repeat(p[:1], len(edge), axis=0), p])
    want = oracle.scalar_mul(pp, k)
    got = emul.smul(pp, k)
    assert oracle.ext_eq(got, want).all()
    assert (oracle.batch_normalize(got) == oracle.batch_normalize(want)).all()
    # constant-time mode (table scan, selects, every addition executed, identity for digit 0): the same points; the
    # projective representation may differ from the variable-time core's only where a zero digit added the identity
    got_ct = emul.smul_ct(pp, k)
    assert oracle.ext_eq(got_ct, want).all()
    assert (oracle.batch_normalize(got_ct) == oracle.batch_normalize(want)).all()
    want = oracle.batch_normalize(oracle.scalar_mul_fixed(oracle.generator(), k))
    for w in (4, 7, 12, 16):  # the window widths of the fixed-base tables (4, 7: shared memory; 12, 16: global memory)
        # the table the device kernel builds, computed here by the oracle: entry e = (j+1) * 2^(w*i) * G
        nw, per = -(-252 // w), 1 << (w - 1)
        mult = [(j + 1) << (w * i) for i in range(nw) for j in range(per)] + [1 << (w * nw)]
        # multiples >= 2^252 are outside multiply()'s 252-bit range (the top-carry entry 2^(w*nw), and for w = 16 the upper
        # part of the top window): build them as [m >> s] G doubled s times
        shift = [0 if m < (1 << 252) else m.bit_length() - 252 for m in mult]
        shift[-1] = w  # the top-carry entry: 2^(w*(nw-1)) G doubled w times
        sc = scalar_bytes(*[m >> s_ for m, s_ in zip(mult, shift)])
        ext = oracle.scalar_mul_fixed(oracle.generator(), sc)
        for s_ in range(1, max(shift) + 1):
            idx = np.flatnonzero(np.array(shift) >= s_)
            ext[idx] = oracle.ext_double(np.ascontiguousarray(ext[idx]))   # one more doubling for everything that owes >= s_
        tbl = oracle.affine_to_niels(oracle.batch_normalize(ext)).view(np.uint32).reshape(-1, 24)
        assert tbl.size == emul.table_words(w)
        for first in (0, per - 1, (nw // 2) * per + 3, nw * per - 2):  # incl. the top-carry entry
            got = emul.fixed_entries(oracle.generator(), w, first, 3)
            assert (got == tbl[first:first + 3][:len(got)]).all(), (w, first)
        assert (oracle.batch_normalize(emul.smul_fixed(tbl, k, w)) == want).all(), w


def test_emulated_shared_scalar_wnaf(built, oracle):
    """The width-5 NAF core behind is_torsion_free: [k]P for one k shared by the batch equals the oracle's ladder
    (affine), for k = r (42 non-zero digits; [r]P = O exactly on the prime-order subgroup), r - 1, 0, 1, 2^251 + ..."""
    lib = C.CDLL(os.path.join(ROOT, "tests", "emul", "libjj_emul.so"))
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    n = 12
    g = oracle.affine_to_extended(oracle.generator())
    t = oracle.fe_to_bytes(FR, oracle.fe_stream(FR, 91, n))
    p = oracle.scalar_mul(np.repeat(g, n, axis=0), t)          # full-order points
    p[n // 2:] = oracle.ext_mul_by_cofactor(p[n // 2:])         # ... and prime-order ones
    r = M.R_ORDER
    for k in (r, r - 1, 0, 1, 2, 31, 32, (1 << 251) + 0x5A5A5A5A5A5A5A5A, (1 << 252) - 1, (1 << 256) - 1):
        kb = scalar_bytes(k)
        out = np.zeros_like(p)
        nz = lib.emul_scalar_mul_wnaf(P(p), P(kb), P(out), C.c_size_t(n))
        want = oracle.scalar_mul(p, np.repeat(kb, n, axis=0))
        assert (oracle.batch_normalize(out) == oracle.batch_normalize(want)).all(), hex(k)
        if k == r:
            assert nz == 42
            aff = oracle.batch_normalize(out)
            one = oracle.fe_one(FQ)[0]
            assert (aff[n // 2:, :4] == 0).all() and (aff[n // 2:, 4:] == one).all()   # [r]P = (0, 1) on the subgroup
            assert not (aff[:n // 2, :4] == 0).all()                                     # but not for full-order points


def test_emulated_sqrt_and_decode(built, oracle):
    from tests.golden import reference_kats as K

    lib = C.CDLL(os.path.join(ROOT, "tests", "emul", "libjj_emul.so"))
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    assert lib.emul_fq_sqrt_tables_ok() == 1  # the 13-bit hash is perfect on the order-256 subgroup
    # random elements, 0, squares of the edge values, and the whole 2^32-torsion ladder g^(2^j), g^(3 * 2^j)
    # (g = 7^T: the inputs whose a^T is not 1, i.e. the ones that exercise every byte of the logarithm)
    edge = edge_field_values(M.Q)
    T = (M.Q - 1) >> 32
    g = pow(7, T, M.Q)
    tors = [pow(g, (c << j) % (1 << 32), M.Q) for j in range(32) for c in (1, 3, 0xFFFFFFFF, 0x9E3779B1)]
    tors += [x * 5 % M.Q for x in tors[:64]] + [x * x * 11 % M.Q for x in tors[:64]]
    tors_m = np.array([M.limbs(M.to_mont(x, M.Q)) for x in tors], dtype=np.uint64)
    a = np.concatenate([oracle.fe_stream(FQ, 77, 600), edge, oracle.fe_batch(FQ, oracle.OP_SQUARE, edge), tors_m])
    a[0] = 0
    want, wok = oracle.fe_sqrt(FQ, a)
    for fn in (lib.emul_fq_sqrt, lib.emul_fq_sqrt_ts):
        out, ok = np.zeros_like(a), np.zeros(len(a), np.uint8)
        fn(P(a), P(out), P(ok), C.c_size_t(len(a)))
        assert (ok == wok).all() and 0.3 * len(a) < ok.sum() < len(a)
        assert (oracle.fe_batch(FQ, oracle.OP_SQUARE, out[ok == 1]) == a[ok == 1]).all()
        assert (out[ok == 1] == want[wok == 1]).all()  # the same root as the oracle's Tonelli-Shanks
    # sqrt(num / den) in one power (the inversion-free decode): residuosity flag and root^2 * den == num, on the same
    # inputs paired with random denominators, with the torsion ladder also used as denominators, and num = 0
    den = np.concatenate([oracle.fe_stream(FQ, 78, len(a) - len(tors_m)), tors_m[::-1]])
    num = oracle.fe_batch(FQ, oracle.OP_MUL, a, den)  # num / den = a
    num2, den2 = np.concatenate([num, a]), np.concatenate([den, den])  # second half: ratio a / den, residuosity unknown
    q2, q2ok = oracle.fe_invert(FQ, den2)
    assert q2ok.all()
    ratio = oracle.fe_batch(FQ, oracle.OP_MUL, num2, q2)
    _, rok = oracle.fe_sqrt(FQ, ratio)
    out, ok = np.zeros_like(num2), np.zeros(len(num2), np.uint8)
    lib.emul_fq_sqrt_ratio(P(num2), P(den2), P(out), P(ok), C.c_size_t(len(num2)))
    assert (ok == rok).all() and (ok[:len(a)] == wok).all()
    assert (oracle.fe_batch(FQ, oracle.OP_SQUARE, out[ok == 1]) == ratio[ok == 1]).all()
    assert ok[0] == 1 and not out[0].any() and not out[ok == 0].any()
    g = oracle.affine_to_extended(oracle.generator())
    t = oracle.fe_to_bytes(FR, oracle.fe_stream(FR, 5, 120))
    enc = oracle.affine_to_bytes(oracle.batch_normalize(oracle.scalar_mul(np.repeat(g, 120, axis=0), t)))
    bad = enc[:40].copy()
    bad[:, 0] ^= 1
    extra = np.array(K.SERIALIZED_MULTIPLES_OF_8G + K.ZIP216_NON_CANONICAL + [[0xFF] * 32, [1] + [0] * 31, [0] * 32],
                     dtype=np.uint8)
    allenc = np.concatenate([enc, bad, extra])
    for zip216 in (1, 0):
        got, ok = np.zeros((len(allenc), 8), np.uint64), np.zeros(len(allenc), np.uint8)
        lib.emul_from_bytes(P(allenc), P(got), P(ok), zip216, C.c_size_t(len(allenc)))
        want, wok = oracle.affine_from_bytes(allenc, zip216=bool(zip216))
        assert (ok == wok).all() and (got[ok == 1] == want[wok == 1]).all()


def _torsion_cosets(oracle, n, seed):
    """n prime-order points (random projective scaling) and, for each, its eight cosets P + T_j over the reference's
    8-torsion table (src/lib.rs:1589-1677): (8n, 20) extended points, coset j of point i at row 8i + j."""
    from tests.golden import reference_kats as K
    from tests.helpers import affine_raw

    g = oracle.affine_to_extended(oracle.generator())
    t = oracle.fe_to_bytes(FR, oracle.fe_stream(FR, seed, n))
    p = oracle.ext_mul_by_cofactor(oracle.scalar_mul(np.repeat(g, n, axis=0), t))       # prime order, z != 1
    tors = oracle.affine_to_extended(affine_raw(oracle, K.EIGHT_TORSION_RAW))           # T_0 .. T_7 (T_7 or so = O)
    pts = np.concatenate([oracle.ext_add(np.repeat(p[i:i + 1], 8, axis=0), tors) for i in range(n)])
    return p, tors, pts


def test_emulated_pairing_torsion_check(built, oracle):
    """The pairing-based is_torsion_free (csrc/torsion.cuh, host emulation of the same source) gives the reference's
    answer, [r]P == O (src/lib.rs:709-711), on every coset of the 8-torsion, on the small-order points themselves,
    on the identity in several projective forms and on random full-order points."""
    lib = C.CDLL(os.path.join(ROOT, "tests", "emul", "libjj_emul.so"))
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731

    def flags(p):
        p = np.ascontiguousarray(p)
        out = np.zeros(len(p), np.uint8)
        lib.emul_is_torsion_free(P(p), P(out), C.c_size_t(len(p)))
        return out

    prime, tors, cosets = _torsion_cosets(oracle, 24, 4242)
    want = oracle.is_torsion_free(cosets)
    assert want.reshape(-1, 8).sum(axis=1).tolist() == [1] * 24   # exactly one coset of each point is torsion free
    assert (flags(cosets) == want).all()
    assert (flags(prime) == 1).all()
    assert (flags(tors) == oracle.is_torsion_free(tors)).all() and flags(tors).sum() == 1
    # doubling changes the projective representation (z != 1, t1 * t2 != u * v / z form): flags are unchanged
    d = oracle.ext_double(cosets)
    assert (flags(d) == oracle.is_torsion_free(d)).all()
    g = oracle.affine_to_extended(oracle.generator())
    t = oracle.fe_to_bytes(FR, oracle.fe_stream(FR, 99, 64))
    full = oracle.scalar_mul(np.repeat(g, 64, axis=0), t)
    assert (flags(full) == oracle.is_torsion_free(full)).all()
    # the identity and the point of order two (0, -1) in scaled projective forms (0, z, z, ..) / (0, -z, z, ..): U = 0
    z = oracle.fe_stream(FQ, 7, 3)
    ident = oracle.identity(3)
    ident[:, 4:8], ident[:, 8:12], ident[:, 16:20] = z, z, z
    assert (oracle.is_torsion_free(ident) == 1).all() and (flags(ident) == 1).all()
    two = ident.copy()
    two[:, 4:8] = oracle.fe_batch(FQ, oracle.OP_NEG, z)
    assert (oracle.is_torsion_free(two) == 0).all() and (flags(two) == 0).all()


def test_build_records_the_validated_nvcc(built):
    """The carry-chain arithmetic was validated (GPU fuzz wall) for one ptxas only: build() records the nvcc it used
    and refuses any other; a library built by an unvalidated toolchain must not pass silently."""
    import json

    info = json.load(open(os.path.join(ROOT, "jubjub_b200", "build_info.json")))
    assert info["nvcc"] == built.VALIDATED_NVCC, info
    # the library on disk was built from exactly the sources on disk (content hash, not file times)
    csrc = os.path.join(ROOT, "jubjub_b200", "csrc")
    srcs = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc))] + [os.path.join(ROOT, "include", "jubjub_b200.h")]
    assert info["sources_sha256"] == built._sources_digest(srcs, " ".join(info["flags"]))
    assert "-lineinfo" in info["flags"] and "arch=compute_100a,code=sm_100a" in info["flags"]


def test_types_star_import_and_vartime_gate():
    """`from jubjub_b200.types import *` works and the `*` operator on points is gated behind an explicit
    acknowledgement that the batch kernels are variable-time (the reference's `*` is constant-time, src/lib.rs:12-17)."""
    ns = {}
    exec("from jubjub_b200.types import *", ns)
    for name in ("Fq", "Fr", "ExtendedPoint", "batch_mul_vartime", "acknowledge_vartime"):
        assert name in ns, name
    import jubjub_b200.types as T

    assert hasattr(T.ExtendedPoint, "mul_vartime") and hasattr(T.AffinePoint, "mul_vartime")
    import jubjub_b200 as jj

    for name in ("scalar_mul_vartime", "scalar_mul_fixed_vartime", "scalar_mul_encoded_vartime", "scalar_mul_sharded_vartime"):
        assert hasattr(jj.Engine, name), name
    # the unsuffixed name is the constant-time-in-the-scalar mode, like the reference's `*`
    assert "JJ_CONST_TIME" in jj.Engine.scalar_mul.__code__.co_names or "constant" in (jj.Engine.scalar_mul.__doc__ or "")
    for name in ("scalar_mul_fixed", "scalar_mul_encoded", "scalar_mul_sharded"):
        assert not hasattr(jj.Engine, name), name  # no constant-time mode there: only the *_vartime names exist


def test_constant_time_kernel_has_no_scalar_dependent_branch(built):
    """Structural check of JJ_CONST_TIME on the shipped SASS: inside the window loop of the constant-time kernel there are
    exactly three loop back-edges (doublings, 8-entry table scan, 63 windows), the calls of the shared Fq product, and
    no BSSY / BSYNC reconvergence pair -- ptxas brackets every branch that may diverge with one, so the only other
    branches left are warp-uniform ones on the (public) loop counter: the first pass doubles once and takes its digit from
    the top part of the scalar by a select.  The table loads sit in the scan loop (4 per iteration), not behind a
    digit-indexed address.  The variable-time kernel, for contrast, has the divergent `if (digit != 0)`."""
    import subprocess

    from jubjub_b200 import _lib

    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        m = re.match(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur is not None:
            cur.append((int(m.group(1), 16), m.group(2).strip()))

    def window_loop(name_part):
        (ins,) = [v for k, v in funcs.items() if name_part in k]
        back = []
        for a, t in ins:
            m = re.match(r"(?:@!?U?P\d+\s+)?BRA\s+(?:!?P\d+,\s*)?0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                back.append((int(m.group(1), 16), a))
        # the window loop: the innermost loop that contains the 8 calls of one addition
        loops = sorted(back, key=lambda r: r[1] - r[0])
        for lo, hi in loops:
            body = [(a, t) for a, t in ins if lo <= a <= hi]
            if sum("CALL" in t for _, t in body) == 8:
                return body, [(l, h) for l, h in back if lo <= l and h <= hi]
        raise AssertionError("window loop not found")

    ct, ct_back = window_loop("k_scalar_mulILi512ELi1ELi1ELb0ELb1")
    vt, _ = window_loop("k_scalar_mulILi512ELi1ELi1ELb0ELb0")
    assert not any(re.search(r"\b(BSSY|BSYNC|WARPSYNC)\b", t) for _, t in ct)
    assert len(ct_back) == 3, ct_back                           # doubling loop, scan loop, window loop
    assert sum(1 for _, t in ct if re.search(r"\bBRA\b", t)) <= 5   # + at most two uniform forward branches on the counter
    scan = min((r for r in ct_back), key=lambda r: r[1] - r[0])
    loads = [(a, t) for a, t in ct if "LDG" in t]
    assert len(loads) == 4 and all(scan[0] <= a <= scan[1] for a, _ in loads)
    assert any(re.search(r"\bBSSY\b", t) for _, t in vt)        # the variable-time kernel does branch on the digit
