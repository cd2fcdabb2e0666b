"""Parity of the CUDA path (through the C ABI) against the CPU oracle -- run on the B200 box:
    python -m pytest tests -m gpu
Bit-exact everywhere: Montgomery limbs for field ops, all 160 bytes for point add/double
(same formula sequence as the reference), normalised affine / 32-byte encodings for scalar-mul
(windowed vs the reference's bitwise ladder differ projectively only)."""
import numpy as np
import pytest

from oracle import model as M
from tests.golden import reference_kats as K
from tests.helpers import affine_raw, b32, edge_field_pairs, fe, scalar_bytes

pytestmark = pytest.mark.gpu

FQ, FR = 0, 1


@pytest.fixture(scope="module")
def eng():
    import jubjub_b200 as jj

    e = jj.Engine(0)
    yield e
    e.close()


def _edge_fe(m):
    vals = [0, 1, 2, m - 1, m - 2, M.to_mont(1, m), M.to_mont(m - 1, m), (1 << 255) % m, 2**32 - 1, 2**64 - 1,
            m >> 1, 1 << 32, 1 << 64, 1 << 224, (1 << 128) - 1]
    return np.array([M.limbs(v) for v in vals], dtype=np.uint64)


def _field_inputs(oracle, which, n=20000):
    m = M.Q if which == FQ else M.R_ORDER
    e = _edge_fe(m)
    a = np.concatenate([e, oracle.fe_stream(which, M.SEED0, n)])
    b = np.concatenate([e[::-1], oracle.fe_stream(which, M.SEED0 + 1, n)])
    return a, b


@pytest.mark.parametrize("which,name", [(FQ, "fq"), (FR, "fr")])
def test_field_ops_montgomery(eng, oracle, which, name):
    a, b = _field_inputs(oracle, which)
    assert (eng.fe_mul(name, a, b) == oracle.fe_batch(which, oracle.OP_MUL, a, b)).all()
    assert (eng.fe_add(name, a, b) == oracle.fe_batch(which, oracle.OP_ADD, a, b)).all()
    assert (eng.fe_sub(name, a, b) == oracle.fe_batch(which, oracle.OP_SUB, a, b)).all()
    assert (eng.fe_square(name, a) == oracle.fe_batch(which, oracle.OP_SQUARE, a)).all()
    assert (eng.fe_neg(name, a) == oracle.fe_batch(which, oracle.OP_NEG, a)).all()
    assert (eng.fe_double(name, a) == oracle.fe_batch(which, oracle.OP_DOUBLE, a)).all()


@pytest.mark.parametrize("which,name", [(FQ, "fq"), (FR, "fr")])
def test_field_ops_edge_pairs(eng, oracle, which, name):
    """All ordered pairs of the adversarial values (fold boundary of the squaring, zero / all-ones limbs that
    make the quotient digit 0 or carry out of every column, neighbours of 0, m/2 and m)."""
    a, b = edge_field_pairs(M.Q if which == FQ else M.R_ORDER)
    assert (eng.fe_mul(name, a, b) == oracle.fe_batch(which, oracle.OP_MUL, a, b)).all()
    assert (eng.fe_add(name, a, b) == oracle.fe_batch(which, oracle.OP_ADD, a, b)).all()
    assert (eng.fe_sub(name, a, b) == oracle.fe_batch(which, oracle.OP_SUB, a, b)).all()
    assert (eng.fe_square(name, a) == oracle.fe_batch(which, oracle.OP_SQUARE, a)).all()
    assert (eng.fe_neg(name, a) == oracle.fe_batch(which, oracle.OP_NEG, a)).all()
    assert (eng.fe_double(name, a) == oracle.fe_batch(which, oracle.OP_DOUBLE, a)).all()
    # values, not only agreement of two implementations: a sample against the big-integer model
    m = M.Q if which == FQ else M.R_ORDER
    rinv = pow(1 << 256, -1, m)
    got = eng.fe_mul(name, a, b)
    for i in range(0, len(a), 53):
        assert M.from_limbs(got[i]) == M.from_limbs(a[i]) * M.from_limbs(b[i]) * rinv % m


@pytest.mark.parametrize("which,name", [(FQ, "fq"), (FR, "fr")])
def test_field_ops_canonical_flag(eng, oracle, which, name):
    import jubjub_b200 as jj

    m = M.Q if which == FQ else M.R_ORDER
    a, b = _field_inputs(oracle, which, 2000)
    ca, cb = oracle.fe_to_bytes(which, a).view(np.uint64), oracle.fe_to_bytes(which, b).view(np.uint64)
    got = eng.fe_mul(name, ca, cb, flags=jj.JJ_CANON)
    for i in range(0, len(a), 97):
        assert M.from_limbs(got[i]) == M.from_limbs(ca[i]) * M.from_limbs(cb[i]) % m
    want = oracle.fe_to_bytes(which, oracle.fe_batch(which, oracle.OP_MUL, a, b)).view(np.uint64)
    assert (got == want).all()
    got = eng.fe_sub(name, ca, cb, flags=jj.JJ_CANON)
    assert (got == oracle.fe_to_bytes(which, oracle.fe_batch(which, oracle.OP_SUB, a, b)).view(np.uint64)).all()


@pytest.mark.parametrize("which,name", [(FQ, "fq"), (FR, "fr")])
def test_field_invert_and_bytes(eng, oracle, which, name):
    m = M.Q if which == FQ else M.R_ORDER
    a, _ = _field_inputs(oracle, which, 3000)
    inv, ok = eng.fe_invert(name, a)
    winv, wok = oracle.fe_invert(which, a)
    assert (ok == wok).all() and ok[0] == 0  # zero has no inverse (CtOption::none, src/fr.rs:539)
    assert (inv == winv).all()
    assert (eng.fe_to_bytes(name, a) == oracle.fe_to_bytes(which, a)).all()
    # from_bytes: canonical accepted, >= m rejected (src/fr.rs:925-960)
    enc = np.concatenate([oracle.fe_to_bytes(which, a),
                          scalar_bytes(m, m + 1, (1 << 256) - 1, m - 1, 0)])
    got, ok = eng.fe_from_bytes(name, enc)
    want, wok = oracle.fe_from_bytes(which, enc)
    assert (ok == wok).all() and ok[-5:].tolist() == [0, 0, 0, 1, 1]
    assert (got[ok == 1] == want[ok == 1]).all()


@pytest.mark.parametrize("which,name", [(FQ, "fq"), (FR, "fr")])
def test_field_invert_long_chains(eng, oracle, which, name):
    """More elements than resident threads: every thread inverts a chain of several elements with Montgomery's
    trick (zeros scattered through the chains are skipped and flagged, like ff::BatchInverter); also with
    canonical-integer I/O and in place on the device."""
    import jubjub_b200 as jj

    n = 2 * 37888 + 517
    a = oracle.fe_stream(which, M.SEED0 + 11, n)
    a[::97] = 0
    a[1::1013] = oracle.fe_one(which)[0]
    inv, ok = eng.fe_invert(name, a)
    winv, wok = oracle.fe_invert(which, a)
    assert (ok == wok).all() and 700 < int((ok == 0).sum()) <= len(a[::97])
    assert (inv == winv).all()
    one = np.repeat(oracle.fe_one(which), n, axis=0)
    assert (eng.fe_mul(name, a, inv)[ok == 1] == one[ok == 1]).all()
    ca = oracle.fe_to_bytes(which, a).view(np.uint64)
    cinv, cok = eng.fe_invert(name, ca, flags=jj.JJ_CANON)
    assert (cok == wok).all() and (cinv == oracle.fe_to_bytes(which, winv).view(np.uint64)).all()
    d = eng.to_device(a)
    okd = eng.empty((n, 1), np.uint8)
    eng._check(getattr(eng.lib, f"jj_{name}_invert")(eng.ctx, d.ptr, d.ptr, okd.ptr, n, jj.JJ_DEVICE_PTRS))  # out aliases a
    assert (d.download() == winv).all() and (okd.download().ravel() == wok).all()


@pytest.mark.parametrize("which,name", [(FQ, "fq"), (FR, "fr")])
def test_field_sqrt(eng, oracle, which, name):
    """src/fr.rs:1205-1227 (47 non-residues among r-2, r-3, ...) and residue flags vs the oracle for both fields."""
    a, _ = _field_inputs(oracle, which, 2000)
    root, ok = eng.fe_sqrt(name, a)
    _, wok = oracle.fe_sqrt(which, a)
    assert (ok == wok).all() and 0 < ok.sum() < len(ok)
    assert (eng.fe_square(name, root[ok == 1]) == a[ok == 1]).all() and (root[ok == 0] == 0).all()
    if which == FR:
        sq, vals = fe(K.FR_R_MINUS_2), []
        for _ in range(100):
            vals.append(sq[0].copy())
            sq = eng.fe_sub("fr", sq, oracle.fe_one(FR))
        _, ok = eng.fe_sqrt("fr", np.array(vals))
        assert int((ok == 0).sum()) == K.FR_SQRT_NONE_COUNT


def test_fq_sqrt_torsion_ladder(eng, oracle):
    """Fq::sqrt is table-driven on the device (logarithm in the 2^32-torsion, 8 bits at a time): inputs whose
    a^T runs over the whole torsion ladder g^(c * 2^j), scaled copies, the edge values and their squares must give
    the oracle's residue flags AND the oracle's root (Tonelli-Shanks, the same root choice)."""
    from tests.helpers import edge_field_values

    T = (M.Q - 1) >> 32
    g = pow(7, T, M.Q)
    tors = [pow(g, (c << j) % (1 << 32), M.Q) for j in range(32) for c in (1, 3, 0xFFFFFFFF, 0x9E3779B1, 0x80000001)]
    tors += [x * 5 % M.Q for x in tors[:96]] + [x * x * 11 % M.Q for x in tors[:96]]
    edge = edge_field_values(M.Q)
    a = np.concatenate([np.array([M.limbs(M.to_mont(x, M.Q)) for x in tors], dtype=np.uint64), edge,
                        oracle.fe_batch(FQ, oracle.OP_SQUARE, edge), oracle.fe_stream(FQ, M.SEED0 + 9, 20000)])
    root, ok = eng.fe_sqrt("fq", a)
    want, wok = oracle.fe_sqrt(FQ, a)
    assert (ok == wok).all() and 0.3 * len(a) < ok.sum() < len(a)
    assert (root[ok == 1] == want[wok == 1]).all() and (root[ok == 0] == 0).all()


def test_fr_reference_kats_on_gpu(eng, oracle):
    """src/fr.rs:1045-1099 (LARGEST add/neg/sub), :1024-1034 (wide max), :1758-1776 (a*b==c)."""
    big, one_raw = fe(K.FR_LARGEST), fe([1, 0, 0, 0])
    assert (eng.fe_add("fr", big, big) == fe(K.FR_LARGEST_PLUS_LARGEST)).all()
    assert (eng.fe_add("fr", big, one_raw) == 0).all()
    assert (eng.fe_neg("fr", big) == one_raw).all()
    assert (eng.fe_neg("fr", fe([0, 0, 0, 0])) == 0).all()
    assert (eng.fe_sub("fr", big, big) == 0).all()
    assert (eng.fe_mul("fr", fe(K.MULC_A), fe(K.MULC_B)) == fe(K.MULC_C)).all()
    assert (eng.fe_from_bytes_wide("fr", np.full((1, 64), 0xFF, np.uint8)) == fe(K.FR_WIDE_MAX_MONT)).all()
    assert (eng.fe_to_bytes("fr", fe(K.FR_R2)) == b32(K.FR_BYTES_R2)).all()
    _, ok = eng.fe_from_bytes("fr", b32(*K.FR_BYTES_REJECTED))
    assert ok.tolist() == [0, 0, 0, 0]


def test_fq_bench_vectors(eng, oracle):
    """benches/fq_bench.rs:25-33: n *= -1 alternates between -1 and 1."""
    one = oracle.fe_one(FQ)
    neg_one = eng.fe_neg("fq", one)
    n = one
    for i in range(6):
        n = eng.fe_mul("fq", n, neg_one)
        assert (n == (neg_one if i % 2 == 0 else one)).all()
    four = eng.fe_double("fq", eng.fe_double("fq", one))
    inv, ok = eng.fe_invert("fq", four)
    assert ok[0] == 1 and (eng.fe_mul("fq", inv, four) == one).all()


@pytest.mark.parametrize("which,name", [(FQ, "fq"), (FR, "fr")])
def test_stream_and_wide(eng, oracle, which, name):
    n = 5000
    assert (eng.fe_stream(name, M.SEED0 + 2, n, first=12345) == oracle.fe_stream(which, M.SEED0 + 2, n, first=12345)).all()
    rng = np.random.default_rng(7)
    wide = rng.integers(0, 256, size=(n, 64), dtype=np.uint8)
    wide[0] = 0xFF
    wide[1] = 0
    assert (eng.fe_from_bytes_wide(name, wide) == oracle.fe_from_bytes_wide(which, wide)).all()


# ------------------------------------------------------------------------------ points
def _points(oracle, n, seed=M.SEED0 + 3):
    """n full-order points [t_i] G as extended with z = 1, plus identity and the 8-torsion points."""
    t = oracle.fe_to_bytes(FR, oracle.fe_stream(FR, seed, n))
    g = oracle.affine_to_extended(oracle.generator())
    rnd = oracle.scalar_mul(np.repeat(g, n, axis=0), t)
    tors = oracle.affine_to_extended(affine_raw(oracle, K.EIGHT_TORSION_RAW))
    return np.concatenate([oracle.identity(), tors, rnd])


def test_point_ops_bit_exact(eng, oracle):
    p = _points(oracle, 600)
    q = oracle.ext_double(p[::-1].copy())  # non-trivial z
    assert (eng.point_double(p) == oracle.ext_double(p)).all()
    assert (eng.point_double(q) == oracle.ext_double(q)).all()
    assert (eng.point_add(q, p) == oracle.ext_add(q, p)).all()
    assert (eng.point_add(q, p, subtract=True) == oracle.ext_sub(q, p)).all()
    nq = oracle.ext_to_niels(q)
    assert (eng.point_to_niels(q) == nq).all()
    assert (eng.point_add_niels(p, nq) == oracle.ext_add_niels(p, nq)).all()
    assert (eng.point_add_niels(p, nq, subtract=True) == oracle.ext_sub_niels(p, nq)).all()
    an = oracle.affine_to_niels(oracle.batch_normalize(q))
    assert (eng.affine_to_niels(oracle.batch_normalize(q)) == an).all()
    assert (eng.point_add_affine_niels(p, an) == oracle.ext_add_affine_niels(p, an)).all()
    assert (eng.point_add_affine_niels(p, an, subtract=True) == oracle.ext_sub_affine_niels(p, an)).all()


def test_point_bench_vectors(eng, oracle):
    """benches/point_bench.rs: identity doubling / identity + (-identity) / cached adds."""
    ident = oracle.identity()
    assert oracle.is_identity(eng.point_double(ident))[0]
    assert oracle.is_identity(eng.point_add(ident, oracle.ext_neg(ident)))[0]
    assert (eng.point_add_niels(ident, oracle.ext_to_niels(ident)) == oracle.ext_add_niels(ident, oracle.ext_to_niels(ident))).all()
    an = oracle.affine_to_niels(oracle.ext_to_affine(ident))
    assert (eng.point_add_affine_niels(ident, an) == oracle.ext_add_affine_niels(ident, an)).all()


EDGE_SCALARS = [0, 1, 2, 7, 8, 9, 15, 16, 17, M.R_ORDER - 1, M.R_ORDER, M.R_ORDER + 1, (1 << 252) - 1, (1 << 252),
                (1 << 256) - 1, int("8" * 64, 16), int("7" * 64, 16), int("f" * 63, 16), 1 << 251, 0x88, 0x80]


def _smul_case(oracle, n):
    p = _points(oracle, n)
    k = np.concatenate([scalar_bytes(*EDGE_SCALARS),
                        oracle.fe_to_bytes(FR, oracle.fe_stream(FR, M.SEED0 + 2, len(p)))])[: len(p)]
    return p, k


# what ships: the default mapping (0 = 13), the two A/B mappings (24, 5 warps-per-SM variants), and the default mapping
# with the fused normalise epilogue forced on / off for converted outputs (200 / 201)
@pytest.mark.parametrize("variant", [0, 13, 24, 5, 200, 201])
def test_scalar_mul_variants(eng, oracle, variant):
    eng.set_scalar_mul_variant(variant)
    try:
        p, k = _smul_case(oracle, 1200 if variant in (0, 200) else 300)
        want_ext = oracle.scalar_mul(p, k)
        want_aff = oracle.batch_normalize(want_ext)
        got = eng.scalar_mul_vartime(p, k)
        assert oracle.ext_eq(got, want_ext).all()  # projective equality, src/lib.rs:153-163
        assert (oracle.batch_normalize(got) == want_aff).all()
        assert (eng.scalar_mul_vartime(p, k, output="affine") == want_aff).all()
        assert (eng.scalar_mul_vartime(p, k, output="bytes") == oracle.affine_to_bytes(want_aff)).all()
    finally:
        eng.set_scalar_mul_variant(0)


def test_scalar_mul_constant_time_mode(eng, oracle):
    """JJ_CONST_TIME (Engine.scalar_mul, and `point * scalar` of the reference-style types): same results as the oracle's
    constant-time ladder for every edge scalar, in every output format, host and device resident."""
    p, k = _smul_case(oracle, 3000)
    want_ext = oracle.scalar_mul(p, k)
    want_aff = oracle.batch_normalize(want_ext)
    got = eng.scalar_mul(p, k)
    assert oracle.ext_eq(got, want_ext).all() and (oracle.batch_normalize(got) == want_aff).all()
    assert (eng.scalar_mul(p, k, output="affine") == want_aff).all()
    assert (eng.scalar_mul(p, k, output="bytes") == oracle.affine_to_bytes(want_aff)).all()
    assert (eng.scalar_mul(eng.to_device(p), eng.to_device(k), output="bytes").download() == oracle.affine_to_bytes(want_aff)).all()
    s = oracle.fe_stream(FR, 123, len(p))
    assert (eng.scalar_mul(p, s, scalar_mont=True, output="affine")
            == oracle.batch_normalize(oracle.scalar_mul(p, oracle.fe_to_bytes(FR, s)))).all()
    # all-zero scalars: 63 identity additions
    z = np.zeros((64, 32), np.uint8)
    assert oracle.is_identity(eng.scalar_mul(p[:64], z)).all()


def test_scalar_mul_vs_bigint_model(eng, oracle):
    p, k = _smul_case(oracle, 40)
    vals = oracle.fe_to_bytes(FQ, oracle.batch_normalize(p).reshape(-1, 4)).reshape(-1, 64)
    got = eng.scalar_mul_vartime(p, k, output="bytes")
    for i in range(len(p)):
        u = int.from_bytes(bytes(vals[i, :32]), "little")
        v = int.from_bytes(bytes(vals[i, 32:]), "little")
        assert bytes(got[i]) == M.encode(M.pmul((u, v), M.scalar_from_bytes_ref(bytes(k[i])))), i


def test_scalar_mul_montgomery_scalars(eng, oracle):
    """`&ExtendedPoint * &Fr` takes the scalar in Montgomery form and calls to_bytes (src/lib.rs:877)."""
    p = _points(oracle, 200)
    s = oracle.fe_stream(FR, 99, len(p))
    got = eng.scalar_mul_vartime(p, s, scalar_mont=True, output="affine")
    assert (got == oracle.batch_normalize(oracle.scalar_mul(p, oracle.fe_to_bytes(FR, s)))).all()


def test_mul_consistency_and_assoc(eng, oracle):
    """src/lib.rs:1505-1527, :1757-1804 with the reference's fixed raw-limb scalars."""
    p = oracle.ext_mul_by_cofactor(oracle.affine_to_extended(affine_raw(oracle, [K.TEST_POINT_RAW])))
    a, b, c = fe(K.MULC_A), fe(K.MULC_B), fe(K.MULC_C)
    assert (eng.fe_mul("fr", a, b) == c).all()
    pc = eng.scalar_mul_vartime(p, c, scalar_mont=True)
    pab = eng.scalar_mul_vartime(eng.scalar_mul_vartime(p, a, scalar_mont=True), b, scalar_mont=True)
    assert oracle.ext_eq(pc, pab)[0]
    assert (eng.batch_normalize(pc) == eng.batch_normalize(pab)).all()
    aff = eng.batch_normalize(p)
    assert (eng.scalar_mul_fixed_vartime(aff, c, scalar_mont=True, output="affine") == eng.batch_normalize(pab)).all()
    lhs = eng.scalar_mul_vartime(eng.scalar_mul_vartime(p, scalar_bytes(1000)), scalar_bytes(3938))
    assert oracle.ext_eq(lhs, eng.scalar_mul_vartime(p, scalar_bytes(3938000)))[0]


# 0: 12-bit windows, 4.1 MB table in global memory (default); 116: 16-bit windows, 50 MB; 107: 7-bit windows, 216 KB in shared
# memory; 100: 4-bit, 47 KB
@pytest.mark.parametrize("variant", [0, 116, 107, 100])
def test_scalar_mul_fixed(eng, oracle, variant):
    eng.set_scalar_mul_variant(variant)
    try:
        _check_fixed(eng, oracle)
    finally:
        eng.set_scalar_mul_variant(0)


def _check_fixed(eng, oracle):
    k = np.concatenate([scalar_bytes(*EDGE_SCALARS), oracle.fe_to_bytes(FR, oracle.fe_stream(FR, 5, 2000))])
    for base in (oracle.generator(), oracle.ext_to_affine(oracle.ext_mul_by_cofactor(oracle.affine_to_extended(oracle.generator())))):
        want = oracle.batch_normalize(oracle.scalar_mul_fixed(base, k))
        assert (eng.scalar_mul_fixed_vartime(base, k, output="affine") == want).all()
        assert (eng.batch_normalize(eng.scalar_mul_fixed_vartime(base, k)) == want).all()
        assert (eng.scalar_mul_fixed_vartime(base, k, output="bytes") == oracle.affine_to_bytes(want)).all()


def test_serialization_golden(eng, oracle):
    """src/lib.rs:1807-1890: encodings of k * (8 G), k = 1..16."""
    g8 = oracle.ext_mul_by_cofactor(oracle.affine_to_extended(oracle.generator()))
    want = b32(*K.SERIALIZED_MULTIPLES_OF_8G)
    k = scalar_bytes(*range(1, 17))
    assert (eng.scalar_mul_vartime(np.repeat(g8, 16, axis=0), k, output="bytes") == want).all()
    assert (eng.scalar_mul_fixed_vartime(oracle.ext_to_affine(g8), k, output="bytes") == want).all()
    # and by repeated addition, as the reference test walks it
    p = g8
    for i in range(16):
        assert (eng.affine_to_bytes(eng.batch_normalize(p))[0] == want[i]).all(), i
        p = eng.point_add(p, g8)


def test_batch_normalize_and_flags(eng, oracle):
    p = _points(oracle, 700)
    q = oracle.ext_double(oracle.ext_double(p))
    q[5, 8:12] = 0  # z = 0 is skipped like ff::BatchInverter and yields (0, 0)
    want = oracle.batch_normalize(q)
    got = eng.batch_normalize(q)
    assert (got == want).all() and (got[5] == 0).all()
    assert (eng.affine_to_bytes(got) == oracle.affine_to_bytes(want)).all()
    # GroupEncoding::to_bytes for ExtendedPoint (src/lib.rs:1419-1421): normalise + encode in one pass (JJ_OUT_BYTES),
    # host buffers (pageable and page-locked: stored in place) and device resident
    wenc = oracle.affine_to_bytes(want)
    assert (eng.batch_normalize_to_bytes(q) == wenc).all()
    assert (eng.batch_normalize_to_bytes(eng.to_device(q)).download() == wenc).all()
    pin = eng.pinned_empty((len(q), 32), np.uint8)
    pin[:] = 0xA5
    assert eng.batch_normalize_to_bytes(q, out=pin) is pin and (pin == wenc).all()
    del pin
    # src/lib.rs:1530-1575: p, 2p, 4p, ... normalised in a batch == one inversion each
    pt = oracle.ext_mul_by_cofactor(oracle.affine_to_extended(affine_raw(oracle, [K.TEST_POINT_RAW])))
    v = []
    for _ in range(10):
        v.append(pt[0].copy())
        pt = eng.point_double(pt)
    v = np.array(v)
    assert (eng.batch_normalize(v) == oracle.ext_to_affine(v)).all()
    # 8-torsion: small order, [8]T = identity, not torsion free except identity (src/lib.rs:1680-1754)
    tors = oracle.affine_to_extended(affine_raw(oracle, K.EIGHT_TORSION_RAW))
    assert eng.is_small_order(tors).all()
    c = eng.point_double(eng.point_double(eng.point_double(tors)))
    assert eng.is_identity(c).all()
    assert (eng.is_identity(p[:20]) == oracle.is_identity(p[:20])).all()
    assert (eng.is_torsion_free(p[:40]) == oracle.is_torsion_free(p[:40])).all()
    assert (eng.is_torsion_free(p[:40], ladder=True) == oracle.is_torsion_free(p[:40])).all()
    g8 = oracle.ext_mul_by_cofactor(p[9:30])
    assert eng.is_torsion_free(g8).all()
    assert (eng.is_small_order(p[:40]) == oracle.is_small_order(p[:40])).all()
    # is_prime_order (src/lib.rs:717-719): torsion free and not the identity; p[0] is the identity
    want_po = oracle.is_torsion_free(p[:40]) & (1 - oracle.is_identity(p[:40]))
    assert (eng.is_prime_order(p[:40]) == want_po).all() and eng.is_prime_order(p[:1])[0] == 0
    assert eng.is_prime_order(g8).all()
    # mul_by_cofactor (src/lib.rs:722-724): all 160 bytes
    assert (eng.mul_by_cofactor(q) == oracle.ext_mul_by_cofactor(q)).all()


def test_point_neg_eq_from_affine(eng, oracle):
    """Neg / PartialEq for ExtendedPoint and From<AffinePoint> (src/lib.rs:153-226) as batch kernels: negation and the
    affine -> extended map bit-exact on all 160 bytes against the oracle's field ops, equality against the oracle's
    normalised points -- equal points in different projective scalings, P vs -P, P vs P + T for the 8-torsion, and the
    points of order two and one, whose negation equals themselves."""
    n = 3000
    g = oracle.affine_to_extended(oracle.generator())
    t = oracle.fe_to_bytes(FR, oracle.fe_stream(FR, M.SEED0 + 41, n))
    p = oracle.scalar_mul(np.repeat(g, n, axis=0), t)  # z != 1
    tors = oracle.affine_to_extended(affine_raw(oracle, K.EIGHT_TORSION_RAW))
    p[:8] = tors
    # negation: (-U, V, Z, -T1, T2)
    want = p.copy()
    want[:, 0:4] = oracle.fe_batch(FQ, oracle.OP_NEG, np.ascontiguousarray(p[:, 0:4]))
    want[:, 12:16] = oracle.fe_batch(FQ, oracle.OP_NEG, np.ascontiguousarray(p[:, 12:16]))
    neg = eng.point_neg(p)
    assert (neg == want).all()
    assert (eng.point_neg(eng.to_device(p)).download() == want).all()
    # from affine
    aff = oracle.batch_normalize(p)
    assert (eng.affine_to_extended(aff) == oracle.affine_to_extended(aff)).all()
    # equality: q = p rescaled by a random lambda (same point), r = p + T_j (different unless T_j = O), -p
    lam = oracle.fe_stream(FQ, M.SEED0 + 42, n)
    q = p.copy()
    for c in range(3):
        q[:, 4 * c:4 * c + 4] = oracle.fe_batch(FQ, oracle.OP_MUL, np.ascontiguousarray(p[:, 4 * c:4 * c + 4]), lam)
    assert eng.point_eq(p, q).all() and eng.point_eq(q, p).all()
    r = eng.point_add(p, np.tile(tors, (n // 8, 1)))
    ar = oracle.batch_normalize(r)
    want_eq = (ar == aff).all(axis=1)
    got = eng.point_eq(p, r)
    assert (got == want_eq).all() and 0.1 * n < got.sum() < 0.15 * n  # T_j = O for one j in eight
    an = oracle.batch_normalize(neg)
    self_neg = (an == aff).all(axis=1)
    assert (eng.point_eq(p, neg) == self_neg).all() and 0 < self_neg.sum() <= 2  # only the identity and (0, -1)
    assert (eng.point_eq(eng.to_device(p), eng.to_device(r)).download().ravel() == want_eq).all()
    with pytest.raises(Exception):
        eng.point_eq(p, q[:-1])


def _oracle_sum(oracle, p):
    """Sum of the rows of p by halving with the oracle's `&ExtendedPoint + &ExtendedPoint` (any order gives the same point)."""
    p = p.copy()
    while len(p) > 1:
        if len(p) % 2:
            p = np.concatenate([p, oracle.identity()])
        h = len(p) // 2
        p = oracle.ext_add(np.ascontiguousarray(p[:h]), np.ascontiguousarray(p[h:]))
    return p


def test_point_sum(eng, oracle):
    """Sum<ExtendedPoint> (src/lib.rs:183-193) batched and grouped: whole-batch sums, group sums (sizes that are and are
    not multiples of the per-thread chain), every output format, host and device; and with jj_scalar_mul in front of it the
    multi-scalar product sum_i [k_i] P_i against the oracle."""
    p = oracle.ext_double(_points(oracle, 5040 - 9))            # 5040 points incl. identity and the 8-torsion, z != 1
    want = oracle.batch_normalize(_oracle_sum(oracle, p))
    got = eng.point_sum(p)
    assert got.shape == (1, 20) and (oracle.batch_normalize(got) == want).all()
    assert (eng.point_sum(p, output="affine") == want).all()
    assert (eng.point_sum(p, output="bytes") == oracle.affine_to_bytes(want)).all()
    assert (eng.point_sum(eng.to_device(p), output="affine").download() == want).all()
    for g in (1, 2, 5, 16, 40, 252, 1008):  # below, at and above the 16-point chain of one thread, with ragged tails
        sums = eng.point_sum(p, group_size=g, output="affine")
        assert sums.shape == (5040 // g, 8)
        for j in (0, 1, 5040 // g - 1):
            assert (sums[j] == oracle.batch_normalize(_oracle_sum(oracle, p[j * g:(j + 1) * g]))[0]).all(), (g, j)
    assert oracle.is_identity(eng.point_sum(np.zeros((0, 20), np.uint64))).all()   # the empty sum is the identity
    with pytest.raises(Exception):
        eng.point_sum(p, group_size=11)                                             # 5040 is not a multiple of 11
    # sum_i [k_i] P_i on the device: scalar-mul results stay in HBM, then one sum
    n = 1 << 16
    t = eng.fe_to_bytes("fr", eng.fe_stream("fr", 9001, n, device=True))
    k = eng.fe_to_bytes("fr", eng.fe_stream("fr", 9002, n, device=True))
    pts = eng.scalar_mul_fixed_vartime(oracle.generator(), t)
    msm = eng.point_sum(eng.scalar_mul_vartime(pts, k), output="affine").download()
    # [k_i][t_i]G summed = [sum k_i t_i mod 8r] G: checked through Fr arithmetic on the prime-order part instead of 65 536
    # oracle scalar-muls: multiply everything by the cofactor so that the scalars live in Fr
    kt = oracle.fe_batch(FR, oracle.OP_MUL, oracle.fe_stream(FR, 9001, n), oracle.fe_stream(FR, 9002, n))
    acc = kt.copy()
    while len(acc) > 1:
        h = len(acc) // 2
        acc = oracle.fe_batch(FR, oracle.OP_ADD, np.ascontiguousarray(acc[:h]), np.ascontiguousarray(acc[h:]))
    g8 = oracle.ext_mul_by_cofactor(oracle.affine_to_extended(oracle.generator()))
    want8 = oracle.batch_normalize(oracle.scalar_mul(g8, oracle.fe_to_bytes(FR, acc)))
    got8 = oracle.batch_normalize(oracle.ext_mul_by_cofactor(oracle.affine_to_extended(msm)))
    assert (got8 == want8).all()


def test_batch_normalize_extended_in_place(eng, oracle):
    """The free function batch_normalize (src/lib.rs:1084-1107): the ExtendedPoints themselves are normalised
    (z = 1, t1 = u, t2 = v), in place on the host and on the device; z = 0 ends as (0, 0, 1, 0, 0)."""
    p = oracle.ext_double(oracle.ext_double(_points(oracle, 5000)))
    p[7, 8:12] = 0
    aff = oracle.batch_normalize(p)
    one = np.repeat(oracle.fe_one(FQ), len(p), axis=0)
    want = np.concatenate([aff, one, aff], axis=1)
    assert (eng.batch_normalize_extended(p) == want).all()
    q = p.copy()
    out = eng.batch_normalize_extended(q, in_place=True)
    assert out is q and (q == want).all() and (q[7, :8] == 0).all() and (q[7, 12:] == 0).all()
    d = eng.to_device(p)
    assert eng.batch_normalize_extended(d, in_place=True) is d and (d.download() == want).all()
    import jubjub_b200.types as T

    pts = T.ExtendedPoint(p.copy(), eng)
    a = T.batch_normalize(pts)  # reference semantics: normalises `pts` and hands back the affine points
    assert (a.data == aff).all() and (pts.data == want).all()


def test_out_buffer_is_validated(eng, oracle):
    """a caller-supplied `out` of the wrong shape or dtype must raise instead of being overrun."""
    a, b = oracle.fe_stream(FQ, 1, 64), oracle.fe_stream(FQ, 2, 64)
    with pytest.raises(ValueError):
        eng.fe_mul("fq", a, b, out=np.empty((32, 4), np.uint64))
    with pytest.raises(ValueError):
        eng.fe_mul("fq", a, b, out=np.empty((64, 4), np.uint32))
    with pytest.raises(ValueError):
        eng.fe_mul("fq", eng.to_device(a), eng.to_device(b), out=eng.empty((64, 8)))
    big = np.zeros((100, 4), np.uint64)
    assert eng.fe_mul("fq", a, b, out=big) is big and (big[:64] == oracle.fe_batch(FQ, oracle.OP_MUL, a, b)).all()


def test_batch_from_bytes(eng, oracle):
    """src/lib.rs:541-627, KATs :1811-1876 and ZIP 216 :1894-1935."""
    p = _points(oracle, 900)
    enc = oracle.affine_to_bytes(oracle.batch_normalize(oracle.ext_double(p)))
    bad = enc[:200].copy()
    bad[:, 0] ^= 1
    extra = b32(*K.SERIALIZED_MULTIPLES_OF_8G, *K.ZIP216_NON_CANONICAL, [0xFF] * 32, [1] + [0] * 31, [0] * 32)
    allenc = np.concatenate([enc, bad, extra])
    want, wok = oracle.batch_from_bytes(allenc)
    got, ok = eng.batch_from_bytes(allenc)
    assert (ok == wok).all() and 0 < ok.sum() < len(ok)
    assert (got[ok == 1] == want[wok == 1]).all()
    assert (got[ok == 0] == 0).all()  # CtOption::none carries no value: the engine returns (0, 0)
    g8 = oracle.ext_mul_by_cofactor(oracle.affine_to_extended(oracle.generator()))
    cur = g8
    base = len(enc) + len(bad)
    for i in range(16):  # the reference test: batch-decoded k*(8G) equals the point itself
        assert ok[base + i] == 1 and (got[base + i] == oracle.ext_to_affine(cur)[0]).all()
        cur = oracle.ext_add(cur, g8)
    z = base + 16
    assert ok[z] == 0 and ok[z + 1] == 0  # ZIP 216: non-canonical (0, 1) and (0, -1) rejected ...
    got2, ok2 = eng.batch_from_bytes(allenc, zip216=False)
    want2, wok2 = oracle.affine_from_bytes(allenc, zip216=False)
    assert ok2[z] == 1 and ok2[z + 1] == 1 and (ok2 == wok2).all()  # ... and accepted by the pre-ZIP-216 API
    assert (got2[ok2 == 1] == want2[wok2 == 1]).all()
    # round trip on the device: decode(encode(P)) == P
    aff = eng.batch_normalize(p)
    back, okb = eng.batch_from_bytes(eng.affine_to_bytes(aff))
    assert okb.all() and (back == aff).all()


def test_batch_from_bytes_long_chains(eng, oracle):
    """More encodings than resident threads, so every thread runs Montgomery's trick over a chain of several
    encodings (the reference's single batched inversion, src/lib.rs:596-600), with rejected encodings of every
    kind (non-canonical v = skipped zero denominator, off-curve v, flipped sign) scattered through the chains."""
    n = 2 * 75776 + 999
    t = oracle.fe_to_bytes(FR, oracle.fe_stream(FR, M.SEED0 + 7, n))
    enc = eng.affine_to_bytes(eng.batch_normalize(eng.scalar_mul_fixed_vartime(oracle.generator(), t)))
    enc = enc.copy()
    enc[5::11, 0] ^= 1          # v + 1 (or - 1): mostly off the curve
    enc[3::97] = 0xFF           # non-canonical v
    enc[50::101, 31] ^= 0x80    # other sign: still valid, u negated
    enc[7::1013] = 0
    enc[7::1013, 0] = 1         # (0, 1), the identity
    got, ok = eng.batch_from_bytes(enc)
    want, wok = oracle.batch_from_bytes(enc)
    assert (ok == wok).all() and 0.8 * n < ok.sum() < n
    assert (got[ok == 1] == want[wok == 1]).all() and (got[ok == 0] == 0).all()


def test_scalar_mul_encoded(eng, oracle):
    """jj_scalar_mul_encoded: 32-byte encodings in, decode + scalar-mul on the device, every output format, host
    buffers (several staged chunks) and device-resident; rejected encodings are flagged like batch_from_bytes."""
    n = 150000
    t = oracle.fe_to_bytes(FR, oracle.fe_stream(FR, M.SEED0 + 17, n))
    k = oracle.fe_to_bytes(FR, oracle.fe_stream(FR, M.SEED0 + 18, n))
    aff = eng.batch_normalize(eng.scalar_mul_fixed_vartime(oracle.generator(), t))
    enc = eng.affine_to_bytes(aff).copy()
    enc[11::37, 0] ^= 1        # mostly off the curve
    enc[5::113] = 0xFF         # non-canonical
    want_aff, wok = eng.batch_from_bytes(enc)     # decode parity itself is covered by test_batch_from_bytes*
    ext = np.concatenate([want_aff, np.repeat(oracle.fe_one(FQ), n, axis=0), want_aff], axis=1)
    want = eng.scalar_mul_vartime(ext, k, output="bytes")
    got, ok = eng.scalar_mul_encoded_vartime(enc, k)
    assert (ok == wok).all() and 0.9 * n < ok.sum() < n
    assert (got[ok == 1] == want[ok == 1]).all()
    # against the oracle end to end on a sample: decode -> ladder -> normalise -> encode
    s = slice(0, 400)
    oa, ook = oracle.batch_from_bytes(enc[s])
    oext = np.concatenate([oa, np.repeat(oracle.fe_one(FQ), 400, axis=0), oa], axis=1)
    owant = oracle.affine_to_bytes(oracle.batch_normalize(oracle.scalar_mul(oext, k[s])))
    assert (ok[s] == ook).all() and (got[s][ook == 1] == owant[ook == 1]).all()
    for output in ("affine", "extended"):
        g2, ok2 = eng.scalar_mul_encoded_vartime(enc[:5000], k[:5000], output=output)
        a2 = g2 if output == "affine" else eng.batch_normalize(g2)
        assert (ok2 == wok[:5000]).all()
        assert (eng.affine_to_bytes(a2)[ok2 == 1] == want[:5000][ok2 == 1]).all()
    gd, okd = eng.scalar_mul_encoded_vartime(eng.to_device(enc[:70000]), eng.to_device(k[:70000]))
    assert (okd.download().ravel() == wok[:70000]).all()
    assert (gd.download()[wok[:70000] == 1] == want[:70000][wok[:70000] == 1]).all()


def test_scalar_mul_encoded_pinned_buffers(eng, oracle):
    """Pinned host buffers: the decode kernel reads the encodings in place, the scalars are uploaded behind the decode,
    and 32-byte results are stored in place (no staging copies) -- same bytes and flags as the staged path with
    pageable buffers, for every output format, with the subgroup check, across several chunks, and for a pinned
    buffer that is not 32-byte aligned (falls back to staging)."""
    import jubjub_b200 as jj
    from bench import pinned

    n = 1300000  # three chunks of eight rounds
    t = oracle.fe_to_bytes(FR, oracle.fe_stream(FR, M.SEED0 + 31, n))
    t[::3, 0] &= 0xF8
    k = oracle.fe_to_bytes(FR, oracle.fe_stream(FR, M.SEED0 + 32, n))
    enc = eng.affine_to_bytes(eng.batch_normalize(eng.scalar_mul_fixed_vartime(oracle.generator(), t))).copy()
    enc[7::53, 0] ^= 1
    enc[3::211] = 0xFF
    henc, p1 = pinned(eng, (n + 1, 32), np.uint8)
    hk, p2 = pinned(eng, (n, 32), np.uint8)
    hok, p3 = pinned(eng, (n,), np.uint8)
    hout, p4 = pinned(eng, (n, 160), np.uint8)
    try:
        henc[:n], hk[:] = enc, k
        for flags, unit in ((jj.JJ_OUT_BYTES, 32), (jj.JJ_OUT_BYTES | jj.JJ_CHECK_SUBGROUP, 32), (jj.JJ_OUT_AFFINE, 64), (0, 160)):
            m = n if unit == 32 else 200000
            want, wok = eng.scalar_mul_encoded_vartime(enc[:m], k[:m], output={32: "bytes", 64: "affine", 160: "extended"}[unit],
                                                       check_subgroup=bool(flags & jj.JJ_CHECK_SUBGROUP))
            hout[:] = 0xA5
            hok[:] = 7
            eng._check(eng.lib.jj_scalar_mul_encoded(eng.ctx, henc.ctypes.data, hk.ctypes.data, hout.ctypes.data, hok.ctypes.data, m, flags))
            got = hout.reshape(-1)[: m * unit].reshape(m, unit)
            assert (hok[:m] == wok).all() and 0 < wok.sum() < m
            assert (got[wok == 1] == want.view(np.uint8).reshape(m, unit)[wok == 1]).all()
            assert (hok[m:] == 7).all() and (hout.reshape(-1)[m * unit:] == 0xA5).all()  # nothing written past the batch
        # pinned but misaligned input (16 bytes in): staged like a pageable buffer, same results
        m = 100000
        mis = henc.reshape(-1)[16: 16 + m * 32].reshape(m, 32)
        mis[:] = enc[:m]
        want, wok = eng.scalar_mul_encoded_vartime(enc[:m], k[:m])
        eng._check(eng.lib.jj_scalar_mul_encoded(eng.ctx, mis.ctypes.data, hk.ctypes.data, hout.ctypes.data, hok.ctypes.data, m, jj.JJ_OUT_BYTES))
        assert (hok[:m] == wok).all() and (hout.reshape(-1)[: m * 32].reshape(m, 32)[wok == 1] == want[wok == 1]).all()
        # the same through the host mirror: Engine.pinned_empty() arrays as inputs, out= and ok_out=
        m = 200000
        pe, pk = eng.pinned_empty((m, 32), np.uint8), eng.pinned_empty((m, 32), np.uint8)
        po, pok = eng.pinned_empty((m, 32), np.uint8), eng.pinned_empty((m,), np.uint8)
        pe[:], pk[:] = enc[:m], k[:m]
        want, wok = eng.scalar_mul_encoded_vartime(enc[:m], k[:m])
        got, gok = eng.scalar_mul_encoded_vartime(pe, pk, out=po, ok_out=pok)
        assert got is po and (gok == wok).all() and (pok == wok).all() and (po[wok == 1] == want[wok == 1]).all()
        del pe, pk, po, pok, got, gok
    finally:
        for p in (p1, p2, p3, p4):
            eng.lib.jj_host_free(eng.ctx, p)


def test_scalar_mul_encoded_subgroup_check(eng, oracle):
    """JJ_CHECK_SUBGROUP: the decode is SubgroupPoint::from_bytes (src/lib.rs:1427-1429) -- ok[i] = decoded AND
    torsion free -- and accepted units are multiplied as before.  Host (staged chunks) and device resident."""
    n = 100000
    t = oracle.fe_to_bytes(FR, oracle.fe_stream(FR, M.SEED0 + 21, n))
    t[::3, 0] &= 0xF8  # a third of the points in the prime-order subgroup: [t]G with 8 | t
    k = oracle.fe_to_bytes(FR, oracle.fe_stream(FR, M.SEED0 + 22, n))
    pts = eng.scalar_mul_fixed_vartime(oracle.generator(), t)
    enc = eng.affine_to_bytes(eng.batch_normalize(pts)).copy()
    enc[13::41] = 0xFF  # some undecodable
    _, dec_ok = eng.batch_from_bytes(enc)
    want_ok = dec_ok & ((t[:, 0] & 7) == 0)
    got, ok = eng.scalar_mul_encoded_vartime(enc, k, check_subgroup=True)
    assert (ok == want_ok).all() and 0.35 * n < ok.sum() < 0.45 * n  # 1/3 forced + 1/8 of the rest, minus undecodable
    plain, _ = eng.scalar_mul_encoded_vartime(enc, k)
    assert (got[ok == 1] == plain[ok == 1]).all()
    s = np.flatnonzero(ok)[:300]
    want = oracle.affine_to_bytes(oracle.batch_normalize(oracle.scalar_mul(pts[s], k[s])))
    assert (got[s] == want).all()
    assert (oracle.is_torsion_free(pts[:400]) == ((t[:400, 0] & 7) == 0)).all()  # the property used above, vs the oracle
    gd, okd = eng.scalar_mul_encoded_vartime(eng.to_device(enc), eng.to_device(k), check_subgroup=True)
    assert (okd.download().ravel() == want_ok).all() and (gd.download()[ok == 1] == got[ok == 1]).all()


def test_is_torsion_free_full_size_all_cosets(eng, oracle):
    """2^20 points (BASELINE size): P_i = [t_i] G with G of order 8r (src/lib.rs:1380-1396) is torsion free exactly
    when 8 | t_i (src/lib.rs:709-711).  The pairing test (default), the reference-style [r]P == O kernel
    (JJ_TORSION_LADDER) and that property must agree on every unit; then every coset P + T_j of prime-order points
    over the reference's eight torsion points (src/lib.rs:1589-1677), in projective forms with z != 1, against the
    oracle's bitwise ladder."""
    n = 1 << 20
    t = oracle.fe_to_bytes(FR, oracle.fe_stream(FR, M.SEED0 + 13, n))
    t[::5, 0] &= 0xF8  # make a fifth of the scalars multiples of 8
    pts = eng.scalar_mul_fixed_vartime(oracle.generator(), t)
    flags = eng.is_torsion_free(pts)
    assert (flags == ((t[:, 0] & 7) == 0)).all() and 0.2 * n < flags.sum() < 0.4 * n
    assert (eng.is_torsion_free(pts, ladder=True) == flags).all()
    p8 = eng.mul_by_cofactor(pts)
    assert eng.is_torsion_free(p8).all() and eng.is_prime_order(p8).sum() >= n - 1
    assert (flags[:600] == oracle.is_torsion_free(pts[:600])).all()
    # all eight cosets: P + T_j is torsion free only for T_j = O
    m = 1 << 15
    tors = oracle.affine_to_extended(affine_raw(oracle, K.EIGHT_TORSION_RAW))
    tors_is_o = oracle.is_identity(tors)
    assert tors_is_o.sum() == 1
    base = eng.point_double(p8[:m])                      # prime order, z != 1
    for j in range(8):
        c = eng.point_add(base, np.repeat(tors[j:j + 1], m, axis=0))
        f = eng.is_torsion_free(c)
        assert (f == tors_is_o[j]).all(), j
        assert (f[:64] == oracle.is_torsion_free(c[:64])).all(), j
        assert (eng.is_torsion_free(c[:4096], ladder=True) == tors_is_o[j]).all(), j
    small = np.concatenate([tors, eng.point_double(tors), oracle.identity(4)])
    assert (eng.is_torsion_free(small) == oracle.is_torsion_free(small)).all()


def test_find_eight_torsion_on_gpu(eng, oracle):
    """src/lib.rs:1680-1696: [r] G walks the 8-torsion subgroup."""
    g = oracle.affine_to_extended(affine_raw(oracle, [K.FULL_GENERATOR_RAW]))
    t = eng.scalar_mul_vartime(g, b32(K.FR_MODULUS_BYTES))
    want = affine_raw(oracle, K.EIGHT_TORSION_RAW)
    cur = t
    for i in range(8):
        assert (eng.batch_normalize(cur) == want[i]).all(), i
        cur = eng.point_add(cur, t)


# ------------------------------------------------------------------------------ residency, sizes, errors
def test_device_resident_and_in_place(eng, oracle):
    a, b = _field_inputs(oracle, FQ, 4096)
    da, db = eng.to_device(a), eng.to_device(b)
    out = eng.fe_mul("fq", da, db, out=da)  # in place, like MulAssign (src/util.rs:126-152)
    assert out is da and (da.download() == oracle.fe_batch(FQ, oracle.OP_MUL, a, b)).all()
    p, k = _smul_case(oracle, 500)
    dp, dk = eng.to_device(p), eng.to_device(k)
    got = eng.scalar_mul_vartime(dp, dk, output="affine").download()
    assert (got == oracle.batch_normalize(oracle.scalar_mul(p, k))).all()
    dd = eng.point_double(dp, out=dp)
    assert (dd.download() == oracle.ext_double(p)).all()


def test_empty_ragged_and_chunk_boundaries(eng, oracle):
    assert eng.fe_mul("fq", np.zeros((0, 4), np.uint64), np.zeros((0, 4), np.uint64)).shape == (0, 4)
    assert eng.scalar_mul_vartime(np.zeros((0, 20), np.uint64), np.zeros((0, 32), np.uint8)).shape == (0, 20)
    for n in (1, 31, 33, (1 << 17) - 1, (1 << 17) + 1, 3 * (1 << 17) + 5):  # staging chunk is 2^17 units
        a, b = oracle.fe_stream(FQ, 1, n), oracle.fe_stream(FQ, 2, n)
        assert (eng.fe_mul("fq", a, b) == oracle.fe_batch(FQ, oracle.OP_MUL, a, b)).all(), n
    p, k = _smul_case(oracle, 70)
    for n in (1, 31, 33, 79):
        assert (eng.scalar_mul_vartime(p[:n], k[:n], output="affine") == oracle.batch_normalize(oracle.scalar_mul(p[:n], k[:n]))).all()


def test_error_behaviour(eng, oracle):
    import ctypes as C

    import jubjub_b200 as jj

    a = oracle.fe_stream(FQ, 1, 8)
    with pytest.raises(jj.JubjubError) as e:  # the reference panics on length mismatch (src/lib.rs:841)
        eng.fe_mul("fq", a, a[:4])
    assert e.value.code == -1
    assert eng.lib.jj_fq_mul(eng.ctx, None, a.ctypes.data, a.ctypes.data, 8, 0) == -1
    d = eng.to_device(a)
    assert eng.lib.jj_fq_mul(eng.ctx, d.ptr + 8, d.ptr, d.ptr, 4, jj.JJ_DEVICE_PTRS) == -1  # misaligned
    assert b"aligned" in eng.lib.jj_last_error(eng.ctx)
    assert eng.lib.jj_set_scalar_mul_variant(eng.ctx, 99) == -1
    assert eng.lib.jj_set_scalar_mul_variant(eng.ctx, 15) == -1  # experimental mappings are not in the default build
    # new entry points validate their arguments the same way
    p, k = _smul_case(oracle, 40)
    enc = oracle.affine_to_bytes(oracle.batch_normalize(p))
    out = np.zeros((len(p), 32), np.uint8)
    rc = eng.lib.jj_scalar_mul_encoded(eng.ctx, enc.ctypes.data, k.ctypes.data, out.ctypes.data, None, len(p),
                                       jj.JJ_OUT_BYTES | jj.JJ_CHECK_SUBGROUP)
    assert rc == -1 and b"ok[]" in eng.lib.jj_last_error(eng.ctx)        # the subgroup flag needs somewhere to go
    rc = eng.lib.jj_scalar_mul_encoded(eng.ctx, enc.ctypes.data, k.ctypes.data, out.ctypes.data, None, len(p), jj.JJ_OUT_BYTES)
    assert rc == 0 and (out == oracle.affine_to_bytes(oracle.batch_normalize(oracle.scalar_mul(p, k)))).all()  # ok[] is optional
    eng.set_scalar_mul_variant(24)
    try:
        with pytest.raises(jj.JubjubError):  # the constant-time mode exists for the default mapping only
            eng.scalar_mul(p, k)
    finally:
        eng.set_scalar_mul_variant(0)
    assert eng.lib.jj_mul_by_cofactor(eng.ctx, None, out.ctypes.data, 4, 0) == -1
    assert eng.lib.jj_is_prime_order(eng.ctx, p.ctypes.data, None, 4, 0) == -1
    assert eng.lib.jj_scalar_mul_sharded_n(eng.ctx, p.ctypes.data, k.ctypes.data, None, None, 4, 0) == -1
    # single rank: the sharded entry points degenerate to the plain call (no communicator needed)
    d_out = eng.empty((len(p), 20))
    eng.scalar_mul_sharded_vartime(p, k, d_out, n_total=len(p))
    assert (eng.batch_normalize(d_out.download()) == oracle.batch_normalize(oracle.scalar_mul(p, k))).all()
    ctx = C.c_void_p()
    assert eng.lib.jj_init(10_000, C.byref(ctx)) == -1


def test_full_size_linearity_1m(eng, oracle):
    """BASELINE config 3 size (2^20 units): [a]P + [b]P == [a + b]P on the prime-order subgroup,
    every unit checked on the device path; a 512-unit sample is re-checked against the oracle."""
    n = 1 << 20
    g8 = oracle.ext_to_affine(oracle.ext_mul_by_cofactor(oracle.affine_to_extended(oracle.generator())))
    t = eng.fe_to_bytes("fr", eng.fe_stream("fr", M.SEED0 + 3, n, device=True))
    pts = eng.scalar_mul_fixed_vartime(g8, t)  # P_i = [t_i] 8G, device resident
    a = eng.fe_stream("fr", M.SEED0 + 2, n, device=True)
    b = eng.fe_stream("fr", M.SEED0 + 4, n, device=True)
    ab = eng.fe_add("fr", a, b)
    pa = eng.scalar_mul_vartime(pts, a, scalar_mont=True)
    pb = eng.scalar_mul_vartime(pts, b, scalar_mont=True)
    lhs = eng.batch_normalize(eng.point_add(pa, pb)).download()
    rhs = eng.scalar_mul_vartime(pts, ab, scalar_mont=True, output="affine").download()
    assert (lhs == rhs).all()
    idx = np.arange(0, n, n // 512)[:512]
    ph, ah = pts.download()[idx], oracle.fe_to_bytes(FR, a.download()[idx])
    assert (eng.batch_normalize(pa).download()[idx] == oracle.batch_normalize(oracle.scalar_mul(ph, ah))).all()


def test_full_size_fixed_base_1m(eng, oracle):
    """BASELINE config 4 size (2^20 units): [k]G through the shared per-window table equals [k]G through the
    variable-base kernel for every unit (two independent kernels), plus an oracle-checked sample."""
    n = 1 << 20
    gen = oracle.generator()
    k = eng.fe_to_bytes("fr", eng.fe_stream("fr", M.SEED0 + 2, n, device=True))
    fixed = eng.scalar_mul_fixed_vartime(gen, k, output="affine").download()
    pts = eng.to_device(np.repeat(oracle.affine_to_extended(gen), n, axis=0))
    var = eng.scalar_mul_vartime(pts, k, output="affine").download()
    assert (fixed == var).all()
    idx = np.arange(0, n, n // 256)[:256]
    kh = k.download()[idx]
    assert (fixed[idx] == oracle.batch_normalize(oracle.scalar_mul_fixed(gen, kh))).all()


def test_scalar_mul_differential_64k(eng, oracle):
    """65 536 variable-base scalar-muls (projective inputs with z != 1, full-range 256-bit scalar strings so
    that bits 252..255 are exercised) compared unit by unit with the oracle's reference ladder."""
    n = 1 << 16
    t = eng.fe_to_bytes("fr", eng.fe_stream("fr", 7001, n))
    p = eng.point_double(eng.scalar_mul_fixed_vartime(oracle.generator(), t))  # z != 1
    rng = np.random.default_rng(2024)
    k = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)  # arbitrary top bits: ignored like multiply_bits
    got = eng.scalar_mul_vartime(p, k, output="bytes")
    want = oracle.affine_to_bytes(oracle.batch_normalize(oracle.scalar_mul(p, k)))
    assert (got == want).all()


def test_cuda_graph_replay(eng, oracle):
    """A captured chain of small field batches (x <- (x*y + y)^2, 16 times) replays bit-exactly."""
    import jubjub_b200 as jj

    n = 4096
    x0, y0 = oracle.fe_stream(FQ, 31, n), oracle.fe_stream(FQ, 32, n)
    x, y, t = eng.to_device(x0), eng.to_device(y0), eng.empty((n, 4))
    A = jj.JJ_ASYNC

    def chain():
        for _ in range(16):
            eng.fe_mul("fq", x, y, out=t, flags=A)
            eng.fe_add("fq", t, y, out=t, flags=A)
            eng.fe_square("fq", t, out=x, flags=A)

    want = x0
    for _ in range(16):
        want = oracle.fe_batch(FQ, oracle.OP_SQUARE, oracle.fe_batch(FQ, oracle.OP_ADD, oracle.fe_batch(FQ, oracle.OP_MUL, want, y0), y0))
    chain()
    eng.sync()
    assert (x.download() == want).all()
    g = eng.graph_capture(chain)
    for _ in range(3):
        x.upload(x0)
        eng.graph_launch(g)
        eng.sync()
        assert (x.download() == want).all()
    eng.graph_destroy(g)


def test_cuda_graph_scratch_safety(oracle):
    """a replayed graph must never write through a freed scratch pointer.  While a captured graph is alive
    a call that would have to grow a main-stream scratch buffer fails (and the graph still replays correctly); calls
    that cannot be captured (host pointers, allocation inside the capture) are refused and leave the capture usable."""
    import jubjub_b200 as jj

    eng = jj.Engine(0)  # own context: scratch sizes start from zero
    try:
        A = jj.JJ_ASYNC
        n1, n2 = 3000, 400000
        a = oracle.fe_stream(FQ, 41, n2)
        a[5] = 0
        d1, o1, ok1 = eng.to_device(a[:n1]), eng.empty((n1, 4)), eng.empty((n1, 1), np.uint8)
        d2, o2, ok2 = eng.to_device(a), eng.empty((n2, 4)), eng.empty((n2, 1), np.uint8)
        inv = lambda d, o, ok, n, f: eng.lib.jj_fq_invert(eng.ctx, d.ptr, o.ptr, ok.ptr, n, jj.JJ_DEVICE_PTRS | f)  # noqa: E731
        want1, _ = oracle.fe_invert(FQ, a[:n1])
        # capturing before the scratch exists: refused, nothing allocated inside the capture
        eng._check(eng.lib.jj_graph_begin(eng.ctx))
        assert inv(d1, o1, ok1, n1, A) == -1 and b"eagerly" in eng.lib.jj_last_error(eng.ctx)
        assert eng.lib.jj_fq_mul(eng.ctx, a.ctypes.data, a.ctypes.data, a.ctypes.data, 8, 0) == -1  # host pointers
        import ctypes as C

        g0 = C.c_void_p()
        assert eng.lib.jj_graph_end(eng.ctx, C.byref(g0)) == 0  # the (empty) capture ends cleanly
        if g0.value:
            eng.graph_destroy(g0)
        assert inv(d1, o1, ok1, n1, 0) == 0  # eager run creates the scratch (n1 x 32 B)
        g = eng.graph_capture(lambda: eng._check(inv(d1, o1, ok1, n1, A)))
        # a larger batch would reallocate the scratch the graph's kernel node points at: refused while the graph lives
        assert inv(d2, o2, ok2, n2, 0) == -1 and b"graph" in eng.lib.jj_last_error(eng.ctx)
        o1.upload(np.zeros((n1, 4), np.uint64))
        eng.graph_launch(g)
        eng.sync()
        assert (o1.download() == want1).all()
        eng.graph_destroy(g)
        assert inv(d2, o2, ok2, n2, 0) == 0  # allowed again once the graph is gone
        assert (o2.download() == oracle.fe_invert(FQ, a)[0]).all()
    finally:
        eng.close()

