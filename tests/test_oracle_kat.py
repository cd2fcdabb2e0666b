"""Pins the CPU oracle (oracle/jj_oracle.c) and the bigint model (oracle/model.py)
against the reference's own known-answer tests: src/fr.rs:787-1244 and
src/lib.rs:1456-1935.  Test names follow the reference's test names."""
import numpy as np

from oracle import model as M
from tests.golden import reference_kats as K
from tests.helpers import affine_raw, affine_values, b32, fe, fe_int, scalar_bytes, to_int

FQ, FR = 0, 1


# ----------------------------------------------------------------- src/fr.rs
def test_constants(oracle):
    assert int(K.FR_MODULUS_HEX, 16) == M.from_limbs(K.FR_MODULUS) == M.R_ORDER
    assert to_int(oracle.fe_one(FR)[0]) == M.from_limbs(K.FR_R) == M.MONT_R % M.R_ORDER
    assert M.from_limbs(K.FR_R2) == pow(2, 512, M.R_ORDER)
    assert M.from_limbs(K.FR_R3) == pow(2, 768, M.R_ORDER)
    two = oracle.fe_from_raw(FR, fe([2, 0, 0, 0]))
    assert (oracle.fe_batch(FR, oracle.OP_MUL, two, fe(K.FR_TWO_INV)) == oracle.fe_one(FR)).all()
    assert M.from_limbs(K.FQ_MODULUS) == M.Q
    # q - 1 as written at src/lib.rs:1629-1634 (EIGHT_TORSION[3].v)
    assert M.from_limbs(K.EIGHT_TORSION_RAW[3][1]) == M.Q - 1


def test_inv():
    for m, want in ((M.R_ORDER, K.FR_INV), (M.Q, 0xFFFFFFFEFFFFFFFF)):
        assert (-pow(m, -1, 1 << 64)) % (1 << 64) == want


def test_debug(oracle):
    def dbg(x):
        return "0x" + bytes(oracle.fe_to_bytes(FR, x)[0])[::-1].hex()

    assert dbg(fe([0, 0, 0, 0])) == K.FR_DEBUG["zero"]
    assert dbg(oracle.fe_one(FR)) == K.FR_DEBUG["one"]
    assert dbg(fe(K.FR_R2)) == K.FR_DEBUG["R2"]


def test_to_bytes(oracle):
    neg_one = oracle.fe_batch(FR, oracle.OP_NEG, oracle.fe_one(FR))
    x = np.concatenate([fe([0, 0, 0, 0]), oracle.fe_one(FR), fe(K.FR_R2), neg_one])
    got = oracle.fe_to_bytes(FR, x)
    want = b32([0] * 32, [1] + [0] * 31, K.FR_BYTES_R2, K.FR_BYTES_NEG_ONE)
    assert (got == want).all()


def test_from_bytes(oracle):
    x, ok = oracle.fe_from_bytes(FR, b32([0] * 32, [1] + [0] * 31, K.FR_BYTES_R2, K.FR_BYTES_NEG_ONE))
    assert ok.tolist() == [1, 1, 1, 1]
    assert (x[0] == 0).all() and (x[1] == oracle.fe_one(FR)[0]).all() and (x[2] == fe(K.FR_R2)[0]).all()
    _, ok = oracle.fe_from_bytes(FR, b32(*K.FR_BYTES_REJECTED))
    assert ok.tolist() == [0, 0, 0, 0]


def _u512(limbs8):
    return np.frombuffer(b"".join(int(v).to_bytes(8, "little") for v in limbs8), dtype=np.uint8).reshape(1, 64)


def test_from_u512(oracle):
    mx = 0xFFFFFFFFFFFFFFFF
    assert (oracle.fe_from_bytes_wide(FR, _u512(K.FR_MODULUS + [0] * 4)) == 0).all()
    assert (oracle.fe_from_bytes_wide(FR, _u512([1] + [0] * 7)) == fe(K.FR_R)).all()
    assert (oracle.fe_from_bytes_wide(FR, _u512([0] * 4 + [1] + [0] * 3)) == fe(K.FR_R2)).all()
    r3_minus_r = oracle.fe_batch(FR, oracle.OP_SUB, fe(K.FR_R3), fe(K.FR_R))
    assert (oracle.fe_from_bytes_wide(FR, _u512([mx] * 8)) == r3_minus_r).all()


def test_from_bytes_wide(oracle):
    wide = np.zeros((2, 64), dtype=np.uint8)
    wide[0, :32] = K.FR_BYTES_R2
    wide[1, :32] = K.FR_BYTES_NEG_ONE
    got = oracle.fe_from_bytes_wide(FR, wide)
    assert (got[0] == fe(K.FR_R2)[0]).all()
    assert (got[1] == oracle.fe_batch(FR, oracle.OP_NEG, oracle.fe_one(FR))[0]).all()
    got = oracle.fe_from_bytes_wide(FR, np.full((1, 64), 0xFF, dtype=np.uint8))
    assert (got == fe(K.FR_WIDE_MAX_MONT)).all()
    # bigint model agrees on the raw Montgomery limbs
    assert M.to_mont(M.from_bytes_wide(b"\xff" * 64, M.R_ORDER), M.R_ORDER) == M.from_limbs(K.FR_WIDE_MAX_MONT)


def test_zero(oracle):
    z = fe([0, 0, 0, 0])
    for op in (oracle.OP_NEG, oracle.OP_ADD, oracle.OP_SUB, oracle.OP_MUL):
        assert (oracle.fe_batch(FR, op, z, z) == 0).all()


def test_addition(oracle):
    big = fe(K.FR_LARGEST)
    assert (oracle.fe_batch(FR, oracle.OP_ADD, big, big) == fe(K.FR_LARGEST_PLUS_LARGEST)).all()
    assert (oracle.fe_batch(FR, oracle.OP_ADD, big, fe([1, 0, 0, 0])) == 0).all()


def test_negation(oracle):
    assert (oracle.fe_batch(FR, oracle.OP_NEG, fe(K.FR_LARGEST)) == fe([1, 0, 0, 0])).all()
    assert (oracle.fe_batch(FR, oracle.OP_NEG, fe([0, 0, 0, 0])) == 0).all()
    assert (oracle.fe_batch(FR, oracle.OP_NEG, fe([1, 0, 0, 0])) == fe(K.FR_LARGEST)).all()


def test_subtraction(oracle):
    big = fe(K.FR_LARGEST)
    assert (oracle.fe_batch(FR, oracle.OP_SUB, big, big) == 0).all()
    a = oracle.fe_batch(FR, oracle.OP_SUB, fe([0, 0, 0, 0]), big)
    b = oracle.fe_batch(FR, oracle.OP_SUB, fe(K.FR_MODULUS), big)
    assert (a == b).all()


def _mul_by_double_and_add(oracle, which, cur):
    """The reference's in-test multiplication check: cur * cur by 256 double/add steps."""
    bits = bytes(oracle.fe_to_bytes(which, cur)[0])
    acc = fe([0, 0, 0, 0])
    for byte in reversed(bits):
        for i in reversed(range(8)):
            acc = oracle.fe_batch(which, oracle.OP_ADD, acc, acc)
            if (byte >> i) & 1:
                acc = oracle.fe_batch(which, oracle.OP_ADD, acc, cur)
    return acc


def test_multiplication_and_squaring(oracle):
    for which, largest in ((FR, fe(K.FR_LARGEST)), (FQ, fe_int(M.Q - 1))):
        cur = largest.copy()
        for _ in range(100):
            want = _mul_by_double_and_add(oracle, which, cur)
            assert (oracle.fe_batch(which, oracle.OP_MUL, cur, cur) == want).all()
            assert (oracle.fe_batch(which, oracle.OP_SQUARE, cur) == want).all()
            cur = oracle.fe_batch(which, oracle.OP_ADD, cur, largest)


def test_inversion(oracle):
    for which, r2 in ((FR, fe(K.FR_R2)), (FQ, fe_int(pow(2, 512, M.Q)))):
        one = oracle.fe_one(which)
        _, ok = oracle.fe_invert(which, fe([0, 0, 0, 0]))
        assert ok[0] == 0
        inv, ok = oracle.fe_invert(which, one)
        assert ok[0] == 1 and (inv == one).all()
        neg_one = oracle.fe_batch(which, oracle.OP_NEG, one)
        assert (oracle.fe_invert(which, neg_one)[0] == neg_one).all()
        tmp = r2.copy()
        for _ in range(100):
            inv, _ = oracle.fe_invert(which, tmp)
            assert (oracle.fe_batch(which, oracle.OP_MUL, inv, tmp) == one).all()
            tmp = oracle.fe_batch(which, oracle.OP_ADD, tmp, r2)


def test_invert_is_pow(oracle):
    r1 = fe(K.FR_R)
    for _ in range(100):
        a = oracle.fe_invert(FR, r1)[0]
        b = oracle.fe_pow_vartime(FR, r1, K.FR_R_MINUS_2)
        assert (a == b).all()
        r1 = oracle.fe_batch(FR, oracle.OP_ADD, a, fe(K.FR_R))


def test_sqrt(oracle):
    square = fe(K.FR_R_MINUS_2)
    none = 0
    for _ in range(100):
        root, ok = oracle.fe_sqrt(FR, square)
        if not ok[0]:
            none += 1
        else:
            assert (oracle.fe_batch(FR, oracle.OP_MUL, root, root) == square).all()
        square = oracle.fe_batch(FR, oracle.OP_SUB, square, oracle.fe_one(FR))
    assert none == K.FR_SQRT_NONE_COUNT


def test_fq_sqrt(oracle):
    a = oracle.fe_stream(FQ, 77, 64)
    root, ok = oracle.fe_sqrt(FQ, a)
    vals = M.stream_field(77, 64, M.Q)
    for i in range(64):
        assert bool(ok[i]) == (pow(vals[i], (M.Q - 1) // 2, M.Q) == 1)
        if ok[i]:
            assert (oracle.fe_batch(FQ, oracle.OP_SQUARE, root[i:i + 1]) == a[i]).all()


def test_from_raw(oracle):
    assert (oracle.fe_from_raw(FR, fe(K.FR_FROM_RAW_ALL_ONES_EQ)) == oracle.fe_from_raw(FR, fe([0xFFFFFFFFFFFFFFFF] * 4))).all()
    assert (oracle.fe_from_raw(FR, fe(K.FR_MODULUS)) == 0).all()
    assert (oracle.fe_from_raw(FR, fe([1, 0, 0, 0])) == fe(K.FR_R)).all()


# ----------------------------------------------------------------- src/lib.rs
def _ext(oracle, pts_raw):
    return oracle.affine_to_extended(affine_raw(oracle, pts_raw))


def _fr_mont_to_scalar_bytes(oracle, mont_limbs):
    return oracle.fe_to_bytes(FR, fe(mont_limbs))


def test_edwards_d(oracle):
    assert M.from_limbs(K.EDWARDS_D_RAW) == M.D
    assert M.from_limbs(K.EDWARDS_D2_RAW) == M.D2
    # test_d_is_non_quadratic_residue :1462-1466
    d = oracle.fe_from_raw(FQ, fe(K.EDWARDS_D_RAW))
    negd = oracle.fe_batch(FQ, oracle.OP_NEG, d)
    for x in (d, negd, oracle.fe_invert(FQ, negd)[0]):
        assert oracle.fe_sqrt(FQ, x)[1][0] == 0


def test_is_on_curve_var(oracle):
    ident = oracle.ext_to_affine(oracle.identity())
    assert oracle.is_on_curve(ident)[0] == 1
    assert oracle.is_on_curve(oracle.generator())[0] == 1
    assert M.on_curve((M.GEN_U, M.GEN_V))
    assert affine_values(oracle, oracle.generator()) == [(M.GEN_U, M.GEN_V)]


def test_niels_point_identity(oracle):
    one = oracle.fe_one(FQ)[0]
    n = oracle.ext_to_niels(oracle.identity())[0]
    assert (n[0:4] == one).all() and (n[4:8] == one).all() and (n[8:12] == one).all() and (n[12:16] == 0).all()
    a = oracle.affine_to_niels(oracle.ext_to_affine(oracle.identity()))[0]
    assert (a[0:4] == one).all() and (a[4:8] == one).all() and (a[8:12] == 0).all()


def test_assoc(oracle):
    p = oracle.ext_mul_by_cofactor(_ext(oracle, [K.TEST_POINT_RAW]))
    assert oracle.is_on_curve(oracle.ext_to_affine(p))[0]
    lhs = oracle.scalar_mul(oracle.scalar_mul(p, scalar_bytes(1000)), scalar_bytes(3938))
    rhs = oracle.scalar_mul(p, scalar_bytes(1000 * 3938))
    assert oracle.ext_eq(lhs, rhs)[0]
    # bigint model, independently
    pv = affine_values(oracle, oracle.ext_to_affine(p))[0]
    assert affine_values(oracle, oracle.ext_to_affine(rhs))[0] == M.pmul(pv, 3938000)


def test_batch_normalize(oracle):
    p = oracle.ext_mul_by_cofactor(_ext(oracle, [K.TEST_POINT_RAW]))
    v = []
    for _ in range(10):
        v.append(p[0].copy())
        p = oracle.ext_double(p)
    v = np.array(v)
    expected = oracle.ext_to_affine(v)
    assert oracle.is_on_curve(expected).all()
    assert (oracle.batch_normalize(v) == expected).all()


def test_find_eight_torsion(oracle):
    g = _ext(oracle, [K.FULL_GENERATOR_RAW])
    assert not oracle.is_small_order(g)[0]
    g = oracle.scalar_mul(g, b32(K.FR_MODULUS_BYTES))
    assert oracle.is_small_order(g)[0]
    want = affine_raw(oracle, K.EIGHT_TORSION_RAW)
    cur = g
    for i in range(8):
        assert (oracle.ext_to_affine(cur) == want[i]).all(), i
        cur = oracle.ext_add(cur, g)


def test_find_curve_generator(oracle):
    trial = np.zeros((1, 32), dtype=np.uint8)
    for _ in range(255):
        a, ok = oracle.affine_from_bytes(trial)
        if ok[0]:
            assert oracle.is_on_curve(a)[0]
            b = oracle.scalar_mul(oracle.affine_to_extended(a), b32(K.FR_MODULUS_BYTES))
            assert oracle.is_small_order(b)[0]
            b = oracle.ext_double(oracle.ext_double(b))
            if not oracle.is_identity(b)[0]:
                b = oracle.ext_double(b)
                assert oracle.is_identity(b)[0]
                assert (a == affine_raw(oracle, [K.FULL_GENERATOR_RAW])).all()
                assert (a == oracle.generator()).all()
                assert oracle.is_torsion_free(oracle.ext_mul_by_cofactor(oracle.affine_to_extended(a)))[0]
                return
        trial[0, 0] += 1
    raise AssertionError("should have found a generator of the curve")


def test_small_order_and_is_identity(oracle):
    pts = _ext(oracle, K.EIGHT_TORSION_RAW)
    assert oracle.is_small_order(pts).all()
    c = oracle.ext_mul_by_cofactor(pts)
    assert oracle.is_identity(c).all()
    a, b = c[0], c[1]
    assert (a[0:4] == b[0:4]).all() and (a[4:8] == a[8:12]).all() and (b[4:8] == b[8:12]).all()
    assert (a[4:8] != b[4:8]).any() and (a[8:12] != b[8:12]).any()


def test_mul_consistency(oracle):
    a, b, c = fe(K.MULC_A), fe(K.MULC_B), fe(K.MULC_C)
    assert (oracle.fe_batch(FR, oracle.OP_MUL, a, b) == c).all()
    assert M.mont_mul(M.from_limbs(K.MULC_A), M.from_limbs(K.MULC_B), M.R_ORDER) == M.from_limbs(K.MULC_C)
    sa, sb, sc = (oracle.fe_to_bytes(FR, x) for x in (a, b, c))
    p = oracle.ext_mul_by_cofactor(_ext(oracle, [K.TEST_POINT_RAW]))
    pc = oracle.scalar_mul(p, sc)
    pab = oracle.scalar_mul(oracle.scalar_mul(p, sa), sb)
    assert oracle.ext_eq(pc, pab)[0]
    # AffineNielsPoint path (src/lib.rs:1798-1803)
    pan = oracle.affine_to_niels(oracle.ext_to_affine(p))
    assert oracle.ext_eq(oracle.affine_niels_mul(pan, sc), pab)[0]
    assert oracle.ext_eq(oracle.scalar_mul(oracle.affine_niels_mul(pan, sa), sb), pc)[0]
    assert oracle.ext_eq(oracle.scalar_mul_fixed(oracle.ext_to_affine(p), sc), pc)[0]


def test_serialization_consistency(oracle):
    gen = oracle.ext_mul_by_cofactor(_ext(oracle, [K.FULL_GENERATOR_RAW]))
    want = b32(*K.SERIALIZED_MULTIPLES_OF_8G)
    batched, ok = oracle.batch_from_bytes(want)
    assert ok.all()
    p = gen
    g8 = M.pmul((M.GEN_U, M.GEN_V), 8)
    for i in range(16):
        affine = oracle.ext_to_affine(p)
        assert oracle.is_on_curve(affine)[0]
        ser = oracle.affine_to_bytes(affine)
        assert (ser[0] == want[i]).all(), i
        de, ok1 = oracle.affine_from_bytes(ser)
        assert ok1[0] and (de == affine).all() and (batched[i] == affine[0]).all()
        # independent bigint arithmetic reproduces the same encoding
        assert M.encode(M.pmul(g8, i + 1)) == bytes(want[i])
        assert M.decode(bytes(want[i])) == M.pmul(g8, i + 1)
        p = oracle.ext_add(p, gen)


def test_zip_216(oracle):
    for enc in K.ZIP216_NON_CANONICAL:
        e = b32(enc)
        assert oracle.affine_from_bytes(e)[1][0] == 0
        assert oracle.batch_from_bytes(e)[1][0] == 0
        assert M.decode(bytes(enc)) is None
        cleared = e.copy()
        cleared[0, 31] &= 0x7F
        assert oracle.affine_from_bytes(cleared)[1][0] == 1
        parsed, ok = oracle.affine_from_bytes(e, zip216=False)
        assert ok[0] == 1
        re = oracle.affine_to_bytes(parsed)
        assert (re != e).any()
        re[0, 31] |= 0x80
        assert (re == e).all()


def test_r_times_8g_is_identity(oracle):
    gen = oracle.ext_mul_by_cofactor(_ext(oracle, [K.FULL_GENERATOR_RAW]))
    assert oracle.is_torsion_free(gen)[0] == 1
    assert oracle.is_torsion_free(_ext(oracle, [K.FULL_GENERATOR_RAW]))[0] == 0


def test_config1_10k_fq_muls_vs_bigint(oracle):
    """BASELINE.json configs[0]: 10k Fq Montgomery muls on the CPU path (SplitMix64 streams of SURVEY 8d plus the
    benches/fq_bench.rs:6-7,25-33 values 1, -1, 4), Montgomery limbs bit-exact against the big-integer model."""
    n = 10_000
    a, b = oracle.fe_stream(FQ, M.SEED0, n), oracle.fe_stream(FQ, M.SEED0 + 1, n)
    am, bm = M.stream_field(M.SEED0, n, M.Q), M.stream_field(M.SEED0 + 1, n, M.Q)
    c = oracle.fe_batch(FQ, oracle.OP_MUL, a, b)
    for i in range(n):
        assert to_int(a[i]) == M.to_mont(am[i], M.Q)
        assert to_int(c[i]) == M.to_mont(am[i] * bm[i] % M.Q, M.Q), i
    one = oracle.fe_one(FQ)
    neg_one = oracle.fe_batch(FQ, oracle.OP_NEG, one)
    x = one
    for i in range(8):  # `n *= -1` of the bench alternates between -1 and 1
        x = oracle.fe_batch(FQ, oracle.OP_MUL, x, neg_one)
        assert (x == (neg_one if i % 2 == 0 else one)).all()
    four = oracle.fe_batch(FQ, oracle.OP_DOUBLE, oracle.fe_batch(FQ, oracle.OP_DOUBLE, one))
    assert to_int(four[0]) == M.to_mont(4, M.Q)
