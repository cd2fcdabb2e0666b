"""The reference's black-box property tests (tests/fq_blackbox.rs:6-121, tests/fr_blackbox.rs, same
lines with Fr) on the CUDA path: 11 algebraic properties x 2000 random elements per field, every
element drawn like tests/common.rs:15-28 (64 random bytes -> from_bytes_wide).  The byte source
restates rand_xorshift 0.3.0's XorShiftRng with the reference's seed [0..15] (tests/common.rs:7-9)
from the crate's published algorithm; the properties do not depend on the exact stream."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NUM_BLACK_BOX_CHECKS = 2000  # tests/common.rs:5


class XorShiftRng:
    def __init__(self, seed=bytes(range(16))):
        self.x, self.y, self.z, self.w = (int.from_bytes(seed[4 * i:4 * i + 4], "little") for i in range(4))

    def next_u32(self):
        t = (self.x ^ (self.x << 11)) & 0xFFFFFFFF
        self.x, self.y, self.z = self.y, self.z, self.w
        self.w = (self.w ^ (self.w >> 19) ^ (t ^ (t >> 8))) & 0xFFFFFFFF
        return self.w

    def fill_bytes(self, n):
        return b"".join(self.next_u32().to_bytes(4, "little") for _ in range(n // 4))


@pytest.fixture(scope="module")
def eng():
    import jubjub_b200 as jj

    e = jj.Engine(0)
    yield e
    e.close()


def _randoms(eng, field, count):
    """`count` interleaved streams of NUM_BLACK_BOX_CHECKS elements, drawn a, b, c, a, b, c, ... like the loops."""
    rng = XorShiftRng()
    wide = np.frombuffer(rng.fill_bytes(64 * count * NUM_BLACK_BOX_CHECKS), dtype=np.uint8).reshape(-1, 64)
    elems = eng.fe_from_bytes_wide(field, wide)
    return [np.ascontiguousarray(elems[i::count]) for i in range(count)]


@pytest.mark.parametrize("field", ["fq", "fr"])
def test_blackbox_properties(eng, field, oracle):
    which = 0 if field == "fq" else 1
    zero = np.zeros((NUM_BLACK_BOX_CHECKS, 4), dtype=np.uint64)
    one = np.repeat(oracle.fe_one(which), NUM_BLACK_BOX_CHECKS, axis=0)
    add, sub, mul, neg = (lambda a, b: eng.fe_add(field, a, b)), (lambda a, b: eng.fe_sub(field, a, b)), \
        (lambda a, b: eng.fe_mul(field, a, b)), (lambda a: eng.fe_neg(field, a))
    (a,) = _randoms(eng, field, 1)
    # test_to_and_from_bytes
    back, ok = eng.fe_from_bytes(field, eng.fe_to_bytes(field, a))
    assert ok.all() and (back == a).all()
    # test_additive_identity / test_subtract_additive_identity / test_additive_inverse
    assert (add(a, zero) == a).all() and (add(zero, a) == a).all()
    assert (sub(a, zero) == a).all() and (sub(zero, neg(a)) == a).all()
    assert (add(a, neg(a)) == 0).all() and (add(neg(a), a) == 0).all()
    # test_multiplicative_identity / test_multiply_additive_identity
    assert (mul(a, one) == a).all() and (mul(one, a) == a).all()
    assert (mul(zero, a) == 0).all() and (mul(a, zero) == 0).all()
    # test_multiplicative_inverse
    inv, ok = eng.fe_invert(field, a)
    nz = a.any(axis=1)
    assert (ok.astype(bool) == nz).all()
    assert (mul(a, inv)[nz] == one[nz]).all() and (mul(inv, a)[nz] == one[nz]).all()
    # commutativity (two draws per iteration)
    a, b = _randoms(eng, field, 2)
    assert (add(a, b) == add(b, a)).all() and (mul(a, b) == mul(b, a)).all()
    # associativity (three draws per iteration)
    a, b, c = _randoms(eng, field, 3)
    assert (add(add(a, b), c) == add(a, add(b, c))).all()
    assert (mul(mul(a, b), c) == mul(a, mul(b, c))).all()
    # and the values themselves against the oracle
    assert (mul(a, b) == oracle.fe_batch(which, oracle.OP_MUL, a, b)).all()


def test_point_group_laws(eng, oracle):
    """Group-law properties on the device path at 4096 points: P + Q == Q + P, (P + Q) - Q == P,
    2P == P + P, [a]([b]P) == [ab]P, [a]P + [b]P == [a + b]P on the prime-order subgroup."""
    from oracle import model as M

    n = 4096
    g8 = oracle.ext_to_affine(oracle.ext_mul_by_cofactor(oracle.affine_to_extended(oracle.generator())))
    t = eng.fe_to_bytes("fr", eng.fe_stream("fr", M.SEED0 + 3, 2 * n))
    pts = eng.scalar_mul_fixed_vartime(g8, t)
    p, q = pts[:n], pts[n:]
    norm = eng.batch_normalize
    assert (norm(eng.point_add(p, q)) == norm(eng.point_add(q, p))).all()
    assert (norm(eng.point_add(eng.point_add(p, q), q, subtract=True)) == norm(p)).all()
    assert (norm(eng.point_double(p)) == norm(eng.point_add(p, p))).all()
    a, b = eng.fe_stream("fr", 11, n), eng.fe_stream("fr", 12, n)
    ab, apb = eng.fe_mul("fr", a, b), eng.fe_add("fr", a, b)
    lhs = eng.scalar_mul_vartime(eng.scalar_mul_vartime(p, b, scalar_mont=True), a, scalar_mont=True, output="affine")
    assert (lhs == eng.scalar_mul_vartime(p, ab, scalar_mont=True, output="affine")).all()
    s = eng.point_add(eng.scalar_mul_vartime(p, a, scalar_mont=True), eng.scalar_mul_vartime(p, b, scalar_mont=True))
    assert (norm(s) == eng.scalar_mul_vartime(p, apb, scalar_mont=True, output="affine")).all()
    assert eng.is_torsion_free(p[:256]).all()
