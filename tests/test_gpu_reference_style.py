"""The reference's own point tests (src/lib.rs:1456-1935) written against the reference-style batch
types of jubjub_b200.types -- same names, same assertions, each object a batch on the GPU."""
import numpy as np
import pytest

from tests.golden import reference_kats as K

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    from jubjub_b200 import types

    # `point * scalar` is constant-time in the scalar like the reference's (src/lib.rs:12-17): it runs the JJ_CONST_TIME
    # kernel mode; mul_vartime() is the explicitly named fast form.  These tests run the operators in BOTH modes.
    types.acknowledge_vartime(False)
    p = types.ExtendedPoint.identity()
    assert (p * types.Fr.one()).is_identity()[0] and p.mul_vartime(types.Fr.one()).is_identity()[0]
    return types


@pytest.fixture(params=["constant-time", "vartime"], autouse=True)
def _operator_mode(request, T):
    T.acknowledge_vartime(request.param == "vartime")
    yield
    T.acknowledge_vartime(False)


def _raw(limbs):
    return sum(int(v) << (64 * i) for i, v in enumerate(limbs))


def _test_point(T):
    """the point used by test_assoc / test_batch_normalize / test_mul_consistency, times the cofactor"""
    a = T.AffinePoint.from_raw_unchecked(T.Fq.from_raw([_raw(K.TEST_POINT_RAW[0])]), T.Fq.from_raw([_raw(K.TEST_POINT_RAW[1])]))
    return T.ExtendedPoint.from_affine(a).mul_by_cofactor()


def test_niels_point_identities(T):  # src/lib.rs:1469-1502
    one, zero = T.Fq.one(), T.Fq.zero()
    n = T.AffineNielsPoint.identity()
    assert (n.data[:, 0:4] == one.limbs).all() and (n.data[:, 4:8] == one.limbs).all() and (n.data[:, 8:12] == zero.limbs).all()
    e = T.ExtendedNielsPoint.identity()
    assert (e.data[:, 0:4] == one.limbs).all() and (e.data[:, 4:8] == one.limbs).all()
    assert (e.data[:, 8:12] == one.limbs).all() and (e.data[:, 12:16] == zero.limbs).all()


def test_assoc(T):  # src/lib.rs:1505-1527
    p = _test_point(T)
    assert (p * T.Fr.from_u64(1000)) * T.Fr.from_u64(3938) == p * (T.Fr.from_u64(1000) * T.Fr.from_u64(3938))


def test_batch_normalize(T):  # src/lib.rs:1530-1575
    p = _test_point(T)
    v = []
    for _ in range(10):
        v.append(p.data[0].copy())
        p = p.double()
    v = T.ExtendedPoint(np.array(v))
    expected = [T.ExtendedPoint(v.data[i:i + 1]).to_affine() for i in range(10)]  # one inversion each
    result = T.batch_normalize(v)
    for i in range(10):
        assert expected[i] == T.AffinePoint(result.data[i:i + 1])
    assert T.ExtendedPoint.from_affine(result) == v
    # the reference normalises `v` itself: z = 1, t1 = u, t2 = v (src/lib.rs:1088-1100)
    assert (v.data[:, 8:12] == T.Fq.one(10).limbs).all()
    assert (v.data[:, 12:16] == v.data[:, 0:4]).all() and (v.data[:, 16:20] == v.data[:, 4:8]).all()


def test_eight_torsion_and_small_order(T):  # src/lib.rs:1589-1754
    tors = T.AffinePoint.from_raw_unchecked(T.Fq.from_raw([_raw(p[0]) for p in K.EIGHT_TORSION_RAW]),
                                            T.Fq.from_raw([_raw(p[1]) for p in K.EIGHT_TORSION_RAW]))
    assert tors.is_small_order().all()
    assert tors.mul_by_cofactor().is_identity().all()
    g = T.ExtendedPoint.from_affine(T.AffinePoint.generator())
    assert not g.is_small_order()[0]
    t = g.multiply_bits(np.array([K.FR_MODULUS_BYTES], dtype=np.uint8))
    assert t.is_small_order()[0]
    cur = t
    for i in range(8):  # find_eight_torsion
        assert cur.to_affine() == T.AffinePoint(tors.data[i:i + 1]), i
        cur = cur + t
    assert g.mul_by_cofactor().is_torsion_free()[0] and not g.is_torsion_free()[0]
    assert g.mul_by_cofactor().is_prime_order()[0] and not T.ExtendedPoint.identity().is_prime_order()[0]


def test_mul_consistency(T):  # src/lib.rs:1757-1804
    a, b, c = T.Fr(np.array([K.MULC_A], dtype=np.uint64)), T.Fr(np.array([K.MULC_B], dtype=np.uint64)), \
        T.Fr(np.array([K.MULC_C], dtype=np.uint64))
    assert a * b == c
    p = _test_point(T)
    assert p * c == (p * a) * b
    # Mul implemented on ExtendedNielsPoint
    assert p * c == (p.to_niels() * a) * b
    assert p.to_niels() * c == (p * a) * b
    assert p.to_niels() * c == (p.to_niels() * a) * b
    # Mul implemented on AffineNielsPoint
    pan = p.to_affine().to_niels()
    assert p * c == (pan * a) * b
    assert pan * c == (p * a) * b
    assert pan * c == (pan * a) * b
    # and on AffinePoint (src/lib.rs:1109-1115)
    assert p.to_affine() * c == (p * a) * b


def test_serialization_consistency(T):  # src/lib.rs:1807-1890
    gen = T.AffinePoint.generator().mul_by_cofactor()
    want = np.array(K.SERIALIZED_MULTIPLES_OF_8G, dtype=np.uint8)
    batched, ok = T.AffinePoint.batch_from_bytes(want)
    assert ok.all()
    p = gen
    for i in range(16):
        affine = p.to_affine()
        serialized = affine.to_bytes()
        deserialized, ok1 = T.AffinePoint.from_bytes(serialized)
        assert ok1[0] and affine == deserialized
        assert affine == T.AffinePoint(batched.data[i:i + 1])
        assert (serialized[0] == want[i]).all()
        p = p + gen


def test_zip_216(T):  # src/lib.rs:1893-1935
    for enc in K.ZIP216_NON_CANONICAL:
        b = np.array([enc], dtype=np.uint8)
        assert T.AffinePoint.from_bytes(b)[1][0] == 0
        cleared = b.copy()
        cleared[0, 31] &= 0x7F
        assert T.AffinePoint.from_bytes(cleared)[1][0] == 1
        parsed, ok = T.AffinePoint.from_bytes_pre_zip216_compatibility(b)
        assert ok[0] == 1
        encoded = parsed.to_bytes()
        assert (encoded != b).any()
        encoded[0, 31] |= 0x80
        assert (encoded == b).all()


def test_subgroup_point_surface(T, oracle):  # SubgroupPoint / CofactorGroup / GroupEncoding, src/lib.rs:1122-1239, 1287-1354, 1407-1434
    n = 64
    g = T.ExtendedPoint.generator(n)
    k = T.Fr.from_u64(list(range(1, n + 1)))
    full = g * k                                   # [k]G, G of order 8r: torsion free iff 8 | k
    sub, ok = full.into_subgroup()                 # CtOption(SubgroupPoint(self), is_torsion_free)
    assert isinstance(sub, T.SubgroupPoint) and sub.to_extended() == full
    assert (ok.astype(bool) == (np.arange(1, n + 1) % 8 == 0)).all()
    assert (full.is_torsion_free() == oracle.is_torsion_free(full.data)).all()
    cleared = full.clear_cofactor()                # SubgroupPoint([8]P)
    assert cleared.to_extended().is_torsion_free().all() and cleared.to_extended() == full.mul_by_cofactor()
    sg = T.SubgroupPoint.generator(n)              # ExtendedPoint::generator().clear_cofactor()
    assert sg.to_extended() == g.mul_by_cofactor() and not sg.is_identity().any()
    assert (sg * k).to_extended() == cleared.to_extended()          # Mul<&Fr> for &SubgroupPoint
    assert (sg * k).mul_vartime(T.Fr.one(n)) == sg * k
    assert (sg + sg) == sg.double() and (sg - sg).is_identity().all() and (-sg + sg).is_identity().all()
    assert (full + sg) == (full + sg.to_extended()) and (full - sg) == (full - sg.to_extended())  # ExtendedPoint +- SubgroupPoint
    assert T.SubgroupPoint.identity(3).is_identity().all()
    # Sum<SubgroupPoint>: sum_{k=1..n} [k] G8 = [n (n + 1) / 2] G8
    tot = (sg * k).sum()
    assert tot == T.SubgroupPoint.generator(1) * T.Fr.from_u64([n * (n + 1) // 2])
    # GroupEncoding: to_bytes == the affine encoding; from_bytes = decode AND torsion free; from_bytes_unchecked = decode only
    enc = full.to_bytes()
    assert (enc == full.to_affine().to_bytes()).all()
    assert (enc == oracle.affine_to_bytes(oracle.batch_normalize(full.data))).all()
    back, ok1 = T.ExtendedPoint.from_bytes(enc)
    assert ok1.all() and back == full
    sp, ok2 = T.SubgroupPoint.from_bytes(enc)
    assert (ok2.astype(bool) == (np.arange(1, n + 1) % 8 == 0)).all()
    assert sp.to_extended().is_identity()[ok2 == 0].all()  # rejected elements are the identity
    assert T.ExtendedPoint(sp.to_extended().data[ok2 == 1]) == T.ExtendedPoint(full.data[ok2 == 1])
    su, ok3 = T.SubgroupPoint.from_bytes_unchecked(enc)
    assert ok3.all() and su.to_extended() == full
    bad = enc.copy()
    bad[0] = 0xFF
    assert T.SubgroupPoint.from_bytes(bad)[1][0] == 0 and T.ExtendedPoint.from_bytes(bad)[1][0] == 0
    # from_raw_unchecked (src/lib.rs:1148-1158)
    a8 = cleared.to_extended().to_affine()
    assert T.SubgroupPoint.from_raw_unchecked(a8.get_u(), a8.get_v()) == cleared


def test_field_surface(T):  # src/fr.rs:1045-1175 through the operator surface
    big = T.Fr(np.array([K.FR_LARGEST], dtype=np.uint64))
    assert big + big == T.Fr(np.array([K.FR_LARGEST_PLUS_LARGEST], dtype=np.uint64))
    assert big + T.Fr(np.array([[1, 0, 0, 0]], dtype=np.uint64)) == T.Fr.zero()
    assert -big == T.Fr(np.array([[1, 0, 0, 0]], dtype=np.uint64))
    assert big - big == T.Fr.zero()
    inv, ok = T.Fr.zero().invert()
    assert ok[0] == 0
    for F in (T.Fq, T.Fr):
        x = F.from_u64([7, 11, 13])
        inv, ok = x.invert()
        assert ok.all() and x * inv == F.one(3)
        assert x.square() == x * x and x.double() == x + x
        back, ok = F.from_bytes(x.to_bytes())
        assert ok.all() and back == x
    with pytest.raises(ValueError):  # the reference panics on a length mismatch (src/lib.rs:841)
        T.ExtendedPoint.identity(2) + T.ExtendedPoint.identity(3)
