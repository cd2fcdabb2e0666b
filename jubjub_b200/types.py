"""Batch-valued mirrors of the reference's public types, so that host code and the parity tests
read like the reference's own (`src/lib.rs`, `src/fr.rs`): every object holds a *batch* of n values
and every operator acts element-wise on the GPU through the C ABI.

    p = ExtendedPoint.from_affine(AffinePoint.generator(n)).mul_by_cofactor()
    assert (p * a) * b == p * (a * b)                       # src/lib.rs:1505-1527
    enc = batch_normalize(p * k).to_bytes()                 # src/lib.rs:1084-1107, 455-464

Names, argument meaning and failure behaviour follow the reference: `invert()` / `from_bytes()`
return `(value, is_some)` like `CtOption` (src/fr.rs:268-292, 438-540), a length mismatch raises
(`assert_eq!`, src/lib.rs:841), `ExtendedPoint.__eq__` is projective equality (src/lib.rs:153-163),
scalar multiplication takes `Fr` in Montgomery form and ignores nothing but what `Fr` cannot hold.
There is no CPU arithmetic here: without a B200 the engine constructor raises.

Constant time: the reference's `*` is constant-time by policy (src/lib.rs:12-17).  Here `point * scalar` runs the
engine's constant-time-in-the-scalar kernel mode (JJ_CONST_TIME: no branch or address depends on the scalar); the
fast variable-time kernels are reached through explicitly named methods, `mul_vartime` / `batch_mul_vartime`, following
the reference's naming rule (src/lib.rs:14-15) -- or through `*` after the caller has opted in once with
`acknowledge_vartime()` (public scalars only).
"""
import numpy as np

from .engine import default_engine

_GEN_RAW = (0x62EDCBB8BF3787C88B0F03DDD60A8187CAF55D1B29BF81AFE4B3D35DF1A7ADFE, 11)  # src/lib.rs:1380-1396
_EDWARDS_D2_RAW = 0x552631CE97F45691EBFB240FCD7AFFA8525AFEDA6EAF3A4C020CBFADAC687D62  # 2d, src/lib.rs:407-412


_VARTIME_ACK = False


def acknowledge_vartime(on=True):
    """Let the `point * scalar` operators run the fast variable-time kernels (public scalars only).  Without it they run
    the constant-time mode."""
    global _VARTIME_ACK
    _VARTIME_ACK = bool(on)


def _limbs(x):
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


class _Field:
    """Common implementation of Fq / Fr batches; `limbs` is (n, 4) uint64 Montgomery form."""

    _name = None

    def __init__(self, limbs, engine=None):
        self.limbs = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1, 4)
        self.eng = engine or default_engine()

    def __len__(self):
        return len(self.limbs)

    # ---- constructors (src/fr.rs:246-349) ----------------------------------------------------
    @classmethod
    def zero(cls, n=1, engine=None):
        return cls(np.zeros((n, 4), dtype=np.uint64), engine)

    @classmethod
    def from_raw(cls, values, engine=None):
        """from_raw([u64; 4]) for each Python int / limb list: the integer's residue in Montgomery form."""
        eng = engine or default_engine()
        raw = np.array([_limbs(v) if isinstance(v, int) else list(v) for v in values], dtype=np.uint64).reshape(-1, 4)
        wide = np.zeros((len(raw), 64), dtype=np.uint8)
        wide[:, :32] = raw.view(np.uint8).reshape(-1, 32)
        return cls(eng.fe_from_bytes_wide(cls._name, wide), eng)  # d0*R2 + 0*R3 == from_raw(d0)

    @classmethod
    def one(cls, n=1, engine=None):
        return cls.from_raw([1] * n, engine)

    @classmethod
    def from_u64(cls, values, engine=None):
        return cls.from_raw([int(v) for v in np.atleast_1d(values)], engine)

    @classmethod
    def from_bytes(cls, b, engine=None):
        """-> (value, is_some); is_some[i] = 0 when the encoding is not canonical (src/fr.rs:268-292)."""
        eng = engine or default_engine()
        v, ok = eng.fe_from_bytes(cls._name, np.ascontiguousarray(b, dtype=np.uint8).reshape(-1, 32))
        return cls(v, eng), ok

    @classmethod
    def from_bytes_wide(cls, b, engine=None):
        eng = engine or default_engine()
        return cls(eng.fe_from_bytes_wide(cls._name, np.ascontiguousarray(b, dtype=np.uint8).reshape(-1, 64)), eng)

    # ---- arithmetic ------------------------------------------------------------------------------
    def _chk(self, o):
        if not isinstance(o, type(self)):
            raise TypeError(f"{type(self).__name__} op {type(o).__name__}")
        return o.limbs

    def __add__(self, o):
        return type(self)(self.eng.fe_add(self._name, self.limbs, self._chk(o)), self.eng)

    def __sub__(self, o):
        return type(self)(self.eng.fe_sub(self._name, self.limbs, self._chk(o)), self.eng)

    def __mul__(self, o):
        if isinstance(o, type(self)):
            return type(self)(self.eng.fe_mul(self._name, self.limbs, o.limbs), self.eng)
        return NotImplemented

    def __neg__(self):
        return type(self)(self.eng.fe_neg(self._name, self.limbs), self.eng)

    def square(self):
        return type(self)(self.eng.fe_square(self._name, self.limbs), self.eng)

    def double(self):
        return type(self)(self.eng.fe_double(self._name, self.limbs), self.eng)

    def invert(self):
        """-> (inverse, is_some) (src/fr.rs:438-540)."""
        v, ok = self.eng.fe_invert(self._name, self.limbs)
        return type(self)(v, self.eng), ok

    def sqrt(self):
        """-> (root, is_some) (src/fr.rs:384-399)."""
        v, ok = self.eng.fe_sqrt(self._name, self.limbs)
        return type(self)(v, self.eng), ok

    def to_bytes(self):
        return self.eng.fe_to_bytes(self._name, self.limbs)

    def __eq__(self, o):
        return isinstance(o, type(self)) and self.limbs.shape == o.limbs.shape and bool((self.limbs == o.limbs).all())

    __hash__ = None

    def __getitem__(self, i):
        return type(self)(self.limbs[i].reshape(-1, 4), self.eng)

    def __repr__(self):  # Debug prints the canonical value big-endian (src/fr.rs:25-40)
        b = self.to_bytes()
        return f"{type(self).__name__}[" + ", ".join("0x" + bytes(r)[::-1].hex() for r in b[:4]) + (", ...]" if len(b) > 4 else "]")


class Fq(_Field):
    """jubjub::Fq = bls12_381::Scalar (src/lib.rs:62)."""
    _name = "fq"


class Fr(_Field):
    """jubjub::Fr (src/fr.rs:23)."""
    _name = "fr"


class AffineNielsPoint:
    """(v+u, v-u, u*v*2d) (src/lib.rs:255-259); `data` is (n, 12) uint64."""

    def __init__(self, data, engine=None):
        self.data = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, 12)
        self.eng = engine or default_engine()

    @classmethod
    def identity(cls, n=1, engine=None):
        return AffinePoint.identity(n, engine).to_niels()

    def __len__(self):
        return len(self.data)

    def to_affine(self):
        """(u, v) = ((vpu - vmu)/2, (vpu + vmu)/2)."""
        vpu, vmu = Fq(self.data[:, 0:4], self.eng), Fq(self.data[:, 4:8], self.eng)
        half = _half(len(self), self.eng)
        return AffinePoint.from_raw_unchecked((vpu - vmu) * half, (vpu + vmu) * half)

    def __mul__(self, k):
        """`&AffineNielsPoint * &Fr` (src/lib.rs:304-310)."""
        if not isinstance(k, Fr):
            return NotImplemented
        return self.to_affine().to_extended() * k

    def multiply_bits(self, by):
        return self.to_affine().to_extended().multiply_bits(by)


def _half(n, eng):
    inv, _ = Fq.from_raw([2] * n, eng).invert()
    return inv


class ExtendedNielsPoint:
    """(V+U, V-U, Z, T1*T2*2d) (src/lib.rs:327-332); `data` is (n, 16) uint64."""

    def __init__(self, data, engine=None):
        self.data = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, 16)
        self.eng = engine or default_engine()

    @classmethod
    def identity(cls, n=1, engine=None):
        return ExtendedPoint.identity(n, engine).to_niels()

    def __len__(self):
        return len(self.data)

    def to_extended(self):
        """An ExtendedPoint whose to_niels() is this value: U = (vpu - vmu)/2, V = (vpu + vmu)/2, Z = z,
        T1 = t2d / 2d, T2 = 1 (so that T1*T2*2d = t2d)."""
        n = len(self)
        vpu, vmu = Fq(self.data[:, 0:4], self.eng), Fq(self.data[:, 4:8], self.eng)
        z, t2d = Fq(self.data[:, 8:12], self.eng), Fq(self.data[:, 12:16], self.eng)
        half = _half(n, self.eng)
        d2 = Fq.from_raw([_EDWARDS_D2_RAW] * n, self.eng)
        inv_d2, _ = d2.invert()
        u, v, t1 = (vpu - vmu) * half, (vpu + vmu) * half, t2d * inv_d2
        return ExtendedPoint(np.concatenate([u.limbs, v.limbs, z.limbs, t1.limbs, Fq.one(n, self.eng).limbs], axis=1),
                             self.eng)

    def __mul__(self, k):
        """`&ExtendedNielsPoint * &Fr` (src/lib.rs:388-394)."""
        if not isinstance(k, Fr):
            return NotImplemented
        return self.to_extended() * k

    def multiply_bits(self, by):
        return self.to_extended().multiply_bits(by)


class AffinePoint:
    """(u, v) (src/lib.rs:81-84); `data` is (n, 8) uint64 Montgomery limbs."""

    def __init__(self, data, engine=None):
        self.data = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, 8)
        self.eng = engine or default_engine()

    def __len__(self):
        return len(self.data)

    @classmethod
    def from_raw_unchecked(cls, u, v):
        assert isinstance(u, Fq) and isinstance(v, Fq) and len(u) == len(v)
        return cls(np.concatenate([u.limbs, v.limbs], axis=1), u.eng)

    @classmethod
    def identity(cls, n=1, engine=None):
        return cls.from_raw_unchecked(Fq.zero(n, engine), Fq.one(n, engine))

    @classmethod
    def generator(cls, n=1, engine=None):
        return cls.from_raw_unchecked(Fq.from_raw([_GEN_RAW[0]] * n, engine), Fq.from_raw([_GEN_RAW[1]] * n, engine))

    @classmethod
    def batch_from_bytes(cls, b, engine=None):
        """-> (points, is_some) (src/lib.rs:541-627)."""
        eng = engine or default_engine()
        pts, ok = eng.batch_from_bytes(np.ascontiguousarray(b, dtype=np.uint8).reshape(-1, 32))
        return cls(pts, eng), ok

    from_bytes = batch_from_bytes

    @classmethod
    def from_bytes_pre_zip216_compatibility(cls, b, engine=None):
        eng = engine or default_engine()
        pts, ok = eng.batch_from_bytes(np.ascontiguousarray(b, dtype=np.uint8).reshape(-1, 32), zip216=False)
        return cls(pts, eng), ok

    def get_u(self):
        return Fq(self.data[:, :4], self.eng)

    def get_v(self):
        return Fq(self.data[:, 4:], self.eng)

    def to_bytes(self):
        return self.eng.affine_to_bytes(self.data)

    def to_extended(self):
        return ExtendedPoint.from_affine(self)

    def to_niels(self):
        return AffineNielsPoint(self.eng.affine_to_niels(self.data), self.eng)

    def __neg__(self):
        return AffinePoint.from_raw_unchecked(-self.get_u(), self.get_v())

    def __eq__(self, o):
        return isinstance(o, AffinePoint) and self.data.shape == o.data.shape and bool((self.data == o.data).all())

    __hash__ = None

    def mul_vartime(self, k):
        """`&AffinePoint * &Fr` (src/lib.rs:1109-1115), variable-time; one shared base uses the fixed-base kernel."""
        if len(self) == 1:
            return ExtendedPoint(self.eng.scalar_mul_fixed_vartime(self.data, k.limbs, scalar_mont=True), self.eng)
        return self.to_extended().mul_vartime(k)

    def __mul__(self, k):
        """`&AffinePoint * &Fr`: constant-time in the scalar unless acknowledge_vartime() was called."""
        if not isinstance(k, Fr):
            return NotImplemented
        return self.mul_vartime(k) if _VARTIME_ACK else self.to_extended() * k

    def mul_by_cofactor(self):
        return self.to_extended().mul_by_cofactor()

    def is_identity(self):
        return self.to_extended().is_identity()

    def is_small_order(self):
        return self.to_extended().is_small_order()

    def is_torsion_free(self):
        return self.to_extended().is_torsion_free()

    def is_prime_order(self):
        return self.to_extended().is_prime_order()


class ExtendedPoint:
    """(U, V, Z, T1, T2) (src/lib.rs:139-145); `data` is (n, 20) uint64 Montgomery limbs."""

    def __init__(self, data, engine=None):
        self.data = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, 20)
        self.eng = engine or default_engine()

    def __len__(self):
        return len(self.data)

    @classmethod
    def from_affine(cls, a):
        return cls(a.eng.affine_to_extended(a.data), a.eng)  # (u, v, 1, u, v), src/lib.rs:214-226

    @classmethod
    def identity(cls, n=1, engine=None):
        return cls.from_affine(AffinePoint.identity(n, engine))

    @classmethod
    def generator(cls, n=1, engine=None):
        """The full-order generator (src/lib.rs:1380-1396) as an ExtendedPoint."""
        return cls.from_affine(AffinePoint.generator(n, engine))

    @classmethod
    def from_bytes(cls, b, engine=None):
        """GroupEncoding for ExtendedPoint (src/lib.rs:1407-1422): -> (points, is_some)."""
        a, ok = AffinePoint.batch_from_bytes(b, engine)
        return cls.from_affine(a), ok

    from_bytes_unchecked = from_bytes  # the curve check cannot be skipped when parsing an encoding (:1414-1417)

    def to_bytes(self):
        """`AffinePoint::from(self).to_bytes()` (src/lib.rs:1419-1421): normalise + encode in one pass on the device."""
        return self.eng.batch_normalize_to_bytes(self.data)

    def clear_cofactor(self):
        """CofactorGroup::clear_cofactor (src/lib.rs:1343-1345): [8]P as a SubgroupPoint."""
        return SubgroupPoint(self.mul_by_cofactor())

    def into_subgroup(self):
        """CofactorGroup::into_subgroup (src/lib.rs:1347-1349): -> (SubgroupPoint, is_some = is_torsion_free)."""
        return SubgroupPoint(self), self.is_torsion_free()

    def _same(self, o):
        if len(o) != len(self):
            raise ValueError(f"length mismatch: {len(self)} != {len(o)}")  # assert_eq!, src/lib.rs:841

    def double(self):
        return ExtendedPoint(self.eng.point_double(self.data), self.eng)

    def to_niels(self):
        return ExtendedNielsPoint(self.eng.point_to_niels(self.data), self.eng)

    def _addsub(self, o, subtract):
        self._same(o)
        if isinstance(o, ExtendedPoint):
            return ExtendedPoint(self.eng.point_add(self.data, o.data, subtract=subtract), self.eng)
        if isinstance(o, ExtendedNielsPoint):
            return ExtendedPoint(self.eng.point_add_niels(self.data, o.data, subtract=subtract), self.eng)
        if isinstance(o, AffineNielsPoint):
            return ExtendedPoint(self.eng.point_add_affine_niels(self.data, o.data, subtract=subtract), self.eng)
        if isinstance(o, AffinePoint):  # src/lib.rs:1012-1028
            return self._addsub(o.to_niels(), subtract)
        if isinstance(o, SubgroupPoint):  # &ExtendedPoint +- &SubgroupPoint -> ExtendedPoint, src/lib.rs:1191-1209
            return self._addsub(o.p, subtract)
        return NotImplemented

    def __add__(self, o):
        return self._addsub(o, False)

    def __sub__(self, o):
        return self._addsub(o, True)

    def __neg__(self):  # (-U, V, Z, -T1, T2), src/lib.rs:196-210
        return ExtendedPoint(self.eng.point_neg(self.data), self.eng)

    def mul_vartime(self, k):
        """`&ExtendedPoint * &Fr` (src/lib.rs:873-879), variable-time in the scalar."""
        self._same(k)
        return ExtendedPoint(self.eng.scalar_mul_vartime(self.data, k.limbs, scalar_mont=True), self.eng)

    def __mul__(self, k):
        """`&ExtendedPoint * &Fr` (src/lib.rs:873-879): constant-time in the scalar (JJ_CONST_TIME) unless
        acknowledge_vartime() was called."""
        if not isinstance(k, Fr):
            return NotImplemented
        if _VARTIME_ACK:
            return self.mul_vartime(k)
        self._same(k)
        return ExtendedPoint(self.eng.scalar_mul(self.data, k.limbs, scalar_mont=True), self.eng)

    def multiply_bits(self, by):
        """[k]P for 32 little-endian bytes per point, top four bits ignored (src/lib.rs:381-385)."""
        return ExtendedPoint(self.eng.scalar_mul_vartime(self.data, np.ascontiguousarray(by, dtype=np.uint8).reshape(-1, 32)), self.eng)

    def mul_by_cofactor(self):
        """src/lib.rs:722-724."""
        return ExtendedPoint(self.eng.mul_by_cofactor(self.data), self.eng)

    def sum(self):
        """Sum<ExtendedPoint> (src/lib.rs:183-193): the sum of the whole batch, a batch of one point."""
        return ExtendedPoint(self.eng.point_sum(self.data), self.eng)

    def is_identity(self):
        return self.eng.is_identity(self.data)

    def is_small_order(self):
        return self.eng.is_small_order(self.data)

    def is_torsion_free(self):
        return self.eng.is_torsion_free(self.data)

    def is_prime_order(self):
        """src/lib.rs:717-719."""
        return self.eng.is_prime_order(self.data)

    def to_affine(self):
        return AffinePoint(self.eng.batch_normalize(self.data), self.eng)

    def __eq__(self, o):
        """(u/z, v/z) == (u'/z', v'/z') via u*z' == u'*z and v*z' == v'*z (src/lib.rs:153-163)."""
        if not isinstance(o, ExtendedPoint) or len(o) != len(self):
            return False
        return bool(self.eq(o).all())

    def eq(self, o):
        """Element by element: flags[i] = (self[i] == o[i]) as points (src/lib.rs:153-163)."""
        return self.eng.point_eq(self.data, o.data).astype(bool)

    __hash__ = None

    def __getitem__(self, i):
        return ExtendedPoint(self.data[i].reshape(-1, 20), self.eng)


class SubgroupPoint:
    """A batch of elements of the prime-order subgroup (src/lib.rs:1122: `SubgroupPoint(ExtendedPoint)`).  Like the
    reference type it can only be built by routes that guarantee membership -- `ExtendedPoint.clear_cofactor()`,
    `into_subgroup()` / `from_bytes()` (which report `is_torsion_free` per element), sums, negations, doublings and
    scalar multiples of subgroup points -- or by the explicitly unchecked constructors."""

    def __init__(self, p):
        assert isinstance(p, ExtendedPoint)
        self.p = p
        self.eng = p.eng

    def __len__(self):
        return len(self.p)

    @classmethod
    def from_raw_unchecked(cls, u, v):
        """src/lib.rs:1148-1158: the caller vouches for membership."""
        return cls(AffinePoint.from_raw_unchecked(u, v).to_extended())

    @classmethod
    def identity(cls, n=1, engine=None):
        return cls(ExtendedPoint.identity(n, engine))

    @classmethod
    def generator(cls, n=1, engine=None):
        """`ExtendedPoint::generator().clear_cofactor()` (src/lib.rs:1304-1306)."""
        return ExtendedPoint.generator(n, engine).clear_cofactor()

    @classmethod
    def from_bytes(cls, b, engine=None):
        """GroupEncoding for SubgroupPoint (src/lib.rs:1427-1429): decoded AND torsion free -> (points, is_some).  The
        subgroup test runs on the device (a pairing, not `[r]P`); rejected elements are the identity."""
        p, ok = ExtendedPoint.from_bytes(b, engine)
        ok = ok.astype(bool) & p.is_torsion_free().astype(bool)
        data = p.data.copy()
        data[~ok] = ExtendedPoint.identity(1, p.eng).data[0]
        return cls(ExtendedPoint(data, p.eng)), ok.astype(np.uint8)

    @classmethod
    def from_bytes_unchecked(cls, b, engine=None):
        """src/lib.rs:1431-1433: curve check only."""
        p, ok = ExtendedPoint.from_bytes(b, engine)
        return cls(p), ok

    def to_bytes(self):
        return self.p.to_bytes()

    def to_extended(self):
        """From<SubgroupPoint> for ExtendedPoint (src/lib.rs:1124-1128)."""
        return self.p

    def __add__(self, o):  # src/lib.rs:1211-1229
        return SubgroupPoint(self.p + o.p) if isinstance(o, SubgroupPoint) else NotImplemented

    def __sub__(self, o):
        return SubgroupPoint(self.p - o.p) if isinstance(o, SubgroupPoint) else NotImplemented

    def __neg__(self):  # src/lib.rs:1173-1189
        return SubgroupPoint(-self.p)

    def double(self):  # src/lib.rs:1312-1315
        return SubgroupPoint(self.p.double())

    def __mul__(self, k):  # src/lib.rs:1231-1239
        r = self.p * k
        return r if r is NotImplemented else SubgroupPoint(r)

    def mul_vartime(self, k):
        return SubgroupPoint(self.p.mul_vartime(k))

    def sum(self):  # Sum<SubgroupPoint>, src/lib.rs:1161-1171
        return SubgroupPoint(self.p.sum())

    def is_identity(self):
        return self.p.is_identity()

    def __eq__(self, o):
        return isinstance(o, SubgroupPoint) and self.p == o.p

    __hash__ = None

    def __getitem__(self, i):
        return SubgroupPoint(self.p[i])


# ---- free functions ---------------------------------------------------------------------------------
def batch_normalize(points):
    """jubjub::batch_normalize (src/lib.rs:1084-1107): normalises the ExtendedPoint batch IN PLACE (z = 1, t1 = u,
    t2 = v, like the reference's `&mut [ExtendedPoint]`) and returns the AffinePoint batch."""
    points.eng.batch_normalize_extended(points.data, in_place=True)
    return AffinePoint(points.data[:, :8], points.eng)


def batch_mul_vartime(points, scalars):
    """New batch entry point: points[i] * scalars[i] (ExtendedPoint x Fr); variable-time in the scalars."""
    return points.mul_vartime(scalars)


def batch_add(p, q):
    """New batch entry point: p[i] + q[i]."""
    return p + q


__all__ = ["Fq", "Fr", "AffinePoint", "ExtendedPoint", "SubgroupPoint", "AffineNielsPoint", "ExtendedNielsPoint", "batch_normalize",
           "batch_mul_vartime", "batch_add", "acknowledge_vartime"]
