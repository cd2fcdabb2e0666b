"""GPU fuzz wall behind the carry-chain arithmetic.

The field kernels carry the PTX condition code across separate `asm volatile` statements, and two ptxas miscompiles
of *other* formulations were met on sm_100a / CUDA 12.9.  The host emulation cannot see a
ptxas bug, so the shipped formulation is differentially tested ON THE DEVICE: 2^24 operand pairs per field --
uniformly random and structured (limbs drawn from {0, 1, 2, 2^31, 2^32-1, 2^32-2, ...} so that quotient digits are 0,
carries ripple through every column, the squaring's fold boundary is straddled) -- through mul (both reduction-row
formulations), square, add, sub, every limb compared with the C oracle (src/fr.rs:353-381, 544-665).
__graft_entry__.VALIDATED_NVCC names the toolchain this wall was last run green on."""
import numpy as np
import pytest

from oracle import model as M

pytestmark = pytest.mark.gpu
FQ, FR = 0, 1
import os

SLAB = 1 << 22
SLABS = int(os.environ.get("JJ_FUZZ_SLABS", "4"))  # default 2^24 pairs per field; JJ_FUZZ_SLABS=32 -> 2^27 (one-off runs)


@pytest.fixture(scope="module")
def eng():
    import jubjub_b200 as jj

    e = jj.Engine(0)
    yield e
    e.close()


def _structured(rng, n, m):
    """(n, 4) uint64 field elements < m whose 32-bit limbs come from a palette of carry-provoking values."""
    top = m >> 224
    palette = np.array([0, 1, 2, 3, 0x80000000, 0x7FFFFFFF, 0xFFFFFFFF, 0xFFFFFFFE, 0xFFFF0000, 0x0000FFFF, 0x00010000,
                        0xAAAAAAAA, 0x55555555], dtype=np.uint64)
    pick = rng.integers(0, len(palette) + 6, size=(n, 8))
    rnd = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64)
    limbs = np.where(pick < len(palette), palette[np.minimum(pick, len(palette) - 1)], rnd)
    # top limb strictly below m's top limb keeps the value canonical; straddle m7/2 (the squaring's fold) and m7 - 1
    tops = np.array([0, 1, (top >> 1) - 1, top >> 1, (top >> 1) + 1, top - 1, top - 2], dtype=np.uint64)
    tpick = rng.integers(0, len(tops) + 4, size=n)
    limbs[:, 7] = np.where(tpick < len(tops), tops[np.minimum(tpick, len(tops) - 1)], rnd[:, 7] % np.uint64(top))
    return np.ascontiguousarray((limbs[:, 0::2] | (limbs[:, 1::2] << np.uint64(32))).astype(np.uint64))


def _random(rng, n, m):
    top = m >> 224
    limbs = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64)
    limbs[:, 7] %= np.uint64(top)
    return np.ascontiguousarray((limbs[:, 0::2] | (limbs[:, 1::2] << np.uint64(32))).astype(np.uint64))


@pytest.mark.parametrize("which,name", [(FQ, "fq"), (FR, "fr")])
def test_fuzz_wall_16m_pairs(eng, oracle, which, name):
    m = M.Q if which == FQ else M.R_ORDER
    rng = np.random.default_rng(0xB200 + which)
    total = 0
    for slab in range(SLABS):
        gen = _structured if slab % 2 == 0 else _random
        a, b = gen(rng, SLAB, m), gen(rng, SLAB, m)
        if slab == 0:  # the generators really produce canonical values, and the palette really bites
            for i in range(0, SLAB, SLAB // 64):
                assert M.from_limbs(a[i]) < m and M.from_limbs(b[i]) < m
            assert (a == 0).any() and ((a & np.uint64(0xFFFFFFFF)) == np.uint64(0xFFFFFFFF)).any()
        da, db = eng.to_device(a), eng.to_device(b)
        for op, fn in ((oracle.OP_MUL, eng.fe_mul), (oracle.OP_ADD, eng.fe_add), (oracle.OP_SUB, eng.fe_sub)):
            got = fn(name, da, db).download()
            want = oracle.fe_batch(which, op, a, b)
            bad = np.flatnonzero((got != want).any(axis=1))
            assert bad.size == 0, (name, op, slab, bad[:4], a[bad[:2]], b[bad[:2]])
        got = eng.fe_square(name, da).download()
        want = oracle.fe_batch(which, oracle.OP_SQUARE, a)
        bad = np.flatnonzero((got != want).any(axis=1))
        assert bad.size == 0, (name, "square", slab, bad[:4], a[bad[:2]])
        total += SLAB
        for x in (da, db):
            x.free()
    assert total == SLABS * SLAB


def test_fuzz_point_formulas_2m(eng, oracle):
    """The point kernels use the OTHER reduction-row formulation (M1MUL = true, subtractive last row): 2^21 doublings
    and additions on points with random projective scaling, all 160 bytes against the oracle."""
    n = 1 << 21
    t = oracle.fe_to_bytes(FR, oracle.fe_stream(FR, 0xF00D, n))
    p = eng.scalar_mul_fixed_vartime(oracle.generator(), eng.to_device(t))
    q = eng.point_double(p)                       # z != 1
    d = eng.point_double(q).download()
    s = eng.point_add(q, p).download()
    ph, qh = p.download(), q.download()
    step = 8  # the oracle is single-threaded for elementwise point ops: every 8th unit = 2^18 units
    assert (d[::step] == oracle.ext_double(np.ascontiguousarray(qh[::step]))).all()
    assert (s[::step] == oracle.ext_add(np.ascontiguousarray(qh[::step]), np.ascontiguousarray(ph[::step]))).all()
