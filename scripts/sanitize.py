"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import jubjub_b200 as jj  # noqa: E402
from scripts.run_smul import SEED0, generator  # noqa: E402

eng = jj.Engine(0)
n = 700
a, b = eng.fe_stream("fq", SEED0, n), eng.fe_stream("fq", SEED0 + 1, n)
for f in ("fq", "fr"):
    eng.fe_mul(f, a, b); eng.fe_add(f, a, b); eng.fe_sub(f, a, b); eng.fe_square(f, a); eng.fe_neg(f, a)
    eng.fe_invert(f, a[:64]); eng.fe_sqrt(f, a[:64]); eng.fe_from_bytes(f, eng.fe_to_bytes(f, a))
g = generator(eng)
k = eng.fe_to_bytes("fr", eng.fe_stream("fr", SEED0 + 2, n))
p = eng.scalar_mul_fixed_vartime(g, k)
eng.set_scalar_mul_variant(100); eng.scalar_mul_fixed_vartime(g, k[:100]); eng.set_scalar_mul_variant(0)
q = eng.point_double(p)
eng.point_add(p, q); eng.point_add_niels(p, eng.point_to_niels(q)); eng.point_add_affine_niels(p, eng.affine_to_niels(eng.batch_normalize(q)))
for v in (0, 24, 5, 200, 201):
    eng.set_scalar_mul_variant(v)
    out = eng.scalar_mul_vartime(p, k, output="bytes")
    dev = eng.scalar_mul_vartime(eng.to_device(p), eng.to_device(k), output="bytes").download()
    assert (dev == out).all()
eng.set_scalar_mul_variant(0)
pts, ok = eng.batch_from_bytes(out)
assert ok.all()
eng.is_torsion_free(p); eng.is_torsion_free(p[:64], ladder=True); eng.is_prime_order(p); eng.is_identity(p); eng.is_small_order(p)
eng.mul_by_cofactor(p); eng.batch_normalize_extended(q)
assert eng.point_eq(p, eng.point_neg(eng.point_neg(p))).all(); eng.affine_to_extended(eng.batch_normalize(q))
enc, ok2 = eng.scalar_mul_encoded_vartime(out, k, check_subgroup=True)
# the same through page-locked buffers: encodings read in place, scalars uploaded behind the decode, results stored in place
from bench import pinned  # noqa: E402
hp = [pinned(eng, (len(k), 32), np.uint8) for _ in range(3)] + [pinned(eng, (len(k),), np.uint8)]
hp[0][0][:], hp[1][0][:] = out, k
eng._check(eng.lib.jj_scalar_mul_encoded(eng.ctx, hp[0][0].ctypes.data, hp[1][0].ctypes.data, hp[2][0].ctypes.data,
                                         hp[3][0].ctypes.data, len(k), jj.JJ_OUT_BYTES | jj.JJ_CHECK_SUBGROUP))
assert (hp[3][0] == ok2).all() and (hp[2][0][ok2 == 1] == enc[ok2 == 1]).all()
d = eng.to_device(p); eng.scalar_mul_vartime(d, eng.to_device(k), output="affine").download()
# chains longer than one element per thread (Montgomery-trick kernels) at a size compute-sanitizer finishes quickly:
# the grids are capped at one 128-thread block per 128 elements, so force chains by calling with few elements is not
# possible -- these calls cover the single-element chains, the long chains are covered by tests/test_gpu_parity.py
eng.fe_invert("fq", a); eng.fe_invert("fr", a, flags=jj.JJ_CANON)
eng.fe_sqrt("fq", eng.fe_square("fq", a[:200]))
print("sanitize run ok")
