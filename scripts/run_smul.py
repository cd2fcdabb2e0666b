"""Runs the variable-base scalar-mul kernel a few times (for ncu captures / launch lists)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import jubjub_b200 as jj  # noqa: E402

SEED0 = 0x4A55424A55420001
GEN_RAW = np.array([[0xe4b3d35df1a7adfe, 0xcaf55d1b29bf81af, 0x8b0f03ddd60a8187, 0x62edcbb8bf3787c8, 0xb, 0, 0, 0]],
                   dtype=np.uint64)


def generator(eng):
    return np.concatenate([eng.fe_from_bytes("fq", GEN_RAW[:, :4].view(np.uint8).reshape(1, 32))[0],
                           eng.fe_from_bytes("fq", GEN_RAW[:, 4:].view(np.uint8).reshape(1, 32))[0]], axis=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, default=17)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--what", default="smul", choices=["smul", "fixed", "fqmul", "torsion", "decode"])
    ap.add_argument("--ct", action="store_true", help="constant-time mode (JJ_CONST_TIME)")
    a = ap.parse_args()
    eng = jj.Engine(0)
    n = 1 << a.logn
    if a.what == "fqmul":
        x, y = eng.fe_stream("fq", SEED0, n, device=True), eng.fe_stream("fq", SEED0 + 1, n, device=True)
        o = eng.empty((n, 4))
        for _ in range(a.reps):
            eng.fe_mul("fq", x, y, out=o)
        return
    t = eng.fe_to_bytes("fr", eng.fe_stream("fr", SEED0 + 3, n, device=True))
    k = eng.fe_to_bytes("fr", eng.fe_stream("fr", SEED0 + 2, n, device=True))
    if a.what == "fixed":
        for _ in range(a.reps):
            eng.scalar_mul_fixed_vartime(generator(eng), k)
        return
    pts = eng.scalar_mul_fixed_vartime(generator(eng), t)
    if a.what == "decode":
        enc = eng.affine_to_bytes(eng.batch_normalize(pts))
        for _ in range(a.reps):
            eng.timer_start()
            eng.batch_from_bytes(enc)
            print(f"batch_from_bytes n={n}: {eng.timer_stop():.3f} ms")
        return
    if a.what == "torsion":
        for _ in range(a.reps):
            eng.timer_start()
            eng.is_torsion_free(pts)
            print(f"is_torsion_free n={n}: {eng.timer_stop():.3f} ms")
        return
    eng.set_scalar_mul_variant(a.variant)
    o = eng.empty((n, 20))
    for _ in range(a.reps):
        eng.timer_start()
        (eng.scalar_mul if a.ct else eng.scalar_mul_vartime)(pts, k, out=o, flags=jj.JJ_ASYNC)
        ms = eng.timer_stop()
        print(f"n={n} variant={a.variant}: {ms:.3f} ms  {n / ms * 1e3:.4e}/s")


if __name__ == "__main__":
    main()
