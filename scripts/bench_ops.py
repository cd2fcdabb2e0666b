"""Throughput of every kernel family on device-resident batches (secondary to bench.py).
Writes gpurun_out/bench_ops.json: ms per launch, units/s, algorithmic GB/s."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import jubjub_b200 as jj  # noqa: E402
from scripts.run_smul import SEED0, generator  # noqa: E402


def timed(eng, fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    eng.sync()
    best = 1e30
    for _ in range(reps):
        eng.timer_start()
        fn()
        best = min(best, eng.timer_stop())
    return best


def main():
    eng = jj.Engine(0)
    n = 1 << int(os.environ.get("OPS_LOGN", "20"))
    res = {"n": n}
    A = jj.JJ_ASYNC

    def rec(name, ms, bytes_per_unit):
        res[name] = {"ms": ms, "units_per_s": n / ms * 1e3, "algorithmic_GBps": n * bytes_per_unit / ms / 1e6}
        print(f"{name:28s} {ms:9.3f} ms  {n / ms * 1e3:.3e}/s  {n * bytes_per_unit / ms / 1e6:8.1f} GB/s", flush=True)

    for field in ("fq", "fr"):
        a = eng.fe_stream(field, SEED0, n, device=True)
        b = eng.fe_stream(field, SEED0 + 1, n, device=True)
        o = eng.empty((n, 4))
        rec(f"{field}_mul", timed(eng, lambda: eng.fe_mul(field, a, b, out=o, flags=A)), 96)
        rec(f"{field}_square", timed(eng, lambda: eng.fe_square(field, a, out=o, flags=A)), 64)
        rec(f"{field}_add", timed(eng, lambda: eng.fe_add(field, a, b, out=o, flags=A)), 96)
        rec(f"{field}_sub", timed(eng, lambda: eng.fe_sub(field, a, b, out=o, flags=A)), 96)
        rec(f"{field}_neg", timed(eng, lambda: eng.fe_neg(field, a, out=o, flags=A)), 64)
        ok = eng.empty((n, 1), np.uint8)
        lib, ctx = eng.lib, eng.ctx
        rec(f"{field}_invert", timed(eng, lambda: eng._check(getattr(lib, f"jj_{field}_invert")(
            ctx, a.ptr, o.ptr, ok.ptr, n, jj.JJ_DEVICE_PTRS | A)), reps=3), 64)
        for x in (a, b, o, ok):
            x.free()
    t = eng.fe_to_bytes("fr", eng.fe_stream("fr", SEED0 + 3, n, device=True))
    k = eng.fe_to_bytes("fr", eng.fe_stream("fr", SEED0 + 2, n, device=True))
    gen = generator(eng)
    p = eng.empty((n, 20))
    rec("scalar_mul_fixed", timed(eng, lambda: eng.scalar_mul_fixed_vartime(gen, t, out=p), reps=3), 192)
    q = eng.point_double(p)
    o = eng.empty((n, 20))
    rec("point_double", timed(eng, lambda: eng._check(eng.lib.jj_point_double(eng.ctx, p.ptr, o.ptr, n, jj.JJ_DEVICE_PTRS | A))), 320)
    rec("point_add (ext+ext)", timed(eng, lambda: eng._check(eng.lib.jj_point_add(eng.ctx, p.ptr, q.ptr, o.ptr, n, jj.JJ_DEVICE_PTRS | A))), 480)
    nq = eng.point_to_niels(q)
    rec("point_add_niels", timed(eng, lambda: eng._check(eng.lib.jj_point_add_niels(eng.ctx, p.ptr, nq.ptr, o.ptr, n, jj.JJ_DEVICE_PTRS | A))), 448)
    rec("scalar_mul (variable base)", timed(eng, lambda: eng.scalar_mul_vartime(p, k, out=o, flags=A), reps=3), 352)
    one = eng.empty((1, 20))
    rec("point_sum (one sum of the batch)", timed(eng, lambda: eng._check(eng.lib.jj_point_sum(
        eng.ctx, o.ptr, one.ptr, 1, n, jj.JJ_DEVICE_PTRS | A))), 160)
    grp = eng.empty((n // 64, 20))
    rec("point_sum (groups of 64)", timed(eng, lambda: eng._check(eng.lib.jj_point_sum(
        eng.ctx, o.ptr, grp.ptr, n // 64, 64, jj.JJ_DEVICE_PTRS | A))), 162.5)
    aff = eng.empty((n, 8))
    rec("batch_normalize", timed(eng, lambda: eng.batch_normalize(o, out=aff)), 224)
    enc = eng.empty((n, 32), np.uint8)
    rec("affine_to_bytes", timed(eng, lambda: eng.affine_to_bytes(aff, out=enc)), 96)
    back, ok = eng.empty((n, 8)), eng.empty((n, 1), np.uint8)
    rec("batch_from_bytes", timed(eng, lambda: eng._check(eng.lib.jj_batch_from_bytes(
        eng.ctx, enc.ptr, back.ptr, ok.ptr, n, jj.JJ_DEVICE_PTRS)), reps=2, warm=1), 96)
    flg = eng.empty((n, 1), np.uint8)
    rec("is_torsion_free (pairing)", timed(eng, lambda: eng._check(eng.lib.jj_is_torsion_free(
        eng.ctx, p.ptr, flg.ptr, n, jj.JJ_DEVICE_PTRS | A)), reps=3, warm=1), 97)
    rec("is_torsion_free (ladder [r]P)", timed(eng, lambda: eng._check(eng.lib.jj_is_torsion_free(
        eng.ctx, p.ptr, flg.ptr, n, jj.JJ_DEVICE_PTRS | A | jj.JJ_TORSION_LADDER)), reps=2, warm=1), 161)
    rec("is_prime_order", timed(eng, lambda: eng._check(eng.lib.jj_is_prime_order(
        eng.ctx, p.ptr, flg.ptr, n, jj.JJ_DEVICE_PTRS | A)), reps=3, warm=1), 97)
    rec("mul_by_cofactor", timed(eng, lambda: eng._check(eng.lib.jj_mul_by_cofactor(
        eng.ctx, p.ptr, o.ptr, n, jj.JJ_DEVICE_PTRS | A))), 320)
    outb = eng.empty((n, 32), np.uint8)
    enc_call = lambda f: eng._check(eng.lib.jj_scalar_mul_encoded(  # noqa: E731
        eng.ctx, enc.ptr, k.ptr, outb.ptr, ok.ptr, n, jj.JJ_DEVICE_PTRS | A | jj.JJ_OUT_BYTES | f))
    rec("scalar_mul_encoded (bytes->bytes)", timed(eng, lambda: enc_call(0), reps=2, warm=1), 97)
    rec("scalar_mul_encoded + subgroup check", timed(eng, lambda: enc_call(jj.JJ_CHECK_SUBGROUP), reps=2, warm=1), 97)
    for v, name in ((200, "fused normalise epilogue"), (201, "separate normalise pass")):
        eng.set_scalar_mul_variant(v)
        rec(f"scalar_mul -> bytes ({name})", timed(eng, lambda: eng._check(eng.lib.jj_scalar_mul(
            eng.ctx, p.ptr, k.ptr, outb.ptr, n, jj.JJ_DEVICE_PTRS | A | jj.JJ_OUT_BYTES)), reps=3, warm=1), 224)
    rec("scalar_mul (JJ_CONST_TIME)", timed(eng, lambda: eng.scalar_mul(p, k, out=o, flags=A), reps=3), 352)
    eng.set_scalar_mul_variant(24)
    rec("scalar_mul (24 warps/SM)", timed(eng, lambda: eng.scalar_mul_vartime(p, k, out=o, flags=A), reps=3), 352)
    eng.set_scalar_mul_variant(0)
    # launch-bound chain: 48 small field batches (n = 4096) eagerly vs replayed as one CUDA graph
    m = 4096
    x, y, t2 = eng.fe_stream("fq", 1, m, device=True), eng.fe_stream("fq", 2, m, device=True), eng.empty((m, 4))

    def chain():
        for _ in range(16):
            eng.fe_mul("fq", x, y, out=t2, flags=A)
            eng.fe_add("fq", t2, y, out=t2, flags=A)
            eng.fe_square("fq", t2, out=x, flags=A)

    eager = timed(eng, chain)
    g = eng.graph_capture(chain)
    graph = timed(eng, lambda: eng.graph_launch(g))
    res["chain48_n4096"] = {"eager_ms": eager, "graph_ms": graph}
    print(f"48-kernel chain at n=4096: eager {eager:.3f} ms, CUDA graph {graph:.3f} ms", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bench_ops.json", "w"), indent=1)


if __name__ == "__main__":
    main()
