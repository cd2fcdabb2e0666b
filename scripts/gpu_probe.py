"""Quick on-box probe: integer-pipe peak, Fq-mul bandwidth, scalar-mul rate per kernel variant.
Writes gpurun_out/probe.json.  Not a bench line (bench.py is); used to pick defaults."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import jubjub_b200 as jj  # noqa: E402

SEED0 = 0x4A55424A55420001


def timed(eng, fn, reps=3, warm=1):
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
    out = {}
    eng = jj.Engine(0)
    out["device"] = eng.device_info()
    out["imad_peak_per_s"] = eng.imad_peak()
    print("imad peak %.3e /s" % out["imad_peak_per_s"], flush=True)
    for logn in (20, 26):
        n = 1 << logn
        a = eng.fe_stream("fq", SEED0, n, device=True)
        b = eng.fe_stream("fq", SEED0 + 1, n, device=True)
        o = eng.empty((n, 4))
        for name, fn, bytes_per in (("mul", lambda: eng.fe_mul("fq", a, b, out=o, flags=jj.JJ_ASYNC), 96),
                                    ("add", lambda: eng.fe_add("fq", a, b, out=o, flags=jj.JJ_ASYNC), 96),
                                    ("square", lambda: eng.fe_square("fq", a, out=o, flags=jj.JJ_ASYNC), 64)):
            ms = timed(eng, fn, reps=5)
            out[f"fq_{name}_2^{logn}"] = {"ms": ms, "gops": n / ms / 1e6, "GBps": n * bytes_per / ms / 1e6}
            print(f"fq_{name} n=2^{logn}: {ms:.3f} ms  {n / ms / 1e6:.2f} Gop/s  {n * bytes_per / ms / 1e6:.0f} GB/s", flush=True)
        for x in (a, b, o):
            x.free()
    n = 1 << int(os.environ.get("PROBE_LOGN", "19"))
    g8 = None
    t = eng.fe_to_bytes("fr", eng.fe_stream("fr", SEED0 + 3, n, device=True))
    gen = np.array([[0xe4b3d35df1a7adfe, 0xcaf55d1b29bf81af, 0x8b0f03ddd60a8187, 0x62edcbb8bf3787c8, 0xb, 0, 0, 0]], dtype=np.uint64)
    gen_m = np.concatenate([eng.fe_from_bytes("fq", gen[:, :4].view(np.uint8).reshape(1, 32))[0],
                            eng.fe_from_bytes("fq", gen[:, 4:].view(np.uint8).reshape(1, 32))[0]], axis=1)
    fo = eng.empty((n, 20))
    ms = timed(eng, lambda: eng.scalar_mul_fixed_vartime(gen_m, t, out=fo), reps=3)
    out["fixed_base"] = {"n": n, "ms": ms, "per_s": n / ms * 1e3}
    print(f"fixed-base n={n}: {ms:.2f} ms  {n / ms * 1e3:.3e}/s", flush=True)
    pts = fo
    k = eng.fe_to_bytes("fr", eng.fe_stream("fr", SEED0 + 2, n, device=True))
    o = eng.empty((n, 20))
    out["variants"] = {}
    for v in (11, 13, 21):
        eng.set_scalar_mul_variant(v)
        try:
            ms = timed(eng, lambda: eng.scalar_mul_vartime(pts, k, out=o, flags=jj.JJ_ASYNC), reps=2)
            out["variants"][v] = {"ms": ms, "per_s": n / ms * 1e3}
            print(f"variant {v}: {ms:.2f} ms  {n / ms * 1e3:.3e} scalar-mul/s", flush=True)
        except Exception as e:  # noqa: BLE001
            out["variants"][v] = {"error": str(e)}
            print(f"variant {v}: {e}", flush=True)
    eng.set_scalar_mul_variant(0)
    ao = eng.empty((n, 8))
    ms = timed(eng, lambda: eng.batch_normalize(o, out=ao), reps=3)
    out["batch_normalize"] = {"n": n, "ms": ms, "per_s": n / ms * 1e3}
    print(f"batch_normalize n={n}: {ms:.2f} ms", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
