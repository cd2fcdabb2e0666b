"""e2e time of the wire-format path (jj_scalar_mul_encoded, pinned host buffers) for the chunking chosen by JJ_WIRE_ROUNDS."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import jubjub_b200 as jj  # noqa: E402
from bench import SEED0, make_inputs, pinned  # noqa: E402

eng = jj.Engine(0)
n = 1 << 20
pts, k = make_inputs(eng, n, 0)
enc = eng.affine_to_bytes(eng.batch_normalize(pts))
henc, _ = pinned(eng, (n, 32), np.uint8)
hk, _ = pinned(eng, (n, 32), np.uint8)
hout, _ = pinned(eng, (n, 32), np.uint8)
hok, _ = pinned(eng, (n,), np.uint8)
henc[:], hk[:] = enc.download(), k.download()


def step():
    eng._check(eng.lib.jj_scalar_mul_encoded(eng.ctx, henc.ctypes.data, hk.ctypes.data, hout.ctypes.data, hok.ctypes.data, n,
                                             jj.JJ_OUT_BYTES))


step()
step()
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    step()
    ts.append((time.perf_counter() - t0) * 1e3)
print(f"JJ_WIRE_ROUNDS={os.environ.get('JJ_WIRE_ROUNDS', 'default')}: best {min(ts):.3f} ms, median {sorted(ts)[2]:.3f} ms, ok={int(hok.sum())}")
