"""The C++ host mirror (include/jubjub_b200.hpp) against oracle-derived expectations, on the GPU."""
import os
import subprocess

import numpy as np
import pytest

from oracle import model as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_host_mirror(oracle, tmp_path):
    exe = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")
    if not os.path.exists(exe):
        import __graft_entry__ as g

        g.build()
    n = 257
    g = oracle.affine_to_extended(oracle.generator())
    t = oracle.fe_to_bytes(1, oracle.fe_stream(1, M.SEED0 + 3, n))
    p = oracle.ext_double(oracle.scalar_mul(np.repeat(g, n, axis=0), t))
    k = oracle.fe_stream(1, M.SEED0 + 2, n)  # Montgomery-form Fr, as the reference's Fr holds it
    want = oracle.affine_to_bytes(oracle.batch_normalize(oracle.scalar_mul(p, oracle.fe_to_bytes(1, k))))
    path = tmp_path / "case.bin"
    with open(path, "wb") as f:
        f.write(np.uint64(n).tobytes())
        f.write(p.tobytes())
        f.write(k.tobytes())
        f.write(want.tobytes())
    r = subprocess.run([exe, str(path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "cpp mirror ok" in r.stdout
