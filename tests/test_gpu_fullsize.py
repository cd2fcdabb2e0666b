"""Parity at the BASELINE.json sizes, every unit against the oracle:
  config 2  2^20 Fq and Fr mul / square / add / sub, all limbs of all units vs the C oracle
            (Fr::mul src/fr.rs:592-616, square :353-381, add :638-647, sub :620-634; Fq = [ext] same shape);
  config 3  all 2^20 variable-base results, normalised, vs the oracle's bitwise ladder
            (src/lib.rs:356-379; ~14 s on 16 host threads) and 4 096 of them vs the big-integer model;
  config 4  all 2^20 fixed-base results vs the oracle's AffineNielsPoint::multiply (src/lib.rs:271-295).
Inputs are the SplitMix64 streams of the bench (same seeds), so this is the benchmark's own batch."""
import os

import numpy as np
import pytest

from oracle import model as M

pytestmark = pytest.mark.gpu
FQ, FR = 0, 1
N = 1 << 20


@pytest.fixture(scope="module")
def eng():
    import jubjub_b200 as jj

    e = jj.Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("which,name", [(FQ, "fq"), (FR, "fr")])
def test_config2_field_ops_all_1m(eng, oracle, which, name):
    a = oracle.fe_stream(which, M.SEED0, N)
    b = oracle.fe_stream(which, M.SEED0 + 1, N)
    # the device generates the same streams (this is how bench.py makes its inputs)
    assert (eng.fe_stream(name, M.SEED0, N) == a).all()
    da, db = eng.to_device(a), eng.to_device(b)
    for op, fn in ((oracle.OP_MUL, eng.fe_mul), (oracle.OP_ADD, eng.fe_add), (oracle.OP_SUB, eng.fe_sub)):
        want = oracle.fe_batch(which, op, a, b)
        assert (fn(name, a, b) == want).all(), op                      # host buffers (staged chunks)
        assert (fn(name, da, db).download() == want).all(), op         # device resident
    want = oracle.fe_batch(which, oracle.OP_SQUARE, a)
    assert (eng.fe_square(name, a) == want).all() and (eng.fe_square(name, da).download() == want).all()
    # values, not only agreement with the C oracle: 4 096 products against Python big integers
    m = M.Q if which == FQ else M.R_ORDER
    rinv = pow(1 << 256, -1, m)
    got = eng.fe_mul(name, a, b)
    for i in range(0, N, N // 4096):
        assert M.from_limbs(got[i]) == M.from_limbs(a[i]) * M.from_limbs(b[i]) * rinv % m, i


def _bench_inputs(eng, oracle, n):
    t = oracle.fe_to_bytes(FR, oracle.fe_stream(FR, M.SEED0 + 3, n))
    k = oracle.fe_to_bytes(FR, oracle.fe_stream(FR, M.SEED0 + 2, n))
    return t, k


def test_config3_variable_base_all_1m(eng, oracle):
    """Every one of the 2^20 results of the headline workload (P_i = [t_i]G, k_i: the bench's streams) equals the
    oracle's reference ladder after normalisation -- in all three output formats -- and 4 096 of them equal the
    big-integer model's affine double-and-add."""
    threads = os.cpu_count() or 1
    t, k = _bench_inputs(eng, oracle, N)
    pts = eng.scalar_mul_fixed_vartime(oracle.generator(), t)   # checked unit by unit in test_config4 below
    want_aff = oracle.batch_normalize(oracle.scalar_mul(pts, k, threads))
    dp, dk = eng.to_device(pts), eng.to_device(k)
    got_ext = eng.scalar_mul_vartime(dp, dk).download()
    assert (eng.batch_normalize(got_ext) == want_aff).all()
    assert (oracle.batch_normalize(got_ext[::64]) == want_aff[::64]).all()   # the oracle's own normalisation too
    want_enc = oracle.affine_to_bytes(want_aff)
    assert (eng.scalar_mul_vartime(dp, dk, output="affine").download() == want_aff).all()   # fused normalise epilogue
    assert (eng.scalar_mul_vartime(dp, dk, output="bytes").download() == want_enc).all()
    assert (eng.scalar_mul_vartime(pts, k, output="bytes") == want_enc).all()              # host buffers, staged chunks
    # wire-format entry: encodings in, encodings out
    enc_in = oracle.affine_to_bytes(oracle.batch_normalize(pts))
    got_w, ok = eng.scalar_mul_encoded_vartime(enc_in, k)
    assert ok.all() and (got_w == want_enc).all()
    # big-integer model on 4 096 units: values of P_i from the encodings, [k_i]P_i by affine arithmetic
    G = (M.GEN_U, M.GEN_V)
    for i in range(0, N, N // 4096):
        P = M.pmul_fast(G, int.from_bytes(bytes(t[i]), "little"))
        assert M.encode(P) == bytes(enc_in[i]), i
        R = M.pmul_fast(P, M.scalar_from_bytes_ref(bytes(k[i])))
        assert M.encode(R) == bytes(want_enc[i]), i


def test_config4_fixed_base_all_1m(eng, oracle):
    threads = os.cpu_count() or 1
    _, k = _bench_inputs(eng, oracle, N)
    want = oracle.batch_normalize(oracle.scalar_mul_fixed(oracle.generator(), k, threads))
    assert (eng.scalar_mul_fixed_vartime(oracle.generator(), k, output="affine") == want).all()
    dk = eng.to_device(k)
    got = eng.scalar_mul_fixed_vartime(oracle.generator(), dk).download()
    assert (eng.batch_normalize(got) == want).all()
    assert (eng.scalar_mul_fixed_vartime(oracle.generator(), dk, output="bytes").download() == oracle.affine_to_bytes(want)).all()
