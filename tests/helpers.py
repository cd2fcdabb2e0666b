"""Small numpy helpers shared by the tests."""
import numpy as np

from oracle import model as M


def fe(*limb_lists):
    """limb lists -> (n, 4) uint64."""
    return np.array(limb_lists, dtype=np.uint64).reshape(-1, 4)


def fe_int(x):
    return np.array(M.limbs(x), dtype=np.uint64).reshape(1, 4)


def to_int(row):
    return M.from_limbs(row)


def b32(*byte_lists):
    return np.array(byte_lists, dtype=np.uint8).reshape(-1, 32)


def scalar_bytes(*ints):
    return np.frombuffer(b"".join(int(k).to_bytes(32, "little") for k in ints), dtype=np.uint8).reshape(-1, 32).copy()


def affine_raw(oracle, pts_raw):
    """[(u_raw, v_raw), ...] -> (n, 8) Montgomery affine, via the oracle's from_raw."""
    u = oracle.fe_from_raw(oracle.FQ, fe(*[p[0] for p in pts_raw]))
    v = oracle.fe_from_raw(oracle.FQ, fe(*[p[1] for p in pts_raw]))
    return np.concatenate([u, v], axis=1)


def affine_values(oracle, aff):
    """(n, 8) Montgomery affine -> list of (u, v) Python ints (canonical values)."""
    ub = oracle.fe_to_bytes(oracle.FQ, aff[:, :4])
    vb = oracle.fe_to_bytes(oracle.FQ, aff[:, 4:])
    return [(int.from_bytes(bytes(ub[i]), "little"), int.from_bytes(bytes(vb[i]), "little"))
            for i in range(len(aff))]


def affine_from_values(pts):
    """[(u, v) ints] -> (n, 8) Montgomery affine (pure bigint conversion)."""
    rows = [M.limbs(M.to_mont(u, M.Q)) + M.limbs(M.to_mont(v, M.Q)) for u, v in pts]
    return np.array(rows, dtype=np.uint64).reshape(-1, 8)


def extended_from_values(pts):
    """[(u, v) ints] -> (n, 20) extended with z = 1, t1 = u, t2 = v."""
    one = M.limbs(M.to_mont(1, M.Q))
    rows = []
    for u, v in pts:
        lu, lv = M.limbs(M.to_mont(u, M.Q)), M.limbs(M.to_mont(v, M.Q))
        rows.append(lu + lv + one + lu + lv)
    return np.array(rows, dtype=np.uint64).reshape(-1, 20)
