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


def edge_field_values(m):
    """Adversarial field values (all < m) for the limb-level code paths of the kernels: the squaring's
    fold boundary (top limb around m7 / 2), limbs of 0 / 0xffffffff (quotient digit 0, carries out of every
    column), single-bit and sparse values, and the neighbours of 0, m / 2 and m."""
    top = m >> 224
    half_top = (m >> 225) << 224  # top limb = m7 >> 1, other limbs 0
    vals = [0, 1, 2, 3, m - 1, m - 2, m - 3, m >> 1, (m >> 1) + 1, (m >> 1) - 1, (m + 1) >> 1,
            M.to_mont(1, m), M.to_mont(m - 1, m), M.to_mont(2, m), (1 << 255) % m,
            half_top, half_top + (1 << 224) - 1, half_top + (1 << 224), half_top + (1 << 224) + 1, half_top - 1,
            (top << 224) - 1, top << 224, (1 << 224) - 1, (1 << 192) - 1, (1 << 128) - 1, (1 << 96) - 1, (1 << 64) - 1,
            (1 << 32) - 1, m - (1 << 32), m - (1 << 224), (m >> 32) << 32, (m >> 96) << 96, (m >> 224) << 224]
    vals += [1 << k for k in (31, 32, 33, 63, 64, 95, 96, 127, 128, 160, 191, 192, 223, 224, 250, 251, 252, 253, 254)
             if (1 << k) < m]
    vals += [((1 << 256) - 1 - (0xFFFFFFFF << (32 * k))) % m for k in range(8)]       # one zero limb
    vals += [(0xFFFFFFFF << (32 * k)) % m for k in range(8)]                            # one all-ones limb
    vals += [int("0000000100000000" * 4, 16) % m, int("00000000ffffffff" * 4, 16) % m, int("ffffffff00000000" * 4, 16) % m]
    vals = [v % m for v in vals]
    return np.array([M.limbs(v) for v in vals], dtype=np.uint64)


def edge_field_pairs(m):
    """All ordered pairs of edge_field_values(m): (a, b) arrays."""
    e = edge_field_values(m)
    i, j = np.meshgrid(np.arange(len(e)), np.arange(len(e)), indexing="ij")
    return e[i.ravel()], e[j.ravel()]
