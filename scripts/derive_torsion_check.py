"""Derivation and exhaustive-coset check of the pairing-based subgroup test used by jj_is_torsion_free.

The reference decides `is_torsion_free` as [r]P == O (src/lib.rs:709-711): a 252-step ladder.  Only the
boolean has to match, so the engine uses the order-8 reduced Tate pairing instead:

  E(Fq) is cyclic of order 8r (one rational point of order 2, (0, -1), and a rational point T of order 8,
  src/lib.rs:1589-1677), mu_8 lies in Fq (q = 1 mod 2^32), so
      t(T, P) = f_{8,T}(P) ^ ((q-1)/8)
  is a non-degenerate character of E/8E = Z/8 : it is 1 exactly for P in 8E = the subgroup of order r.

Miller's algorithm for the fixed point T on the birationally equivalent Montgomery curve
  B y^2 = x^3 + A x^2 + x,  x = (1+v)/(1-v),  y = x/u     (a = -1: A = 2(a+d)/(a-d), B = 4/(a-d))
gives  f_8 = l1^4 l2^2 / (v2^4 v4)  with  l1, l2 the tangents at T and 2T, v2, v4 the verticals at 2T and 4T = (0,0).
Modulo 8th powers  f_8 = l1^4 l2^2 v2^4 v4^7, and with the projective substitution
  x = Nx / D, y = Ny / D,  Nx = (Z+V) U, Ny = (Z+V) Z, D = (Z-V) U
every factor is linear over the common denominator D; 17 copies of 1/D are D^7 modulo 8th powers:
  F(P) = L1^4 L2^2 V2^4 Nx^7 D^7,   L_k = Ny - lam_k Nx + c_k D,   V2 = Nx - x2 D.
P is torsion free  <=>  U == 0 ? (V == Z) : F(P)^((q-1)/8) == C   (C: the value on the subgroup, calibrated below;
it absorbs the normalisation of f at infinity).

Run:  python scripts/derive_torsion_check.py   -> prints the constants (plain integers mod q) used by
jubjub_b200/csrc/torsion.cuh and checks the criterion on every coset P + jT, j = 0..7, of random subgroup points,
on all eight small-order points and on random full-order points.
"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model as M  # noqa: E402

Q, D = M.Q, M.D
inv = lambda x: pow(x % Q, -1, Q)  # noqa: E731


def find_order8_point():
    """A point of exact order 8: [r] of a point whose 8-part is full."""
    rng = random.Random(1)
    while True:
        v = rng.randrange(Q)
        p = M.decode(v.to_bytes(32, "little"))
        if p is None:
            continue
        t = M.pmul(p, M.R_ORDER)
        if M.pmul(t, 4) != M.IDENTITY:
            return t


def to_mont_xy(p):
    u, v = p
    x = (1 + v) * inv(1 - v) % Q
    return x, x * inv(u) % Q


def constants():
    a = Q - 1
    A = 2 * (a + D) * inv(a - D) % Q
    B = 4 * inv(a - D) % Q
    T = find_order8_point()
    T2 = M.padd(T, T)
    assert M.pmul(T, 4) == (0, Q - 1)
    x1, y1 = to_mont_xy(T)
    x2, y2 = to_mont_xy(T2)
    for x, y in ((x1, y1), (x2, y2)):
        assert (B * y * y - (x * x * x + A * x * x + x)) % Q == 0
    lam1 = (3 * x1 * x1 + 2 * A * x1 + 1) * inv(2 * B * y1) % Q
    lam2 = (3 * x2 * x2 + 2 * A * x2 + 1) * inv(2 * B * y2) % Q
    c1 = (lam1 * x1 - y1) % Q
    c2 = (lam2 * x2 - y2) % Q
    return {"T": T, "lam1": lam1, "c1": c1, "lam2": lam2, "c2": c2, "x2": x2}


def F_miller(p_proj, k):
    """The Miller value as derived: L1^4 L2^2 V2^4 Nx^7 D^7."""
    U, V, Z = p_proj
    Nx = (Z + V) * U % Q
    Ny = (Z + V) * Z % Q
    Dn = (Z - V) * U % Q
    L1 = (Ny - k["lam1"] * Nx + k["c1"] * Dn) % Q
    L2 = (Ny - k["lam2"] * Nx + k["c2"] * Dn) % Q
    V2 = (Nx - k["x2"] * Dn) % Q
    return pow(L1, 4, Q) * pow(L2, 2, Q) * pow(V2, 4, Q) * pow(Nx, 7, Q) * pow(Dn, 7, Q) % Q


def F(p_proj, k):
    """The same class modulo 8th powers and constants, as the kernel evaluates it (12 M + 4 S).
    On Jubjub x(2T) = 1 and c2 = 0, so V2 = Nx - D = 2UV and L2 = (Z+V)(Z - lam2 U); with a = Z+V, b = Z-V
      F = L1^4 (Z - lam2 U)^2 U^2 V^4 a b^7 * (8th powers)  =  (G b^3)^2 a b,   G = L1^2 (Z - lam2 U) U V^2."""
    U, V, Z = p_proj
    a, b = (Z + V) % Q, (Z - V) % Q
    aU, aZ, bU = a * U % Q, a * Z % Q, b * U % Q
    L1 = (aZ - k["lam1"] * aU + k["c1"] * bU) % Q
    G = L1 * L1 % Q * ((Z - k["lam2"] * U) % Q) % Q * (U * V % Q * V % Q) % Q
    H = G * (b * b % Q * b % Q) % Q
    return H * H % Q * (a * b % Q) % Q


def is_torsion_free(p_proj, k, C):
    U, V, Z = p_proj
    if U % Q == 0:
        return (V - Z) % Q == 0
    return pow(F(p_proj, k), (Q - 1) // 8, Q) == C


def main():
    k = constants()
    rng = random.Random(7)
    G = (M.GEN_U, M.GEN_V)
    G8 = M.pmul(G, 8)  # generator of the prime-order subgroup (src/lib.rs:1811: 8*G)
    # calibrate C on one subgroup point, then check everything else against it
    C = pow(F((G8[0], G8[1], 1), k), (Q - 1) // 8, Q)
    assert pow(C, 8, Q) == 1
    T = k["T"]
    small = [M.pmul(T, j) for j in range(8)]
    checked = 0
    for _ in range(200):
        s = rng.randrange(1, M.R_ORDER)
        P = M.pmul(G8, s)
        z = rng.randrange(1, Q)
        for j in range(8):
            Pj = M.padd(P, small[j])
            want = M.pmul(Pj, M.R_ORDER) == M.IDENTITY
            assert want == (j == 0)
            got = is_torsion_free((Pj[0] * z % Q, Pj[1] * z % Q, z), k, C)
            assert got == want, (j, s)
            checked += 1
    for j in range(8):  # the small-order points themselves (F = 0 or U = 0 there)
        z = rng.randrange(1, Q)
        Pj = small[j]
        assert is_torsion_free((Pj[0] * z % Q, Pj[1] * z % Q, z), k, C) == (j == 0), j
    ratio = set()
    for _ in range(64):  # the simplified form differs from the Miller value by a constant 8th-power class only
        P = M.pmul(G, rng.randrange(1, 8 * M.R_ORDER))
        z = rng.randrange(1, Q)
        pp = (P[0] * z % Q, P[1] * z % Q, z)
        ratio.add(pow(F(pp, k) * inv(F_miller(pp, k)), (Q - 1) // 8, Q))
    assert len(ratio) == 1
    chars = set()
    for j in range(8):  # the character is injective on E/8E
        Pj = M.padd(G8, small[j]) if j else G8
        chars.add(pow(F((Pj[0], Pj[1], 1), k), (Q - 1) // 8, Q))
    assert len(chars) == 8
    for _ in range(200):  # random points of the whole group
        P = M.pmul(G, rng.randrange(1, 8 * M.R_ORDER))
        want = M.pmul(P, M.R_ORDER) == M.IDENTITY
        assert is_torsion_free((P[0], P[1], 1), k, C) == want
    print(f"criterion verified on {checked} coset points, 8 small-order points, 200 random points")
    for name in ("lam1", "c1", "lam2", "c2", "x2"):
        print(f"{name:5s} = 0x{k[name]:064x}")
    print(f"C     = 0x{C:064x}")
    print(f"T     = (0x{T[0]:064x}, 0x{T[1]:064x})")

    def words(x):
        x = M.to_mont(x, Q)
        return ", ".join(f"0x{(x >> (32 * i)) & 0xFFFFFFFF:08x}u" for i in range(8))

    print("\n// Montgomery-form 32-bit limbs for torsion.cuh")
    for name in ("lam1", "c1", "lam2", "c2", "x2"):
        print(f"{name.upper():5s}: {{{words(k[name])}}}")
    print(f"C    : {{{words(C)}}}")
    e = (Q - 1) >> 32  # T: F^T then 29 squarings
    print("T (q-1 = 2^32 T) words:", ", ".join(f"0x{(e >> (32 * i)) & 0xFFFFFFFF:08x}u" for i in range(7)))


if __name__ == "__main__":
    main()
