"""Python big-integer model of the Jubjub hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *mathematical* cross-check for the C oracle (oracle/jj_oracle.c)
and, through it, for the CUDA engine.  It restates values, not algorithms: field
elements are Python ints, points are affine (u, v) pairs, the group law is the
affine twisted-Edwards addition law.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import it; the product (jubjub_b200/) never does.

Reference anchors (relative to /root/reference):
  * r, Fr Montgomery constants ............. src/fr.rs:77-82, 214, 217-238
  * q (= bls12_381::Scalar modulus, [ext]) .. q-1 appears at src/lib.rs:1629-1634
  * d = -(10240/10241), 2d ................. src/lib.rs:399-412
  * generator (u, v=11) .................... src/lib.rs:1380-1396
  * encoding ............................... src/lib.rs:455-464, 492-534
  * scalar-mul ignores bits 252..255 ....... src/lib.rs:356-379
  * from_bytes_wide = d0*R2 + d1*R3 ........ src/fr.rs:312-343
"""

MASK64 = (1 << 64) - 1

# --- moduli -----------------------------------------------------------------
Q = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
R_ORDER = 0x0E7DB4EA6533AFA906673B0101343B00A6682093CCC81082D0970E5ED6F72CB7

MONT_R = 1 << 256  # Montgomery radix for both fields (4 x u64 limbs)

# --- curve ------------------------------------------------------------------
D = (-10240 * pow(10241, -1, Q)) % Q
D2 = (2 * D) % Q
GEN_U = 0x62EDCBB8BF3787C88B0F03DDD60A8187CAF55D1B29BF81AFE4B3D35DF1A7ADFE
GEN_V = 11
FR_MODULUS_BYTES = R_ORDER.to_bytes(32, "little")


def limbs(x):
    """int -> 4 little-endian u64 limbs."""
    return [(x >> (64 * i)) & MASK64 for i in range(4)]


def from_limbs(l):
    return sum(int(v) << (64 * i) for i, v in enumerate(l))


def to_mont(x, m):
    return (x * MONT_R) % m


def from_mont(x, m):
    return (x * pow(MONT_R, -1, m)) % m


def mont_mul(a, b, m):
    """Montgomery product of two Montgomery-form residues (src/fr.rs:592-616)."""
    return (a * b * pow(MONT_R, -1, m)) % m


def from_bytes_wide(b64, m):
    """Value (not Montgomery form) of a 512-bit LE integer mod m (src/fr.rs:312-343)."""
    assert len(b64) == 64
    return int.from_bytes(b64, "little") % m


# --- SplitMix64 input streams -------------------------
SEED0 = 0x4A55424A55420001


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & MASK64

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)


def stream_wide_bytes(seed, n):
    """n elements, each 8 consecutive SplitMix64 outputs as 64 LE bytes."""
    g = SplitMix64(seed)
    out = []
    for _ in range(n):
        out.append(b"".join(g.next().to_bytes(8, "little") for _ in range(8)))
    return out


def stream_field(seed, n, m):
    return [from_bytes_wide(b, m) for b in stream_wide_bytes(seed, n)]


# --- affine twisted Edwards group law: -u^2 + v^2 = 1 + d u^2 v^2 ------------
IDENTITY = (0, 1)


def on_curve(p):
    u, v = p
    return (v * v - u * u - 1 - D * u * u % Q * v * v) % Q == 0


def padd(p1, p2):
    u1, v1 = p1
    u2, v2 = p2
    t = D * u1 % Q * u2 % Q * v1 % Q * v2 % Q
    u3 = (u1 * v2 + v1 * u2) * pow(1 + t, -1, Q) % Q
    v3 = (v1 * v2 + u1 * u2) * pow(1 - t, -1, Q) % Q
    return (u3, v3)


def pneg(p):
    return ((-p[0]) % Q, p[1])


def pmul(p, k):
    """[k]P with plain double-and-add over the integer k."""
    acc = IDENTITY
    for i in reversed(range(k.bit_length())):
        acc = padd(acc, acc)
        if (k >> i) & 1:
            acc = padd(acc, p)
    return acc


def pmul_fast(p, k):
    """[k]P, same value as pmul, without a modular inversion per step: projective (X : Y : Z) twisted-Edwards
    arithmetic with the unified addition law (add-2008-bbjlp, a = -1), one inversion at the end.  Independent of the
    C oracle and of the kernels (they use extended coordinates with Niels tables)."""
    if k == 0:
        return IDENTITY
    x1, y1 = p
    X, Y, Z = 0, 1, 1
    for i in reversed(range(k.bit_length())):
        # doubling (dbl-2008-bbjlp, a = -1)
        B = (X + Y) * (X + Y) % Q
        C, Dd = X * X % Q, Y * Y % Q
        E = (-C) % Q
        F = (E + Dd) % Q
        H = Z * Z % Q
        J = (F - 2 * H) % Q
        X, Y, Z = (B - C - Dd) * J % Q, F * (E - Dd) % Q, F * J % Q
        if (k >> i) & 1:  # mixed addition with the affine (x1, y1)
            C, Dd = X * x1 % Q, Y * y1 % Q
            E = D * C % Q * Dd % Q
            B = Z * Z % Q
            F, G = (B - E) % Q, (B + E) % Q
            X3 = Z * F % Q * ((X + Y) * (x1 + y1) - C - Dd) % Q
            Y3 = Z * G % Q * (Dd + C) % Q  # Dd - a*C with a = -1
            X, Y, Z = X3, Y3, F * G % Q
    zi = pow(Z, -1, Q)
    return (X * zi % Q, Y * zi % Q)


def scalar_from_bytes_ref(b32):
    """The integer the reference's multiply() actually uses: low 252 bits
    (src/lib.rs:363-372: MSB-first bits, first 4 skipped)."""
    return int.from_bytes(b32, "little") & ((1 << 252) - 1)


def encode(p):
    """32-byte compressed encoding (src/lib.rs:455-464)."""
    u, v = p
    b = bytearray(v.to_bytes(32, "little"))
    b[31] |= (u & 1) << 7
    return bytes(b)


def fq_sqrt(a):
    """Any square root of a mod q, or None (Tonelli-Shanks; q-1 = 2^32 * t)."""
    a %= Q
    if a == 0:
        return 0
    if pow(a, (Q - 1) // 2, Q) != 1:
        return None
    s, t = 32, (Q - 1) >> 32
    z = pow(7, t, Q)  # 7 generates the 2-Sylow subgroup ([ext] bls12_381 GENERATOR)
    x = pow(a, (t + 1) // 2, Q)
    b = pow(a, t, Q)
    m = s
    while b != 1:
        i, bb = 0, b
        while bb != 1:
            bb = bb * bb % Q
            i += 1
        w = pow(z, 1 << (m - i - 1), Q)
        x = x * w % Q
        z = w * w % Q
        b = b * z % Q
        m = i
    return x


def decode(b32, zip216=True):
    """AffinePoint::from_bytes_inner (src/lib.rs:492-534). Returns (u, v) or None."""
    b = bytearray(b32)
    sign = b[31] >> 7
    b[31] &= 0x7F
    v = int.from_bytes(b, "little")
    if v >= Q:
        return None
    v2 = v * v % Q
    den = (1 + D * v2) % Q
    u2 = (v2 - 1) * pow(den, -1, Q) % Q if den else 0
    u = fq_sqrt(u2)
    if u is None:
        return None
    flip = (u & 1) ^ sign
    final_u = (-u) % Q if flip else u
    if zip216 and u == 0 and flip:
        return None
    return (final_u, v)
