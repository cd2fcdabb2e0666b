// scalarmul.cuh -- variable-base and fixed-base scalar multiplication cores.
//
// Replaces ExtendedNielsPoint::multiply / ExtendedPoint::multiply / AffineNielsPoint::multiply
// (src/lib.rs:356-379, 830-833, 271-295).  The reference walks the scalar one bit at a time
// (252 doublings + 252 constant-time additions of "P or identity").  Here one thread owns
// one scalar-mul and uses a signed radix-16 window:
//
//   k (low 252 bits, exactly the bits the reference consumes, src/lib.rs:366-372)
//     = T * 16^62 + sum_{i<62} d_i * 16^i,   d_i in [-8, 7],  T in [0, 16]
//
// obtained by adding 0x0088...8 to k and reading nibbles (nibble - 8) below the top part T.  A per-scalar-mul
// table holds the eight extended-Niels multiples 1P..8P; negative digits use the
// reference's own subtraction formula (src/lib.rs:922-940), so no field negation is
// needed.  Cost: 7 additions + 8 to_niels for the table, then 252 doublings and <= 63
// additions: ~2.3k field multiplications instead of the reference ladder's 3.8k.  The
// projective output differs from the reference's ladder; the affine point (and its
// 32-byte encoding) is identical -- parity is checked after normalisation.
//
// This batch engine is variable-time in the scalar (zero digits skip their addition);
// it is for public scalars unless the caller accepts that (the reference's constant-time
// policy is src/lib.rs:12-17; the Rust shim names these entry points *_vartime).
#pragma once
#include "point.cuh"

namespace jj {

// t = (k mod 2^252) + 0x0088...8 (an 8 in each of the nibbles 0..61).  Nibble i < 62 of t, minus 8, is the signed digit d_i;
// the top byte of t is T = (nibble 62 of k) + carry in [0, 16].  Returns T and K = t << 8, whose top nibble is digit 61.
JJ_DEVICE uint32_t recode_scalar(uint32_t K[8], const uint32_t k[8]) {
    uint32_t t[8];
    add_cc(t[0], k[0], 0x88888888u);
#pragma unroll
    for (int i = 1; i < 7; i++) addc_cc(t[i], k[i], 0x88888888u);
    addc(t[7], k[7] & 0x0fffffffu, 0x00888888u);
#pragma unroll
    for (int i = 7; i > 0; i--) K[i] = (t[i] << 8) | (t[i - 1] >> 24);
    K[0] = t[0] << 8;
    return t[7] >> 24;
}
JJ_DEVICE void shl_256(uint32_t K[8], int s) {  // 0 < s < 32
#pragma unroll
    for (int i = 7; i > 0; i--) K[i] = (K[i] << s) | (K[i - 1] >> (32 - s));
    K[0] <<= s;
}

// Table policy used by the host emulation and as the plain fallback layout: a private array.
// load_signed(j, neg, n): entry j with its v+u / v-u halves exchanged when neg -- i.e. the Niels form of -(j+1)P up to
// the sign of t2d, which the addition's output side handles (point_add_niels_preswapped).
struct LocalTable {
    ext_niels t[8];
    JJ_DEVICE_SPEC void store(int j, const ext_niels& n) { t[j] = n; }
    JJ_DEVICE_SPEC void load(int j, ext_niels& n) const { n = t[j]; }
    JJ_DEVICE_SPEC void load_signed(int j, bool neg, ext_niels& n) const {
        n = t[j];
        if (neg) {
            n.vpu = t[j].vmu;
            n.vmu = t[j].vpu;
        }
    }
};

#ifndef JJ_DBL_UNROLL
#define JJ_DBL_UNROLL 1
#endif
constexpr int kDoubleUnroll = JJ_DBL_UNROLL;  // unroll factor of the 4-doubling loop (code size vs register moves)

// Branch-free, address-uniform lookup for the constant-time mode: every entry is read (all lanes read the same address, so
// the accesses are fully coalesced) and the wanted one is kept by selects; digit 0 yields the Niels identity (1, 1, 1, 0),
// whose addition leaves the point unchanged (src/lib.rs:347-354), so the addition is always executed.
template <class Table>
JJ_DEVICE void table_scan(ext_niels& n, const Table& tbl, int mag) {  // mag = |digit| in [0, 8]
    fe_set_one<FqP>(n.vpu);
    fe_set_one<FqP>(n.vmu);
    fe_set_one<FqP>(n.z);
    fe_set_zero(n.t2d);
#pragma unroll 1
    for (int j = 0; j < 8; j++) {
        ext_niels e;
        tbl.load(j, e);
        const bool hit = mag == j + 1;
        fe_select(n.vpu, n.vpu, e.vpu, hit);
        fe_select(n.vmu, n.vmu, e.vmu, hit);
        fe_select(n.z, n.z, e.z, hit);
        fe_select(n.t2d, n.t2d, e.t2d, hit);
    }
}

// acc = [k] P.  `tbl` provides storage for the 8-entry window table.
// CT = false (the *_vartime entry points): a zero digit skips its addition and the table is indexed by the digit.
// CT = true (JJ_CONST_TIME): no branch and no address depends on the scalar -- the table is scanned, the sign is applied by
// selects and the addition always runs -- the batch analogue of the reference's constant-time ladder, which always adds
// "P or identity" (src/lib.rs:356-379).  Same values either way.
template <class Table, bool CT = false>
JJ_DEVICE void scalar_mul_core(ext_point& acc, const ext_point& P, const uint32_t k[8], Table& tbl) {
    {
        ext_niels n1, nj;
        ext_point cur = P;
        point_to_niels(n1, P);
        tbl.store(0, n1);
#pragma unroll 1
        for (int j = 1; j < 8; j++) {
            point_add_niels(cur, cur, n1, false);
            point_to_niels(nj, cur);
            tbl.store(j, nj);
        }
    }
    uint32_t K[8];
    const uint32_t T = recode_scalar(K, k);
    // k = T * 16^62 + sum_{i < 62} d_i 16^i.  The top part T = 2h + b, h in [0, 8]: start from [h]P read back from the table
    // (Niels (v+u, v-u, z) -> (2u : 2v : 2z), the same point), and let the first pass of the loop be ONE doubling and the
    // addition of [b]P instead of four doublings of "P or identity" and the addition of digit 62: three doublings fewer per
    // scalar multiplication.
    {
        ext_niels n;
        const int h = (int)(T >> 1);
        if (CT) {
            table_scan(n, tbl, h);
        } else {
            fe_set_one<FqP>(n.vpu);
            fe_set_one<FqP>(n.vmu);
            fe_set_one<FqP>(n.z);
            if (h != 0) tbl.load(h - 1, n);
        }
        fe_sub<FqP>(acc.u, n.vpu, n.vmu);
        fe_add<FqP>(acc.v, n.vpu, n.vmu);
        fe_dbl<FqP>(acc.z, n.z);
        acc.t1 = acc.u;  // not read before the next addition: every doubling recomputes t1, t2
        acc.t2 = acc.v;
    }
#pragma unroll 1
    for (int i = 62; i >= 0; i--) {
        const int nd = i == 62 ? 1 : 4;
#pragma unroll kDoubleUnroll
        for (int j = 0; j < nd; j++) point_double(acc, acc);
        int d;
        if (i == 62) {
            d = (int)(T & 1u);
        } else {
            d = (int)(K[7] >> 28) - 8;
            shl_256(K, 4);
        }
        if (CT) {
            ext_niels n;
            table_scan(n, tbl, d < 0 ? -d : d);
            point_add_niels(acc, acc, n, d < 0);  // operand and output halves chosen by selects
        } else if (d != 0) {
            ext_niels n;
            tbl.load_signed((d < 0 ? -d : d) - 1, d < 0, n);
            point_add_niels_preswapped(acc, acc, n, d < 0);
        }
    }
}

// ---- one scalar shared by the whole batch (is_torsion_free: the scalar is r) --------------------------
// Every lane has the same digits, so a width-5 sliding-window NAF costs no divergence: digits are odd, in
// [-15, 15], at most one in any 5 consecutive positions -- for r, 42 additions instead of the 58 non-zero
// signed radix-16 digits.  The table holds the odd multiples 1P, 3P, ..., 15P (one doubling + 7 additions).
struct NafDigits {
    int8_t d[256];  // little-endian: k = sum d[i] 2^i
    int top;        // index of the highest non-zero digit, -1 for k = 0
};
// Host side: width-5 NAF of the low 252 bits of a 256-bit little-endian scalar (the bits `multiply` consumes,
// src/lib.rs:366-372).
inline void wnaf5_recode(NafDigits& out, const uint32_t k_in[8]) {
    uint32_t k[9];
    for (int i = 0; i < 8; i++) k[i] = k_in[i];
    k[7] &= 0x0fffffffu;
    k[8] = 0;
    out.top = -1;
    for (int i = 0; i < 256; i++) {
        int t = 0;
        if (k[0] & 1u) {
            t = (int)(k[0] & 31u);
            if (t >= 16) t -= 32;
            // k -= t
            uint64_t carry = 0;
            if (t > 0) {
                uint64_t borrow = (uint64_t)t;
                for (int w = 0; w < 9 && borrow; w++) {
                    uint64_t v = (uint64_t)k[w];
                    k[w] = (uint32_t)(v - borrow);
                    borrow = v < borrow ? 1 : 0;
                }
            } else {
                carry = (uint64_t)(-t);
                for (int w = 0; w < 9 && carry; w++) {
                    uint64_t v = (uint64_t)k[w] + carry;
                    k[w] = (uint32_t)v;
                    carry = v >> 32;
                }
            }
            out.top = i;
        }
        out.d[i] = (int8_t)t;
        for (int w = 0; w < 8; w++) k[w] = (k[w] >> 1) | (k[w + 1] << 31);
        k[8] >>= 1;
    }
}
// acc = [k] P for the batch-wide scalar whose NAF is `naf`.
template <class Table>
JJ_DEVICE void scalar_mul_wnaf_core(ext_point& acc, const ext_point& P, const NafDigits& naf, Table& tbl) {
    {
        ext_niels n2, nj;
        ext_point cur, P2;
        point_double(P2, P);
        point_to_niels(n2, P2);
        cur = P;
        point_to_niels(nj, cur);
        tbl.store(0, nj);
#pragma unroll 1
        for (int j = 1; j < 8; j++) {
            point_add_niels(cur, cur, n2, false);
            point_to_niels(nj, cur);
            tbl.store(j, nj);
        }
    }
    point_set_identity(acc);
#pragma unroll 1
    for (int i = naf.top; i >= 0; i--) {
        point_double(acc, acc);
        const int d = naf.d[i];
        if (d != 0) {
            ext_niels n;
            tbl.load_signed(((d < 0 ? -d : d) - 1) >> 1, d < 0, n);
            point_add_niels_preswapped(acc, acc, n, d < 0);
        }
    }
}

// ---- fixed base -------------------------------------------------------------------------
// Shared per-window table for one base B: entry (i, j) = affine-Niels of (j+1) * 2^(W*i) * B for
// window i < NW = ceil(252 / W) and j < 2^(W-1), plus one entry 2^(W*NW) * B for the recoding's top
// carry.  [k]B = sum_i d_i * 2^(W*i) * B with signed digits d_i in [-2^(W-1), 2^(W-1) - 1] needs no
// doublings: <= NW + 1 mixed additions of 7M each (src/lib.rs:944-988).  W = 7: 36 windows x 64
// entries x 96 B = 216 KB, the whole shared memory of an SM; W = 4: 63 windows, 47 KB.
template <int W>
struct FixedGeom {
    static_assert(W >= 2 && W <= 16 && W * ((252 + W - 1) / W) <= 256, "the recoding needs W * NW <= 256 (W = 4, 7, 12, 14, 16 ...)");
    static constexpr int NW = (252 + W - 1) / W;
    static constexpr int PER = 1 << (W - 1);
    static constexpr int ENTRIES = NW * PER + 1;
    static constexpr uint32_t BYTES = (uint32_t)ENTRIES * 96u;
    // When W * NW > 252 (W = 16) the top window's multiples (j+1) * 2^(W (NW-1)) can reach 2^252 and beyond, which the
    // 252-bit scalar-mul core cannot take as a scalar: the table builders multiply by (j+1) * 2^(W (NW-1) - EXCESS) and
    // double EXCESS times.  (A digit of the top window never exceeds 2^(252 - W (NW-1)) + 1, but the table is complete.)
    static constexpr int EXCESS = W * NW > 252 ? W * NW - 252 : 0;
};
struct fixed_table_view {
    const uint32_t* base;  // [entry][24 words]
    JJ_DEVICE_SPEC void load(int entry, aff_niels& n) const {
        const uint32_t* p = base + entry * 24;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            n.vpu.w[w] = p[w];
            n.vmu.w[w] = p[8 + w];
            n.t2d.w[w] = p[16 + w];
        }
    }
    // halves exchanged by address for a negative digit (see LocalTable::load_signed)
    JJ_DEVICE_SPEC void load_signed(int entry, bool neg, aff_niels& n) const {
        const uint32_t* p = base + entry * 24;
        const int o = neg ? 8 : 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            n.vpu.w[w] = p[o + w];
            n.vmu.w[w] = p[8 - o + w];
            n.t2d.w[w] = p[16 + w];
        }
    }
};
// t = (k mod 2^252) + sum_{i < NW} 2^(W-1) * 2^(W*i): window i of t, minus 2^(W-1), is digit d_i.
template <int W>
JJ_DEVICE void recode_fixed(uint32_t t[8], const uint32_t k[8]) {
    uint32_t c[8];
#pragma unroll
    for (int w = 0; w < 8; w++) c[w] = 0;
#pragma unroll
    for (int i = 0; i < FixedGeom<W>::NW; i++) {
        const int bit = W * i + W - 1;
        c[bit >> 5] |= 1u << (bit & 31);
    }
    add_cc(t[0], k[0], c[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) addc_cc(t[i], k[i], c[i]);
    addc(t[7], k[7] & 0x0fffffffu, c[7]);
}
template <int W, bool INL, class View = fixed_table_view>
JJ_DEVICE void scalar_mul_fixed_core(ext_point& acc, const uint32_t k[8], const View& tbl) {
    constexpr int NW = FixedGeom<W>::NW, PER = FixedGeom<W>::PER;
    uint32_t t[8];
    recode_fixed<W>(t, k);
    point_set_identity(acc);
#pragma unroll 1
    for (int i = 0; i <= NW; i++) {
        int d = (int)(t[0] & ((1u << W) - 1u)) - (i < NW ? PER : 0);
#pragma unroll
        for (int w = 0; w < 7; w++) t[w] = (t[w] >> W) | (t[w + 1] << (32 - W));
        t[7] >>= W;
        if (d != 0) {
            aff_niels n;
            tbl.load_signed(i * PER + (d < 0 ? -d : d) - 1, d < 0, n);
            point_add_aff_niels_t<INL, kM1MulDefault, true>(acc, acc, n, d < 0);  // <INL, false> (112-multiply rows) measures the same
        }
    }
}

}  // namespace jj
