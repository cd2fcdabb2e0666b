// torsion.cuh -- subgroup membership without a scalar multiplication.
//
// Replaces ExtendedPoint::is_torsion_free (src/lib.rs:709-711: `self.multiply(&FR_MODULUS_BYTES).is_identity()`,
// a 252-step ladder) for the batch engine.  Only the boolean has to match the reference, so the test is the
// order-8 reduced Tate pairing with the fixed 8-torsion point T (the reference's 8-torsion table is
// src/lib.rs:1589-1677):
//
//   E(Fq) is cyclic of order 8r and mu_8 lies in Fq, so  chi(P) = f_{8,T}(P)^((q-1)/8)  is a character of E / 8E
//   of exact order 8: it takes the value it has on the subgroup of order r  <=>  P is in 8E  <=>  [r]P = O.
//
// Miller's loop for the fixed T on the birationally equivalent Montgomery curve collapses, modulo 8th powers and
// constants, to a polynomial in the extended coordinates (derivation, constants and an exhaustive check over all
// eight cosets P + jT: scripts/derive_torsion_check.py, run on the CPU with the big-integer model):
//
//   a = Z + V, b = Z - V,  L1 = aZ - lam1 aU + c1 bU,  G = L1^2 (Z - lam2 U) U V^2,  F = (G b^3)^2 a b
//   torsion free  <=>  U == 0 ? V == Z : F^((q-1)/8) == CHI
//
// (U == 0 are the identity and the point of order 2; F vanishes exactly on T, 2T and -2T, which are not in the
// subgroup.)  Cost: 12 M + 4 S for F, a 223-bit fixed-window power and 29 squarings: about 340 field products against
// the ~2 300 of [r]P -- tests/test_gpu_parity.py checks the flags against the oracle's [r]P == O on every coset.
#pragma once
#include "point.cuh"

namespace jj {

struct TorsionK {
    // Montgomery-form limbs printed by scripts/derive_torsion_check.py
    JJ_CONST_FN uint32_t LAM1(int i) {
        constexpr uint32_t t[8] = {0xa905d6dcu, 0x2b0eab27u, 0x5a49e59du, 0xb847aea0u, 0x3a5d91cdu, 0xf2284c15u, 0x6b820f00u, 0x19186a3fu};
        return t[i];
    }
    JJ_CONST_FN uint32_t C1(int i) {
        constexpr uint32_t t[8] = {0x538fa68du, 0x1ebf019cu, 0xba509f9cu, 0xbf76afe0u, 0x5fc40bf3u, 0xc0c259d4u, 0x98db9275u, 0x66454e44u};
        return t[i];
    }
    JJ_CONST_FN uint32_t LAM2(int i) {
        constexpr uint32_t t[8] = {0xaa89cfb1u, 0xf3b05674u, 0x6006b9feu, 0x072f0140u, 0x25667a26u, 0xce9a0dbfu, 0x2d598374u, 0x4d2ce405u};
        return t[i];
    }
    JJ_CONST_FN uint32_t CHI(int i) {
        constexpr uint32_t t[8] = {0x55763050u, 0x0c4fa98au, 0x9ff7a200u, 0x4c8ea2c2u, 0xe43b5ddfu, 0x649fca48u, 0xfc43f9d3u, 0x26c0c34du};
        return t[i];
    }
};
// T = (q - 1) / 2^32 (223 bits): F^((q-1)/8) = (F^T)^(2^29)
struct ExpFqOddPart {
    static constexpr int NW = 7;
    JJ_CONST_FN uint32_t word(int i) {
        constexpr uint32_t t[7] = {0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return t[i];
    }
};

// The Miller value F(U, V, Z) (any projective scaling of the point gives the same class modulo 8th powers).
JJ_DEVICE void torsion_miller_value(fe& F, const fe& U, const fe& V, const fe& Z) {
    fe a, b, aU, aZ, bU, L1, t, s, k;
    fe_add<FqP>(a, Z, V);
    fe_sub<FqP>(b, Z, V);
    fq_mul(aU, a, U);
    fq_mul(aZ, a, Z);
    fq_mul(bU, b, U);
    JJ_LOAD_CONST(k, TorsionK::LAM1);
    fq_mul(t, k, aU);
    fe_sub<FqP>(L1, aZ, t);
    JJ_LOAD_CONST(k, TorsionK::C1);
    fq_mul(t, k, bU);
    fe_add<FqP>(L1, L1, t);           // L1 = aZ - lam1 aU + c1 bU
    JJ_LOAD_CONST(k, TorsionK::LAM2);
    fq_mul(t, k, U);
    fe_sub<FqP>(t, Z, t);             // Z - lam2 U
    fq_sqr(L1, L1);
    fq_mul(L1, L1, t);                // L1^2 (Z - lam2 U)
    fq_sqr(s, V);
    fq_mul(s, s, U);                  // U V^2
    fq_mul(L1, L1, s);                // G
    fq_sqr(s, b);
    fq_mul(s, s, b);                  // b^3
    fq_mul(L1, L1, s);                // G b^3
    fq_sqr(L1, L1);
    fq_mul(s, a, b);
    fq_mul(F, L1, s);                 // (G b^3)^2 a b
}
// r = a^e with the shared (noinline) product / square bodies: the variant of fe_pow_const the point kernels use
// when code size matters more than call overhead.
template <class EXP>
JJ_DEVICE void fq_pow_const_shared(fe& r, const fe& a) {
    fe tbl[16];
    fe_set_one<FqP>(tbl[0]);
    tbl[1] = a;
#pragma unroll 1
    for (int i = 2; i < 16; i++) fq_mul(tbl[i], tbl[i - 1], a);
    fe acc;
    fe_set_one<FqP>(acc);
    bool started = false;  // the exponent is warp-uniform: leading zeros and zero digits are skipped (fe_pow_const)
#pragma unroll 1
    for (int wi = EXP::NW - 1; wi >= 0; wi--) {
        uint32_t e = EXP::word(wi);
#pragma unroll 1
        for (int s = 28; s >= 0; s -= 4) {
            const uint32_t d = (e >> s) & 15u;
            if (!started) {
                if (d) {
                    acc = tbl[d];
                    started = true;
                }
                continue;
            }
            fq_sqr(acc, acc);
            fq_sqr(acc, acc);
            fq_sqr(acc, acc);
            fq_sqr(acc, acc);
            if (d) fq_mul(acc, acc, tbl[d]);
        }
    }
    r = acc;
}
// is_torsion_free for a point given by its (U, V, Z) -- t1, t2 are not needed.
JJ_DEVICE bool point_is_torsion_free(const fe& U, const fe& V, const fe& Z) {
    fe F, chi;
    torsion_miller_value(F, U, V, Z);
    fq_pow_const_shared<ExpFqOddPart>(F, F);
#pragma unroll 1
    for (int i = 0; i < 29; i++) fq_sqr(F, F);
    JJ_LOAD_CONST(chi, TorsionK::CHI);
    const bool pairing_ok = fe_eq(F, chi);
    return fe_is_zero(U) ? fe_eq(V, Z) : pairing_ok;
}

}  // namespace jj
