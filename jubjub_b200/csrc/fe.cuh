// fe.cuh -- 256-bit Montgomery field arithmetic for Jubjub's Fq and Fr on sm_100a.
//
// A field element is 8 x u32 little-endian limbs held in registers; in memory it is
// byte-identical to the reference's 4 x u64 Montgomery limbs (src/fr.rs:23, R = 2^256).
// Every operation returns the fully reduced representative in [0, m), so results are
// the same bits the reference produces (uniqueness argument).
//
// What replaces what (paths relative to /root/reference):
//   mont_mul    <- Fr::mul + montgomery_reduce   src/fr.rs:592-616, 544-588  ([ext] Fq same shape)
//   mont_sqr    <- Fr::square                    src/fr.rs:353-381
//   fe_add/sub/neg/dbl <- Fr::add/sub/neg/double src/fr.rs:638-647, 620-634, 651-665, 261-263
//   fe_to_canonical    <- Fr::to_bytes           src/fr.rs:296-308
//   fe_invert          <- Fr::invert / [ext] Fq::invert   src/fr.rs:438-540
//
// Multiplication is an operand-scanning (CIOS) Montgomery product on 32-bit limbs with
// the even/odd column split: products whose column index is even accumulate into one
// 8-word array, odd ones into another that sits one word higher, so every row is two
// independent carry chains of four IMAD.WIDE.U32 each and no chain ever has to
// propagate a carry across the other's words.  Per product: 64 IMAD.WIDE (a*b) +
// 64 IMAD.WIDE (q*m) + 8 IMAD (q) = 136 integer-pipe instructions; everything else
// (merges, conditional subtract) is IADD3/LOP3/SEL work on the other pipe.
//
// Precondition shared with the reference's type invariant: the first operand `a` of
// mont_mul is canonical (< m).  The second operand may be any 256-bit value (this is
// what from_raw / from_bytes rely on, src/fr.rs:347-349).
#pragma once
#include "ptx_ops.cuh"

namespace jj {

struct fe {
    uint32_t w[8];
};

// ---- field parameters (32-bit limbs of the constants in SURVEY.md section 8a) ------
struct FqP {  // [ext] bls12_381::Scalar modulus q; q-1 is at src/lib.rs:1629-1634
    static constexpr uint32_t M0 = 0x00000001u, M1 = 0xffffffffu, M2 = 0xfffe5bfeu, M3 = 0x53bda402u,
                              M4 = 0x09a1d805u, M5 = 0x3339d808u, M6 = 0x299d7d48u, M7 = 0x73eda753u;
    static constexpr uint32_t INV = 0xffffffffu;  // -q^-1 mod 2^32
    // R = 2^256 mod q, R2 = 2^512 mod q, R3 = 2^768 mod q
    JJ_CONST_FN uint32_t R(int i) {
        constexpr uint32_t t[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
        return t[i];
    }
    JJ_CONST_FN uint32_t R2(int i) {
        constexpr uint32_t t[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
        return t[i];
    }
    JJ_CONST_FN uint32_t R3(int i) {
        constexpr uint32_t t[8] = {0x439b73afu, 0xc62c1807u, 0x8cf06990u, 0x1b3e0d18u, 0xc7b5f418u, 0x73d13c71u, 0xc8db33e9u, 0x6e2a5bb9u};
        return t[i];
    }
    // m - 2, the Fermat inversion exponent
    JJ_CONST_FN uint32_t M_MINUS_2(int i) {
        constexpr uint32_t t[8] = {0xffffffffu, 0xfffffffeu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return t[i];
    }
    // q - 1 = 2^32 * T with T odd.  (T - 1) / 2 (222 bits, 7 words) and the 2^32-th root of unity
    // 7^T in Montgomery form ([ext] bls12_381 ROOT_OF_UNITY; SURVEY.md section 8f, recomputed).
    JJ_CONST_FN uint32_t T_MINUS_1_HALF(int i) {
        constexpr uint32_t t[8] = {0x7fffffffu, 0x7fff2dffu, 0xa9ded201u, 0x04d0ec02u, 0x199cec04u, 0x94cebea4u, 0x39f6d3a9u, 0u};
        return t[i];
    }
    JJ_CONST_FN uint32_t ROOT_OF_UNITY(int i) {
        constexpr uint32_t t[8] = {0x5f0e466au, 0xb9b58d8cu, 0x1819d7ecu, 0x5b1b4c80u, 0x52a31e64u, 0x0af53ae3u, 0x19e9b27bu, 0x5bf3addau};
        return t[i];
    }
};
struct FrP {  // src/fr.rs:77-82 (MODULUS), :214 (INV), :217-238 (R, R2, R3)
    static constexpr uint32_t M0 = 0xd6f72cb7u, M1 = 0xd0970e5eu, M2 = 0xccc81082u, M3 = 0xa6682093u,
                              M4 = 0x01343b00u, M5 = 0x06673b01u, M6 = 0x6533afa9u, M7 = 0x0e7db4eau;
    static constexpr uint32_t INV = 0xef788ef9u;
    JJ_CONST_FN uint32_t R(int i) {
        constexpr uint32_t t[8] = {0xb99607d9u, 0x25f80bb3u, 0x66b6e750u, 0xf315d62fu, 0xeb8814f4u, 0x932514eeu, 0x479155c6u, 0x09a6fc6fu};
        return t[i];
    }
    JJ_CONST_FN uint32_t R2(int i) {
        constexpr uint32_t t[8] = {0x95e57731u, 0x67719aa4u, 0x9ce3fc26u, 0x51b0cef0u, 0xc026e9a5u, 0x69dab7fau, 0x8d127688u, 0x04f6547bu};
        return t[i];
    }
    JJ_CONST_FN uint32_t R3(int i) {
        constexpr uint32_t t[8] = {0x3d830544u, 0xe0d6c656u, 0x598d0f85u, 0x323e3883u, 0x4c2e2ba8u, 0xf0fea300u, 0x946737ecu, 0x05874f84u};
        return t[i];
    }
    JJ_CONST_FN uint32_t M_MINUS_2(int i) {
        constexpr uint32_t t[8] = {0xd6f72cb5u, 0xd0970e5eu, 0xccc81082u, 0xa6682093u, 0x01343b00u, 0x06673b01u, 0x6533afa9u, 0x0e7db4eau};
        return t[i];
    }
};

// word i of m - 2 with a run-time index (constexpr arrays cannot be ODR-used on the device)
template <class F>
JJ_DEVICE uint32_t exp_word_m_minus_2(int i) {
    switch (i) {
        case 0: return F::M_MINUS_2(0); case 1: return F::M_MINUS_2(1);
        case 2: return F::M_MINUS_2(2); case 3: return F::M_MINUS_2(3);
        case 4: return F::M_MINUS_2(4); case 5: return F::M_MINUS_2(5);
        case 6: return F::M_MINUS_2(6); default: return F::M_MINUS_2(7);
    }
}

// ---- small helpers ---------------------------------------------------------------
JJ_DEVICE void fe_set_zero(fe& r) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.w[i] = 0;
}
template <class F>
JJ_DEVICE void fe_set_one(fe& r) {  // Montgomery one = R
#pragma unroll
    for (int i = 0; i < 8; i++) r.w[i] = F::R(i);
}
JJ_DEVICE bool fe_is_zero(const fe& a) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t |= a.w[i];
    return t == 0;
}
JJ_DEVICE bool fe_eq(const fe& a, const fe& b) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t |= a.w[i] ^ b.w[i];
    return t == 0;
}
// r = c ? b : a   (the reference's conditional_select, src/fr.rs:66-75)
JJ_DEVICE void fe_select(fe& r, const fe& a, const fe& b, bool c) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.w[i] = c ? b.w[i] : a.w[i];
}
JJ_DEVICE void fe_cswap(fe& a, fe& b, bool c) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t x = a.w[i], y = b.w[i];
        a.w[i] = c ? y : x;
        b.w[i] = c ? x : y;
    }
}

// ---- predicated modulus add / subtract / negate ----------------------------------------------
// if (cond) d (+|-)= m, or d = m - d: eight predicated IADD3.X in one carry chain (the instructions issue
// either way; what is saved is the 8 SEL / LOP3 of the select- or mask-based forms).
#if defined(JJ_HOST_EMUL)
template <class F>
JJ_DEVICE void fe_cadd_mod(uint32_t d[8], uint32_t cond) {
    if (!cond) return;
    JJ_ADD_CC_I(d[0], d[0], F::M0); JJ_ADDC_CC_I(d[1], d[1], F::M1); JJ_ADDC_CC_I(d[2], d[2], F::M2);
    JJ_ADDC_CC_I(d[3], d[3], F::M3); JJ_ADDC_CC_I(d[4], d[4], F::M4); JJ_ADDC_CC_I(d[5], d[5], F::M5);
    JJ_ADDC_CC_I(d[6], d[6], F::M6); JJ_ADDC_I(d[7], d[7], F::M7);
}
template <class F>
JJ_DEVICE void fe_cadd_mod_if_negative(uint32_t d[8], uint32_t cond) {
    fe_cadd_mod<F>(d, cond >> 31);
}
template <class F>
JJ_DEVICE void fe_csub_mod(uint32_t d[8], uint32_t cond) {
    if (!cond) return;
    JJ_SUB_CC_I(d[0], d[0], F::M0); JJ_SUBC_CC_I(d[1], d[1], F::M1); JJ_SUBC_CC_I(d[2], d[2], F::M2);
    JJ_SUBC_CC_I(d[3], d[3], F::M3); JJ_SUBC_CC_I(d[4], d[4], F::M4); JJ_SUBC_CC_I(d[5], d[5], F::M5);
    JJ_SUBC_CC_I(d[6], d[6], F::M6); JJ_SUBC_CC_I(d[7], d[7], F::M7);
}
template <class F>
JJ_DEVICE void fe_cneg_mod(uint32_t d[8], uint32_t cond) {
    if (!cond) return;
    sub_cc(d[0], F::M0, d[0]); subc_cc(d[1], F::M1, d[1]); subc_cc(d[2], F::M2, d[2]); subc_cc(d[3], F::M3, d[3]);
    subc_cc(d[4], F::M4, d[4]); subc_cc(d[5], F::M5, d[5]); subc_cc(d[6], F::M6, d[6]); subc(d[7], F::M7, d[7]);
}
#else
#define JJ_PRED8(SETP, OP0, OPC, OPL, SWAP)                                                                       \
    asm volatile(                                                                                                 \
        "{\n\t.reg .pred p;\n\t" SETP " p, %8, 0;\n\t"                                                            \
        "@p " OP0 " %0, " SWAP("%0", "%9") ";\n\t@p " OPC " %1, " SWAP("%1", "%10") ";\n\t"                         \
        "@p " OPC " %2, " SWAP("%2", "%11") ";\n\t@p " OPC " %3, " SWAP("%3", "%12") ";\n\t"                        \
        "@p " OPC " %4, " SWAP("%4", "%13") ";\n\t@p " OPC " %5, " SWAP("%5", "%14") ";\n\t"                        \
        "@p " OPC " %6, " SWAP("%6", "%15") ";\n\t@p " OPL " %7, " SWAP("%7", "%16") ";\n\t}"                       \
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]), "+r"(d[4]), "+r"(d[5]), "+r"(d[6]), "+r"(d[7])          \
        : "r"(cond), "n"(F::M0), "n"(F::M1), "n"(F::M2), "n"(F::M3), "n"(F::M4), "n"(F::M5), "n"(F::M6), "n"(F::M7))
#define JJ_ORD_DM(D, M) D ", " M
#define JJ_ORD_MD(D, M) M ", " D
template <class F>
JJ_DEVICE void fe_cadd_mod(uint32_t d[8], uint32_t cond) {
    JJ_PRED8("setp.ne.u32", "add.cc.u32", "addc.cc.u32", "addc.u32", JJ_ORD_DM);
}
// if ((int32_t)cond < 0) d += m: the sign test is the predicate itself (no shift)
template <class F>
JJ_DEVICE void fe_cadd_mod_if_negative(uint32_t d[8], uint32_t cond) {
    JJ_PRED8("setp.lt.s32", "add.cc.u32", "addc.cc.u32", "addc.u32", JJ_ORD_DM);
}
template <class F>
JJ_DEVICE void fe_csub_mod(uint32_t d[8], uint32_t cond) {
    JJ_PRED8("setp.ne.u32", "sub.cc.u32", "subc.cc.u32", "subc.u32", JJ_ORD_DM);
}
template <class F>
JJ_DEVICE void fe_cneg_mod(uint32_t d[8], uint32_t cond) {
    JJ_PRED8("setp.ne.u32", "sub.cc.u32", "subc.cc.u32", "subc.u32", JJ_ORD_MD);
}
#endif

// r in [0, 2m)  ->  [0, m): trial-subtract m, keep the difference unless it borrowed.
// Same value as the reference's sub(&MODULUS) with mask add-back (src/fr.rs:587, 620-634).
template <class F>
JJ_DEVICE void fe_reduce_once(uint32_t r[8]) {
    uint32_t d[8], borrow;
    JJ_SUB_CC_I(d[0], r[0], F::M0);
    JJ_SUBC_CC_I(d[1], r[1], F::M1);
    JJ_SUBC_CC_I(d[2], r[2], F::M2);
    JJ_SUBC_CC_I(d[3], r[3], F::M3);
    JJ_SUBC_CC_I(d[4], r[4], F::M4);
    JJ_SUBC_CC_I(d[5], r[5], F::M5);
    JJ_SUBC_CC_I(d[6], r[6], F::M6);
    JJ_SUBC_CC_I(d[7], r[7], F::M7);
    subc(borrow, 0u, 0u);  // 0 or 0xffffffff
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = borrow ? r[i] : d[i];
}

// ---- add / sub / neg / double ------------------------------------------------------
template <class F>
JJ_DEVICE void fe_add(fe& r, const fe& a, const fe& b) {
    uint32_t s[8];
    add_cc(s[0], a.w[0], b.w[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) addc_cc(s[i], a.w[i], b.w[i]);
    addc(s[7], a.w[7], b.w[7]);  // 2m < 2^256: no carry out
    fe_reduce_once<F>(s);
#pragma unroll
    for (int i = 0; i < 8; i++) r.w[i] = s[i];
}
template <class F>
JJ_DEVICE void fe_dbl(fe& r, const fe& a) {
    fe_add<F>(r, a, a);
}
// a + b WITHOUT the final reduction: the sum of two canonical values lies in [0, 2m) and 2m < 2^256 for both fields,
// so it still fits 8 limbs.  Such a "lazy" value may only be used where any 256-bit integer is allowed: as the SECOND
// operand of mont_mul (its precondition, and what from_raw relies on) or as the minuend of fe_sub, whose add-back then
// lands in [0, 2m) again.  Saves the 8 trial subtractions + 8 selects (+ borrow) of fe_reduce_once; every value that
// leaves a point formula is still the fully reduced representative.
template <class F>
JJ_DEVICE void fe_add_lazy(fe& r, const fe& a, const fe& b) {
    uint32_t s[8];
    add_cc(s[0], a.w[0], b.w[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) addc_cc(s[i], a.w[i], b.w[i]);
    addc(s[7], a.w[7], b.w[7]);
#pragma unroll
    for (int i = 0; i < 8; i++) r.w[i] = s[i];
}
template <class F>
JJ_DEVICE void fe_sub(fe& r, const fe& a, const fe& b) {
    uint32_t d[8], mask;
    sub_cc(d[0], a.w[0], b.w[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) subc_cc(d[i], a.w[i], b.w[i]);
    subc(mask, 0u, 0u);  // all-ones when a < b
#if defined(JJ_PRED_SUB)
    fe_cadd_mod<F>(d, mask);
#else
    add_cc(d[0], d[0], F::M0 & mask);
    addc_cc(d[1], d[1], F::M1 & mask);
    addc_cc(d[2], d[2], F::M2 & mask);
    addc_cc(d[3], d[3], F::M3 & mask);
    addc_cc(d[4], d[4], F::M4 & mask);
    addc_cc(d[5], d[5], F::M5 & mask);
    addc_cc(d[6], d[6], F::M6 & mask);
    addc(d[7], d[7], F::M7 & mask);
#endif
#pragma unroll
    for (int i = 0; i < 8; i++) r.w[i] = d[i];
}
template <class F>
JJ_DEVICE void fe_neg(fe& r, const fe& a) {
    uint32_t d[8], nz = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) nz |= a.w[i];
    uint32_t mask = nz ? 0xffffffffu : 0u;  // -0 = 0 (src/fr.rs:661-664)
    sub_cc(d[0], F::M0, a.w[0]);
    subc_cc(d[1], F::M1, a.w[1]);
    subc_cc(d[2], F::M2, a.w[2]);
    subc_cc(d[3], F::M3, a.w[3]);
    subc_cc(d[4], F::M4, a.w[4]);
    subc_cc(d[5], F::M5, a.w[5]);
    subc_cc(d[6], F::M6, a.w[6]);
    subc(d[7], F::M7, a.w[7]);
#pragma unroll
    for (int i = 0; i < 8; i++) r.w[i] = d[i] & mask;
}

// ---- Montgomery multiplication -------------------------------------------------------
// One reduction row: q = E[0] * (-m^-1); E += q*(m0,m2,m4,m6); O += q*(m1,m3,m5,m7).
// E is the array whose word 0 sits at the current column, O sits one column higher.
// After the row E[0] == 0.  The O chain cannot carry out (running total < 2^(32*9) at
// this column, see DESIGN.md "carry bounds"); the E chain's carry lands in O[7].
template <class F>
JJ_DEVICE void redc_row_generic(uint32_t E[8], uint32_t O[8]) {
    uint32_t q = E[0] * F::INV;
    JJ_MAD_LO_CC_I(O[0], q, F::M1, O[0]);
    JJ_MADC_HI_CC_I(O[1], q, F::M1, O[1]);
    JJ_MADC_LO_CC_I(O[2], q, F::M3, O[2]);
    JJ_MADC_HI_CC_I(O[3], q, F::M3, O[3]);
    JJ_MADC_LO_CC_I(O[4], q, F::M5, O[4]);
    JJ_MADC_HI_CC_I(O[5], q, F::M5, O[5]);
    JJ_MADC_LO_CC_I(O[6], q, F::M7, O[6]);
    JJ_MADC_HI_I(O[7], q, F::M7, O[7]);
    JJ_MAD_LO_CC_I(E[0], q, F::M0, E[0]);
    JJ_MADC_HI_CC_I(E[1], q, F::M0, E[1]);
    JJ_MADC_LO_CC_I(E[2], q, F::M2, E[2]);
    JJ_MADC_HI_CC_I(E[3], q, F::M2, E[3]);
    JJ_MADC_LO_CC_I(E[4], q, F::M4, E[4]);
    JJ_MADC_HI_CC_I(E[5], q, F::M4, E[5]);
    JJ_MADC_LO_CC_I(E[6], q, F::M6, E[6]);
    JJ_MADC_HI_CC_I(E[7], q, F::M6, E[7]);
    addc(O[7], O[7], oz());
}
// Fq specialisation.  q's low limbs are m0 = 1, m1 = 2^32 - 1 and -q^-1 = 2^32 - 1, so
//   k       = -E[0]                                   (no multiply)
//   k * m0  : column 0 cancels E[0]; its carry c0 = [E[0] != 0] moves one column up
//   k * m1  = k * 2^32 - k : low word = E[0], high word = k - c0
// i.e. O[0] += E[0] + c0 and O[1] += k - c0.  Done with IADD3s that is 14 multiplier instructions per
// row instead of the generic 17 (112 per product); done as ONE IMAD.WIDE.X with carry-in c0 it is 15 per
// row (119 per product with the subtractive last row below) and three ALU instructions fewer.  E[1] is
// untouched either way.
// M1MUL selects how k*m1 + c0 is added: true = one IMAD.WIDE.X (default of the point kernels, whose bound is the
// issue rate), false = three IADD3 (the elementwise field kernels, which are HBM- or multiplier-bound and want
// the fewest multiplier instructions).
template <bool M1MUL>
JJ_DEVICE void redc_row_fq(uint32_t E[8], uint32_t O[8]) {
    uint32_t e0 = E[0], q, hi, t;
    if (M1MUL) {
    // (O[1]:O[0]) += k * m1 + c0 as ONE IMAD.WIDE.X whose carry-in is c0: 2 ALU + 1 multiplier
    // instruction instead of 5 ALU instructions.  Measured faster: the ALU/issue side of this loop
    // costs about as much as the multiplier side.
    // (The borrow of the negation IS c0, but a madc/addc that consumes the flag of a sub.cc gets the
    // un-inverted SASS predicate on sm_100a / CUDA 12.9 -- wrong results -- so c0 is regenerated.)
    sub_cc(q, oz(), e0);         // q = -e0
    add_cc(t, e0, 0xffffffffu);  // carry = c0 = [e0 != 0]
    JJ_MADC_LO_CC_I(O[0], q, FqP::M1, O[0]);
    JJ_MADC_HI_CC_I(O[1], q, FqP::M1, O[1]);
    } else {
    sub_cc(q, oz(), e0);   // q = -e0, borrow = c0
    subc(hi, q, 0u);     // hi = q - c0
    add_cc(t, e0, 0xffffffffu);  // carry = c0
    addc_cc(O[0], O[0], e0);
    addc_cc(O[1], O[1], hi);
    }
    JJ_MADC_LO_CC_I(O[2], q, FqP::M3, O[2]);
    JJ_MADC_HI_CC_I(O[3], q, FqP::M3, O[3]);
    JJ_MADC_LO_CC_I(O[4], q, FqP::M5, O[4]);
    JJ_MADC_HI_CC_I(O[5], q, FqP::M5, O[5]);
    JJ_MADC_LO_CC_I(O[6], q, FqP::M7, O[6]);
    JJ_MADC_HI_I(O[7], q, FqP::M7, O[7]);
    JJ_MAD_LO_CC_I(E[2], q, FqP::M2, E[2]);
    JJ_MADC_HI_CC_I(E[3], q, FqP::M2, E[3]);
    JJ_MADC_LO_CC_I(E[4], q, FqP::M4, E[4]);
    JJ_MADC_HI_CC_I(E[5], q, FqP::M4, E[5]);
    JJ_MADC_LO_CC_I(E[6], q, FqP::M6, E[6]);
    JJ_MADC_HI_CC_I(E[7], q, FqP::M6, E[7]);
    addc(O[7], O[7], oz());
    E[0] = 0;
    (void)t;
    (void)hi;
}
#if defined(JJ_REDC_M1_IMAD)
constexpr bool kM1MulDefault = true;
#else
constexpr bool kM1MulDefault = false;
#endif
template <class F, bool M1MUL>
struct RedcRow {
    static JJ_DEVICE_SPEC void run(uint32_t E[8], uint32_t O[8]) { redc_row_generic<F>(E, O); }
};
template <bool M1MUL>
struct RedcRow<FqP, M1MUL> {
    static JJ_DEVICE_SPEC void run(uint32_t E[8], uint32_t O[8]) { redc_row_fq<M1MUL>(E, O); }
};
template <class F, bool M1MUL = kM1MulDefault>
JJ_DEVICE void redc_row(uint32_t E[8], uint32_t O[8]) {
    RedcRow<F, M1MUL>::run(E, O);
}
// One product row for multiplier word bi.  On entry O is the array that was column-
// aligned in the previous row: O[0] is dead (zeroed by redc_row), O[1] sits at this
// row's column and is folded into E[0]; O[2..7] slide down two words while the odd
// products are added, so O ends up one column above E again.
JJ_DEVICE void mul_row(uint32_t E[8], uint32_t O[8], const uint32_t a[8], uint32_t bi) {
    add_cc(E[0], E[0], O[1]);
    madc_lo_cc(O[0], a[1], bi, O[2]);
    madc_hi_cc(O[1], a[1], bi, O[3]);
    madc_lo_cc(O[2], a[3], bi, O[4]);
    madc_hi_cc(O[3], a[3], bi, O[5]);
    madc_lo_cc(O[4], a[5], bi, O[6]);
    madc_hi_cc(O[5], a[5], bi, O[7]);
    madc_lo_cc(O[6], a[7], bi, 0u);
    madc_hi(O[7], a[7], bi, 0u);
    mad_lo_cc(E[0], a[0], bi, E[0]);
    madc_hi_cc(E[1], a[0], bi, E[1]);
    madc_lo_cc(E[2], a[2], bi, E[2]);
    madc_hi_cc(E[3], a[2], bi, E[3]);
    madc_lo_cc(E[4], a[4], bi, E[4]);
    madc_hi_cc(E[5], a[4], bi, E[5]);
    madc_lo_cc(E[6], a[6], bi, E[6]);
    madc_hi_cc(E[7], a[6], bi, E[7]);
    addc(O[7], O[7], oz());
}
// First row: nothing accumulated yet, plain 32x32->64 products (IMAD.WIDE.U32 with RZ).
JJ_DEVICE void mul_row0(uint32_t E[8], uint32_t O[8], const uint32_t a[8], uint32_t b0) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint64_t e = (uint64_t)a[2 * k] * b0, o = (uint64_t)a[2 * k + 1] * b0;
        E[2 * k] = (uint32_t)e;
        E[2 * k + 1] = (uint32_t)(e >> 32);
        O[2 * k] = (uint32_t)o;
        O[2 * k + 1] = (uint32_t)(o >> 32);
    }
}
// After the last row: E holds columns 8..14 in E[1..7] (E[0] dead), O holds 8..15.
template <class F>
JJ_DEVICE void mont_finish(fe& r, const uint32_t E[8], const uint32_t O[8]) {
    uint32_t s[8];
    add_cc(s[0], O[0], E[1]);
#pragma unroll
    for (int i = 1; i < 7; i++) addc_cc(s[i], O[i], E[i + 1]);
    addc(s[7], O[7], 0u);
    fe_reduce_once<F>(s);
#pragma unroll
    for (int i = 0; i < 8; i++) r.w[i] = s[i];
}

// Fq: the LAST reduction row subtracts E[0] * q instead of adding (2^32 - E[0]) * q (q = 1 mod 2^32, so the
// quotient digit of the subtractive form is E[0] itself).  The result is the additive one minus q, i.e. it lies
// in (-q, q) instead of [0, 2q): its sign is bit 255 of the 8-word value and the final correction becomes
// "add q if negative" -- 1 + 8 predicated instructions instead of the 8 subtractions + 8 selects of
// fe_reduce_once -- and the row itself needs no negation, no carry regeneration and no k * m1 product.
// Subtracting is adding E[0] * (2^288 - q) modulo the 9-word window: 2^288 - q has limbs
// (2^32 - 1, 0, ~m2, ..., ~m7, 2^32 - 1); column 0 cancels into a plain E[0] at column 1, column 8 receives
// -E[0] and everything above the window is dropped, as any carry out of it always was.
#if defined(JJ_LAST_ROW_SUB)
constexpr bool kLastRowSub = true;
#else
constexpr bool kLastRowSub = false;
#endif
JJ_DEVICE void redc_row_fq_last(uint32_t E[8], uint32_t O[8]) {
    const uint32_t e0 = E[0];
    add_cc(O[0], O[0], e0);
    addc_cc(O[1], O[1], oz());
    JJ_MADC_LO_CC_I(O[2], e0, (uint32_t)~FqP::M3, O[2]);
    JJ_MADC_HI_CC_I(O[3], e0, (uint32_t)~FqP::M3, O[3]);
    JJ_MADC_LO_CC_I(O[4], e0, (uint32_t)~FqP::M5, O[4]);
    JJ_MADC_HI_CC_I(O[5], e0, (uint32_t)~FqP::M5, O[5]);
    JJ_MADC_LO_CC_I(O[6], e0, (uint32_t)~FqP::M7, O[6]);
    JJ_MADC_HI_I(O[7], e0, (uint32_t)~FqP::M7, O[7]);
    JJ_MAD_LO_CC_I(E[2], e0, (uint32_t)~FqP::M2, E[2]);
    JJ_MADC_HI_CC_I(E[3], e0, (uint32_t)~FqP::M2, E[3]);
    JJ_MADC_LO_CC_I(E[4], e0, (uint32_t)~FqP::M4, E[4]);
    JJ_MADC_HI_CC_I(E[5], e0, (uint32_t)~FqP::M4, E[5]);
    JJ_MADC_LO_CC_I(E[6], e0, (uint32_t)~FqP::M6, E[6]);
    JJ_MADC_HI_CC_I(E[7], e0, (uint32_t)~FqP::M6, E[7]);
    addc(O[7], O[7], oz());
    O[7] -= e0;
    E[0] = 0;
}
// merge + "add q if negative" (pairs with redc_row_fq_last)
JJ_DEVICE void mont_finish_signed_fq(fe& r, const uint32_t E[8], const uint32_t O[8]) {
    uint32_t s[8];
    add_cc(s[0], O[0], E[1]);
#pragma unroll
    for (int i = 1; i < 7; i++) addc_cc(s[i], O[i], E[i + 1]);
    addc(s[7], O[7], 0u);
    fe_cadd_mod_if_negative<FqP>(s, s[7]);
#pragma unroll
    for (int i = 0; i < 8; i++) r.w[i] = s[i];
}
template <class F, bool M1MUL>
struct LastRow {
    static JJ_DEVICE_SPEC void run(fe& r, uint32_t E[8], uint32_t O[8]) {
        redc_row<F, M1MUL>(E, O);
        mont_finish<F>(r, E, O);
    }
};
template <bool M1MUL>
struct LastRow<FqP, M1MUL> {
    static JJ_DEVICE_SPEC void run(fe& r, uint32_t E[8], uint32_t O[8]) {
        if (kLastRowSub) {
            redc_row_fq_last(E, O);
            mont_finish_signed_fq(r, E, O);
        } else {
            redc_row<FqP, M1MUL>(E, O);
            mont_finish<FqP>(r, E, O);
        }
    }
};

// r = a * b * 2^-256 mod m.  `a` canonical, `b` any 256-bit value.  r may alias a or b.
template <class F, bool M1MUL = kM1MulDefault>
JJ_DEVICE void mont_mul(fe& r, const fe& a, const fe& b) {
    uint32_t X[8], Y[8];
    mul_row0(X, Y, a.w, b.w[0]);
    redc_row<F, M1MUL>(X, Y);
    mul_row(Y, X, a.w, b.w[1]);
    redc_row<F, M1MUL>(Y, X);
    mul_row(X, Y, a.w, b.w[2]);
    redc_row<F, M1MUL>(X, Y);
    mul_row(Y, X, a.w, b.w[3]);
    redc_row<F, M1MUL>(Y, X);
    mul_row(X, Y, a.w, b.w[4]);
    redc_row<F, M1MUL>(X, Y);
    mul_row(Y, X, a.w, b.w[5]);
    redc_row<F, M1MUL>(Y, X);
    mul_row(X, Y, a.w, b.w[6]);
    redc_row<F, M1MUL>(X, Y);
    mul_row(Y, X, a.w, b.w[7]);
    LastRow<F, M1MUL>::run(r, Y, X);
}

// ---- Montgomery squaring ---------------------------------------------------------------
// a^2 = sum_i a_i B^i (a_i B^i + 2 * sum_{j>i} a_j B^j): row i multiplies a_i by the vector
// (a_i, 2a_{i+1}, ..., 2a_7) taken from the bit-doubled operand, so only 8 - i products are
// issued (36 IMAD.WIDE instead of 64; with Fq's reduction rows 84 .. 91 in total).  Columns below the first
// product of a chain only propagate the merge carry (addc).
//
// Row i already contains the doubled cross terms of later rows, so the running total is bounded
// by (2a + m)(B + 1) instead of 2m(B + 1).  It still fits the 9-word window -- and every carry
// argument of mont_mul still holds -- as long as 2a + m < 2^256 (1 - 2^-32).  That is always true
// for Fr (3r < 2^254); for Fq it is true for a <= q/2, so larger inputs are squared as q - a
// (same square), selected by the top limb: one 8-limb subtraction and 8 selects, all ALU work,
// instead of tracking a 257th bit on the multiplier pipe.  Needs a < m (canonical input).
template <int I, int J>
JJ_DEVICE uint32_t sqr_operand(const uint32_t a[8], const uint32_t dw[8]) {
    return J == I ? a[I] : (J == I + 1 ? (dw[J] & 0xfffffffeu) : dw[J]);
}
// Row I >= 1 of the squaring.  Same column bookkeeping as mul_row.
template <int I>
JJ_DEVICE void sqr_row(uint32_t E[8], uint32_t O[8], const uint32_t a[8], const uint32_t dw[8]) {
    const uint32_t bi = a[I];
    add_cc(E[0], E[0], O[1]);
    // odd columns J = 1, 3, 5, 7 -> O pairs (0,1), (2,3), (4,5), (6,7), sliding down two words
    if (1 >= I) {
        madc_lo_cc(O[0], sqr_operand<I, 1>(a, dw), bi, O[2]);
        madc_hi_cc(O[1], sqr_operand<I, 1>(a, dw), bi, O[3]);
    } else {
        addc_cc(O[0], O[2], 0u);
        addc_cc(O[1], O[3], 0u);
    }
    if (3 >= I) {
        madc_lo_cc(O[2], sqr_operand<I, 3>(a, dw), bi, O[4]);
        madc_hi_cc(O[3], sqr_operand<I, 3>(a, dw), bi, O[5]);
    } else {
        addc_cc(O[2], O[4], 0u);
        addc_cc(O[3], O[5], 0u);
    }
    if (5 >= I) {
        madc_lo_cc(O[4], sqr_operand<I, 5>(a, dw), bi, O[6]);
        madc_hi_cc(O[5], sqr_operand<I, 5>(a, dw), bi, O[7]);
    } else {
        addc_cc(O[4], O[6], 0u);
        addc_cc(O[5], O[7], 0u);
    }
    madc_lo_cc(O[6], sqr_operand<I, 7>(a, dw), bi, 0u);
    madc_hi(O[7], sqr_operand<I, 7>(a, dw), bi, 0u);
    // even columns J = 0, 2, 4, 6 -> E pairs, in place; the chain starts at the first J >= I
    if (I <= 6) {
        if (I <= 0) {
            mad_lo_cc(E[0], sqr_operand<I, 0>(a, dw), bi, E[0]);
            madc_hi_cc(E[1], sqr_operand<I, 0>(a, dw), bi, E[1]);
        }
        if (I <= 2) {
            if (I > 0) {
                mad_lo_cc(E[2], sqr_operand<I, 2>(a, dw), bi, E[2]);
            } else {
                madc_lo_cc(E[2], sqr_operand<I, 2>(a, dw), bi, E[2]);
            }
            madc_hi_cc(E[3], sqr_operand<I, 2>(a, dw), bi, E[3]);
        }
        if (I <= 4) {
            if (I > 2) {
                mad_lo_cc(E[4], sqr_operand<I, 4>(a, dw), bi, E[4]);
            } else {
                madc_lo_cc(E[4], sqr_operand<I, 4>(a, dw), bi, E[4]);
            }
            madc_hi_cc(E[5], sqr_operand<I, 4>(a, dw), bi, E[5]);
        }
        if (I > 4) {
            mad_lo_cc(E[6], sqr_operand<I, 6>(a, dw), bi, E[6]);
        } else {
            madc_lo_cc(E[6], sqr_operand<I, 6>(a, dw), bi, E[6]);
        }
        madc_hi_cc(E[7], sqr_operand<I, 6>(a, dw), bi, E[7]);
        addc(O[7], O[7], oz());
    }
}
// x = a or m - a, whichever has its top limb <= m7 / 2 (then 2x + m < 2^256 (1 - 2^-32)).
template <class F>
JJ_DEVICE void sqr_fold(uint32_t x[8], const fe& a) {
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = a.w[i];
    constexpr bool needed = !(3.0 * (double)F::M7 < 4294967295.0 * 0.999);  // 3m < 2^256 already (Fr): no fold
    if (needed) {
        const bool big = a.w[7] > (F::M7 >> 1);
#if defined(JJ_PRED_FOLD)
        fe_cneg_mod<F>(x, big ? 1u : 0u);
#else
        uint32_t n[8];
        sub_cc(n[0], F::M0, a.w[0]);
        subc_cc(n[1], F::M1, a.w[1]);
        subc_cc(n[2], F::M2, a.w[2]);
        subc_cc(n[3], F::M3, a.w[3]);
        subc_cc(n[4], F::M4, a.w[4]);
        subc_cc(n[5], F::M5, a.w[5]);
        subc_cc(n[6], F::M6, a.w[6]);
        subc(n[7], F::M7, a.w[7]);
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = big ? n[i] : a.w[i];
#endif
    }
}
template <class F, bool M1MUL = kM1MulDefault>
JJ_DEVICE void mont_sqr(fe& r, const fe& a) {
    uint32_t X[8], Y[8], dw[8], x[8];
    sqr_fold<F>(x, a);
    dw[0] = 0;
#pragma unroll
    for (int j = 1; j < 8; j++) dw[j] = (x[j] << 1) | (x[j - 1] >> 31);
    {   // row 0: x_0 * (x_0, 2x_1, ..., 2x_7), nothing accumulated yet
        const uint32_t b0 = x[0];
        uint64_t e, o;
        e = (uint64_t)x[0] * b0;                   X[0] = (uint32_t)e; X[1] = (uint32_t)(e >> 32);
        o = (uint64_t)(dw[1] & 0xfffffffeu) * b0;  Y[0] = (uint32_t)o; Y[1] = (uint32_t)(o >> 32);
        e = (uint64_t)dw[2] * b0;                  X[2] = (uint32_t)e; X[3] = (uint32_t)(e >> 32);
        o = (uint64_t)dw[3] * b0;                  Y[2] = (uint32_t)o; Y[3] = (uint32_t)(o >> 32);
        e = (uint64_t)dw[4] * b0;                  X[4] = (uint32_t)e; X[5] = (uint32_t)(e >> 32);
        o = (uint64_t)dw[5] * b0;                  Y[4] = (uint32_t)o; Y[5] = (uint32_t)(o >> 32);
        e = (uint64_t)dw[6] * b0;                  X[6] = (uint32_t)e; X[7] = (uint32_t)(e >> 32);
        o = (uint64_t)dw[7] * b0;                  Y[6] = (uint32_t)o; Y[7] = (uint32_t)(o >> 32);
    }
    redc_row<F, M1MUL>(X, Y);
    sqr_row<1>(Y, X, x, dw);
    redc_row<F, M1MUL>(Y, X);
    sqr_row<2>(X, Y, x, dw);
    redc_row<F, M1MUL>(X, Y);
    sqr_row<3>(Y, X, x, dw);
    redc_row<F, M1MUL>(Y, X);
    sqr_row<4>(X, Y, x, dw);
    redc_row<F, M1MUL>(X, Y);
    sqr_row<5>(Y, X, x, dw);
    redc_row<F, M1MUL>(Y, X);
    sqr_row<6>(X, Y, x, dw);
    redc_row<F, M1MUL>(X, Y);
    sqr_row<7>(Y, X, x, dw);
    LastRow<F, M1MUL>::run(r, Y, X);
}

// Montgomery form -> canonical integer: one reduction of (a, 0), i.e. a * 1 (src/fr.rs:296-308).
template <class F>
JJ_DEVICE void fe_to_canonical(fe& r, const fe& a) {
    fe one;
    fe_set_zero(one);
    one.w[0] = 1;
    mont_mul<F>(r, a, one);
}
// canonical (or any 256-bit) integer -> Montgomery form: R2 * v (src/fr.rs:347-349).
template <class F>
JJ_DEVICE void fe_from_raw(fe& r, const fe& v) {
    fe r2;
#pragma unroll
    for (int i = 0; i < 8; i++) r2.w[i] = F::R2(i);
    mont_mul<F>(r, r2, v);
}
// v < m ?  (the canonical check of from_bytes, src/fr.rs:277-286)
template <class F>
JJ_DEVICE bool fe_is_canonical(const fe& v) {
    uint32_t d, borrow;
    JJ_SUB_CC_I(d, v.w[0], F::M0);
    JJ_SUBC_CC_I(d, v.w[1], F::M1);
    JJ_SUBC_CC_I(d, v.w[2], F::M2);
    JJ_SUBC_CC_I(d, v.w[3], F::M3);
    JJ_SUBC_CC_I(d, v.w[4], F::M4);
    JJ_SUBC_CC_I(d, v.w[5], F::M5);
    JJ_SUBC_CC_I(d, v.w[6], F::M6);
    JJ_SUBC_CC_I(d, v.w[7], F::M7);
    subc(borrow, 0u, 0u);
    (void)d;
    return borrow != 0;
}

// r = a^e for a constant exponent (EXP::word(i), EXP::NW 32-bit words): 4-bit fixed window.  The exponent is the same for
// every lane, so its zero digits are skipped by a warp-uniform branch and the squarings of the leading one are not run
// (the accumulator starts at the first non-zero digit's table entry).
template <class F, class EXP>
JJ_DEVICE void fe_pow_const(fe& r, const fe& a) {
    fe tbl[16];
    fe_set_one<F>(tbl[0]);
    tbl[1] = a;
#pragma unroll 1
    for (int i = 2; i < 16; i++) mont_mul<F>(tbl[i], tbl[i - 1], a);
    fe acc;
    fe_set_one<F>(acc);
    bool started = false;
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
            mont_sqr<F>(acc, acc);
            mont_sqr<F>(acc, acc);
            mont_sqr<F>(acc, acc);
            mont_sqr<F>(acc, acc);
            if (d) mont_mul<F>(acc, acc, tbl[d]);
        }
    }
    r = acc;
}
template <class F>
struct ExpInvert {
    static constexpr int NW = 8;
    JJ_CONST_FN uint32_t word(int i) { return F::M_MINUS_2(i); }
};
struct ExpFqSqrt {
    static constexpr int NW = 7;
    JJ_CONST_FN uint32_t word(int i) { return FqP::T_MINUS_1_HALF(i); }
};
// a^(m-2): the result is the unique inverse (0 for a = 0, flagged by the caller like CtOption,
// src/fr.rs:539), so the reference's particular addition chain (src/fr.rs:438-538) need not be replayed.
template <class F>
JJ_DEVICE void fe_invert(fe& r, const fe& a) {
    fe_pow_const<F, ExpInvert<F>>(r, a);
}
// Square root in Fq ([ext] bls12_381::Scalar::sqrt; call sites src/lib.rs:515, :603): Tonelli-Shanks
// with q - 1 = 2^32 * T.  Returns false for a non-residue.  Which root comes back does not matter to
// the callers on this path: the sign is fixed from the parity afterwards (src/lib.rs:518-520).
// Fr::sqrt (src/fr.rs:384-399): r = 3 (mod 4), so the candidate root is a^((r+1)/4); Some iff it squares to a.
struct ExpFrSqrt {
    static constexpr int NW = 8;
    JJ_CONST_FN uint32_t word(int i) {
        constexpr uint32_t t[8] = {0xb5bdcb2eu, 0xb425c397u, 0xf3320420u, 0x299a0824u, 0x404d0ec0u, 0x4199cec0u, 0x994cebeau, 0x039f6d3au};
        return t[i];
    }
};
JJ_DEVICE bool fr_sqrt(fe& r, const fe& a) {
    fe s, s2;
    fe_pow_const<FrP, ExpFrSqrt>(s, a);
    mont_sqr<FrP>(s2, s);
    r = s;
    return fe_eq(s2, a);
}
// Tonelli-Shanks loop form (kept as the cross-check of the table-driven fq_sqrt below; same root).
JJ_DEVICE bool fq_sqrt_ts(fe& r, const fe& a) {
    if (fe_is_zero(a)) {
        fe_set_zero(r);
        return true;
    }
    fe w, x, b, z, one;
    fe_set_one<FqP>(one);
    fe_pow_const<FqP, ExpFqSqrt>(w, a);  // a^((T-1)/2)
    mont_mul<FqP>(x, a, w);              // a^((T+1)/2)
    mont_mul<FqP>(b, x, w);              // a^T
#pragma unroll
    for (int i = 0; i < 8; i++) z.w[i] = FqP::ROOT_OF_UNITY(i);
    int m = 32;
#pragma unroll 1
    while (!fe_eq(b, one)) {
        int i = 0;
        fe bb = b;
#pragma unroll 1
        while (!fe_eq(bb, one)) {
            mont_sqr<FqP>(bb, bb);
            if (++i >= m) return false;  // order 2^m: a is a non-residue
        }
        fe t = z;
#pragma unroll 1
        for (int k = 0; k < m - i - 1; k++) mont_sqr<FqP>(t, t);
        mont_mul<FqP>(x, x, t);
        mont_sqr<FqP>(z, t);
        mont_mul<FqP>(b, b, z);
        m = i;
    }
    r = x;
    return true;
}

// ---- table-driven square root in Fq ------------------------------------------------------------------
// q - 1 = 2^32 * T.  With x = a^((T+1)/2) and b = a^T (an element of the cyclic 2^32-torsion <g>, g = 7^T),
// sqrt(a) = x * g^s where 2s = -log_g(b) (mod 2^32); a is a residue iff that logarithm is even.  The loop form
// above finds s one set bit at a time (about 16 data-dependent rounds of up to 31 squarings, and a warp pays
// the maximum over its lanes); here log_g(b) is read 8 bits at a time (Pohlig-Hellman): raising to 2^24,
// 2^16, 2^8, 1 lands in the order-256 subgroup <g^(2^24)>, whose elements are recognised by a perfect hash of
// 13 bits of their Montgomery limbs, and each recovered byte is divided out with one table product.
// Fixed cost after the 222-bit power: 48 squarings + 9 products + 8 table reads, no data-dependent branches.
// s is taken in [0, 2^31) like the loop form, so both return the same root.
constexpr int kSqrtHashBits = 13, kSqrtHashLimb = 1, kSqrtHashShift = 11;
struct alignas(32) FqSqrtTables {
    fe neg[3][256];                    // g^(-i * 2^(8k)), k = 0, 1, 2
    fe pos[4][256];                    // g^(+i * 2^(8k)), k = 0..3; pos[3] is the order-256 subgroup
    uint8_t hash[1 << kSqrtHashBits];  // key(pos[3][i]) -> i
    uint32_t ready;
};
#if defined(JJ_HOST_EMUL)
static FqSqrtTables g_fq_sqrt_tab;
#else
__device__ FqSqrtTables g_fq_sqrt_tab;
#endif
JJ_DEVICE uint32_t fq_sqrt_key(const fe& y) {
    return (y.w[kSqrtHashLimb] >> kSqrtHashShift) & ((1u << kSqrtHashBits) - 1u);
}
JJ_DEVICE void fe_load_tab(fe& r, const fe* p) {
#if defined(JJ_HOST_EMUL)
    r = *p;
#else
    const uint4* q4 = reinterpret_cast<const uint4*>(p);
    uint4 lo = q4[0], hi = q4[1];
    r.w[0] = lo.x; r.w[1] = lo.y; r.w[2] = lo.z; r.w[3] = lo.w;
    r.w[4] = hi.x; r.w[5] = hi.y; r.w[6] = hi.z; r.w[7] = hi.w;
#endif
}
// Fills the tables (one thread; ~1 900 products).  Returns false if the hash is not perfect on the subgroup
// (it is, for the constants above: checked here at every initialisation and by the tests).
JJ_DEVICE bool fq_sqrt_tables_build(FqSqrtTables& t) {
    fe g, gi, one, sp, sn;
#pragma unroll
    for (int i = 0; i < 8; i++) g.w[i] = FqP::ROOT_OF_UNITY(i);
    fe_invert<FqP>(gi, g);
    fe_set_one<FqP>(one);
    sp = g;
    sn = gi;
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        fe cp = one, cn = one;
#pragma unroll 1
        for (int i = 0; i < 256; i++) {
            t.pos[k][i] = cp;
            mont_mul<FqP>(cp, cp, sp);
            if (k < 3) {
                t.neg[k][i] = cn;
                mont_mul<FqP>(cn, cn, sn);
            }
        }
#pragma unroll 1
        for (int j = 0; j < 8; j++) {
            mont_sqr<FqP>(sp, sp);
            mont_sqr<FqP>(sn, sn);
        }
    }
#pragma unroll 1
    for (int i = 0; i < (1 << kSqrtHashBits); i++) t.hash[i] = 0xff;
    bool perfect = true;
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
        uint32_t key = fq_sqrt_key(t.pos[3][i]);
        if (t.hash[key] != 0xff) perfect = false;
        t.hash[key] = (uint8_t)i;
    }
    t.ready = perfect ? 1u : 0u;
    return perfect;
}
JJ_DEVICE bool fq_sqrt(fe& r, const fe& a) {
#if defined(JJ_HOST_EMUL)
    if (!g_fq_sqrt_tab.ready) fq_sqrt_tables_build(g_fq_sqrt_tab);
#endif
    if (fe_is_zero(a)) {
        fe_set_zero(r);
        return true;
    }
    const FqSqrtTables& t = g_fq_sqrt_tab;
    fe w, x, b, y, f;
    fe_pow_const<FqP, ExpFqSqrt>(w, a);  // a^((T-1)/2)
    mont_mul<FqP>(x, a, w);              // a^((T+1)/2)
    mont_mul<FqP>(b, x, w);              // a^T
    uint32_t dlog = 0;
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        y = b;
#pragma unroll 1
        for (int j = 0; j < 24 - 8 * k; j++) mont_sqr<FqP>(y, y);
        uint32_t d = t.hash[fq_sqrt_key(y)];
        dlog |= d << (8 * k);
        if (k < 3) {
            fe_load_tab(f, &t.neg[k][d & 255u]);
            mont_mul<FqP>(b, b, f);  // divide the recovered byte out
        }
    }
    if (dlog & 1u) return false;  // odd logarithm: a is a non-residue
    uint32_t s = (0x80000000u - (dlog >> 1)) & 0x7fffffffu;
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        fe_load_tab(f, &t.pos[k][(s >> (8 * k)) & 255u]);
        mont_mul<FqP>(x, x, f);
    }
    r = x;
    return true;
}
// r = sqrt(num / den) without an inversion (den != 0).  With w = num * den and x = w^((T-1)/2): w x^2 = w^T = b lies in
// the 2^32-torsion, and g^s = b^(-1/2) as in fq_sqrt, so x g^s = w^(-1/2) and num * w^(-1/2) = sqrt(num / den).  num / den is a
// residue iff w is (they differ by den^2).  Costs two products more than fq_sqrt and replaces the batched inversion of
// the denominators in batch_from_bytes (src/lib.rs:596-600): the decoded point is the same, because the root's sign is
// fixed from the encoding afterwards (src/lib.rs:518-520).
JJ_DEVICE bool fq_sqrt_ratio(fe& r, const fe& num, const fe& den) {
#if defined(JJ_HOST_EMUL)
    if (!g_fq_sqrt_tab.ready) fq_sqrt_tables_build(g_fq_sqrt_tab);
#endif
    fe w, x, b, y, f;
    mont_mul<FqP>(w, num, den);
    fe_set_zero(r);
    if (fe_is_zero(w)) return fe_is_zero(num);
    const FqSqrtTables& t = g_fq_sqrt_tab;
    fe_pow_const<FqP, ExpFqSqrt>(x, w);  // w^((T-1)/2)
    mont_mul<FqP>(b, w, x);              // w^((T+1)/2)
    mont_mul<FqP>(b, b, x);              // w^T
    uint32_t dlog = 0;
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        y = b;
#pragma unroll 1
        for (int j = 0; j < 24 - 8 * k; j++) mont_sqr<FqP>(y, y);
        uint32_t d = t.hash[fq_sqrt_key(y)];
        dlog |= d << (8 * k);
        if (k < 3) {
            fe_load_tab(f, &t.neg[k][d & 255u]);
            mont_mul<FqP>(b, b, f);
        }
    }
    if (dlog & 1u) return false;
    uint32_t s = (0x80000000u - (dlog >> 1)) & 0x7fffffffu;
    mont_mul<FqP>(x, x, num);
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        fe_load_tab(f, &t.pos[k][(s >> (8 * k)) & 255u]);
        mont_mul<FqP>(x, x, f);
    }
    r = x;
    return true;
}

}  // namespace jj
