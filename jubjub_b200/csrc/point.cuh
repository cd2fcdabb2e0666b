// point.cuh -- Jubjub twisted-Edwards group law on register-resident field elements.
//
// Coordinate systems and formulas are the reference's, verbatim in *sequence* so that
// point_add / point_double outputs are bit-identical 160-byte extended points
// (SURVEY.md section 8c).  Paths relative to /root/reference:
//   ext_point   (U, V, Z, T1, T2), T1*T2 = UV/Z        src/lib.rs:139-145
//   ext_niels   (V+U, V-U, Z, T1*T2*2d)                src/lib.rs:327-332
//   aff_niels   (v+u, v-u, u*v*2d)                     src/lib.rs:255-259
//   point_double      4S + 3M                          src/lib.rs:739-828
//   point_add_niels   8M  (sub: swapped halves)        src/lib.rs:883-940
//   point_add_aff_niels 7M                             src/lib.rs:944-988
//   into_extended     3M                               src/lib.rs:1052-1060
//   to_niels          2M                               src/lib.rs:728-735, 652-658
#pragma once
#include "fe.cuh"

namespace jj {

struct aff_point { fe u, v; };
struct ext_point { fe u, v, z, t1, t2; };
struct ext_niels { fe vpu, vmu, z, t2d; };
struct aff_niels { fe vpu, vmu, t2d; };

// Montgomery-form curve constants: d = -(10240/10241) (src/lib.rs:399-404), 2d (:407-412),
// generator (u, 11) (:1380-1396).  Values are from_raw(...) of the reference's raw limbs,
// recomputed in oracle/model.py and checked in tests/test_oracle_kat.py.
struct Curve {
    JJ_CONST_FN uint32_t D(int i) {
        constexpr uint32_t t[8] = {0xb974f6b0u, 0x2a522455u, 0x0d9acab3u, 0xfc6cc9efu, 0xc27628d1u, 0x7a08fb94u, 0xfe0e262eu, 0x57f8f6a8u};
        return t[i];
    }
    JJ_CONST_FN uint32_t D2(int i) {
        constexpr uint32_t t[8] = {0x72e9ed5fu, 0x54a448acu, 0x1b373967u, 0xa51befdbu, 0x7b4a799eu, 0xc0d81f21u, 0xd27ecf14u, 0x3c0445feu};
        return t[i];
    }
    JJ_CONST_FN uint32_t GEN_U(int i) {
        constexpr uint32_t t[8] = {0xc166eca5u, 0x50c87a58u, 0xc0051afcu, 0x8046fd74u, 0x695b0493u, 0x406355eeu, 0x1bdc7e0au, 0x0d5a8d93u};
        return t[i];
    }
    JJ_CONST_FN uint32_t GEN_V(int i) {
        constexpr uint32_t t[8] = {0xffffffe8u, 0x00000017u, 0x00276018u, 0x26389fb8u, 0x18d3bf80u, 0x3293bf3fu, 0x193c413bu, 0x21b85034u};
        return t[i];
    }
};

// Fq product / square as used by the point formulas.  On the device they are real (noinline)
// functions: arguments and result travel in registers (no stack frame), and every point formula
// shares ONE copy of the product and squaring bodies (119 / 91 multiplier instructions).  Fully inlined, the scalar-mul loop body is
// 53 KB of straight-line code, which misses the instruction cache as soon as more than 8 warps per
// SM run it (ncu: stall_no_instruction 1.7 per issue at 16 warps); with shared bodies the loop is
// ~15 KB and 12-16 warps/SM keep the multiplier pipe busy.
#if defined(JJ_HOST_EMUL) || defined(JJ_INLINE_FQ)
JJ_DEVICE void fq_mul(fe& r, const fe& a, const fe& b) { mont_mul<FqP>(r, a, b); }
JJ_DEVICE void fq_sqr(fe& r, const fe& a) { mont_sqr<FqP>(r, a); }
#else
#if !defined(JJ_FQ_CALL_BY_POINTER)
static __device__ __noinline__ fe fq_mul_body(fe a, fe b) {
    fe r;
    mont_mul<FqP>(r, a, b);
    return r;
}
static __device__ __noinline__ fe fq_sqr_body(fe a) {
    fe r;
    mont_sqr<FqP>(r, a);
    return r;
}
JJ_DEVICE void fq_mul(fe& r, const fe& a, const fe& b) { r = fq_mul_body(a, b); }
JJ_DEVICE void fq_sqr(fe& r, const fe& a) { r = fq_sqr_body(a); }
#else
// Operands travel through the thread's local-memory frame (L1-resident, 128-bit LDL/STL on the
// otherwise idle LSU pipe) instead of ~24 register moves per call: ptxas turns half of such
// moves into IMAD.MOV on the FMA-heavy pipe, the very pipe the multiplier saturates.
static __device__ __noinline__ void fq_mul_body(fe* r, const fe* a, const fe* b) {
    fe x = *a, y = *b, z;
    mont_mul<FqP>(z, x, y);
    *r = z;
}
static __device__ __noinline__ void fq_sqr_body(fe* r, const fe* a) {
    fe x = *a, z;
    mont_sqr<FqP>(z, x);
    *r = z;
}
JJ_DEVICE void fq_mul(fe& r, const fe& a, const fe& b) { fq_mul_body(&r, &a, &b); }
JJ_DEVICE void fq_sqr(fe& r, const fe& a) { fq_sqr_body(&r, &a); }
#endif
#endif

#define JJ_LOAD_CONST(r, ACCESSOR)                          \
    do {                                                   \
        _Pragma("unroll") for (int i_ = 0; i_ < 8; i_++)(r).w[i_] = ACCESSOR(i_); \
    } while (0)

JJ_DEVICE void point_set_identity(ext_point& p) {  // (0, 1, 1, 0, 0)  src/lib.rs:680-688
    fe_set_zero(p.u);
    fe_set_one<FqP>(p.v);
    fe_set_one<FqP>(p.z);
    fe_set_zero(p.t1);
    fe_set_zero(p.t2);
}
JJ_DEVICE void point_from_affine(ext_point& p, const aff_point& a) {  // src/lib.rs:214-226
    p.u = a.u;
    p.v = a.v;
    fe_set_one<FqP>(p.z);
    p.t1 = a.u;
    p.t2 = a.v;
}
JJ_DEVICE void point_neg(ext_point& r, const ext_point& p) {  // src/lib.rs:196-210
    fe_neg<FqP>(r.u, p.u);
    r.v = p.v;
    r.z = p.z;
    fe_neg<FqP>(r.t1, p.t1);
    r.t2 = p.t2;
}

// Every formula exists in two instantiations: INL = true inlines the Fq product / square bodies
// (best for kernels that use a formula once: the elementwise point kernels, the fixed-base loop, the
// doubling of the variable-base loop), INL = false calls the shared noinline bodies (small code).
template <bool INL, bool M1 = kM1MulDefault>
JJ_DEVICE void fqm(fe& r, const fe& a, const fe& b) {
    if (INL) mont_mul<FqP, M1>(r, a, b);
    else fq_mul(r, a, b);
}
template <bool INL>
JJ_DEVICE void fqs(fe& r, const fe& a) {
    if (INL) mont_sqr<FqP>(r, a);
    else fq_sqr(r, a);
}
#if defined(JJ_DOUBLE_INLINE)
constexpr bool kInlineDouble = true;
#else
constexpr bool kInlineDouble = false;
#endif

// completed point (u, v, z, t) -> extended (u*t, v*z, z*t, u, v)   (src/lib.rs:1052-1060)
template <bool INL, bool M1 = kM1MulDefault>
JJ_DEVICE void into_extended_t(ext_point& r, const fe& cu, const fe& cv, const fe& cz, const fe& ct) {
    fe u, v, z;
    fqm<INL, M1>(u, cu, ct);
    fqm<INL, M1>(v, cv, cz);
    fqm<INL, M1>(z, cz, ct);
    r.t1 = cu;
    r.t2 = cv;
    r.u = u;
    r.v = v;
    r.z = z;
}
// JJ_DBL_ORDER: source order of the doubling's independent pieces (same values, same 160 output bytes).  ptxas schedules
// the inlined doubling as one basic block but largely keeps the PTX order between independent groups, so the order
// decides which modular add/sub runs (pure ALU work) can hide under which products.  Measured per 2^20 scalar-muls on one
// box: order 0 (round 1) 33.62 ms, order 1 33.49 / 33.43 ms, order 2 33.71 ms -- small, but order 1 is the default.
#ifndef JJ_DBL_ORDER
#define JJ_DBL_ORDER 1
#endif
template <bool INL>
JJ_DEVICE void point_double_t(ext_point& r, const ext_point& p) {
    fe uu, vv, zz2, uv2, vpu, vmu, cu, ct;
    // the reference's values (src/lib.rs:812-826)
#if JJ_DBL_ORDER == 0
    // ordered so that u, v and the squares die as early as possible: fewer live registers across the four squarings
    fe_add<FqP>(uv2, p.u, p.v);
    fqs<INL>(uu, p.u);
    fqs<INL>(vv, p.v);
    fqs<INL>(uv2, uv2);
    fe_add<FqP>(vpu, vv, uu);
    fe_sub<FqP>(vmu, vv, uu);
    fe_sub<FqP>(cu, uv2, vpu);
    fqs<INL>(zz2, p.z);
#if defined(JJ_NO_LAZY_ADD)
    fe_dbl<FqP>(zz2, zz2);
#else
    // 2z^2 stays unreduced in [0, 2q): it is only the minuend of the next subtraction, whose result ct in [0, 2q) is only
    // ever the second operand of a product (u*t and z*t below) -- 17 ALU instructions fewer, same 160 output bytes
    fe_add_lazy<FqP>(zz2, zz2, zz2);
#endif
    fe_sub<FqP>(ct, zz2, vmu);
    into_extended_t<INL>(r, cu, vpu, vmu, ct);
#elif JJ_DBL_ORDER == 1
    // v^2 + u^2 and v^2 - u^2 sit in front of the two squarings that do not need them; cu, ct in front of v*z
    fe_add<FqP>(uv2, p.u, p.v);
    fqs<INL>(uu, p.u);
    fqs<INL>(vv, p.v);
    fe_add<FqP>(vpu, vv, uu);
    fe_sub<FqP>(vmu, vv, uu);
    fqs<INL>(uv2, uv2);
    fqs<INL>(zz2, p.z);
    fe_sub<FqP>(cu, uv2, vpu);
    fe_add_lazy<FqP>(zz2, zz2, zz2);
    fe_sub<FqP>(ct, zz2, vmu);
    {
        fe u, v, z;
        fqm<INL>(v, vpu, vmu);
        fqm<INL>(u, cu, ct);
        fqm<INL>(z, vmu, ct);
        r.t1 = cu; r.t2 = vpu; r.u = u; r.v = v; r.z = z;
    }
#else
    // z^2 first; the product v*z = (v^2+u^2)(v^2-u^2) as soon as its operands exist
    fqs<INL>(zz2, p.z);
    fe_add<FqP>(uv2, p.u, p.v);
    fqs<INL>(uu, p.u);
    fqs<INL>(vv, p.v);
    fe_add_lazy<FqP>(zz2, zz2, zz2);
    fe_add<FqP>(vpu, vv, uu);
    fe_sub<FqP>(vmu, vv, uu);
    fqs<INL>(uv2, uv2);
    fe_sub<FqP>(ct, zz2, vmu);
    {
        fe u, v, z;
        fqm<INL>(v, vpu, vmu);
        fe_sub<FqP>(cu, uv2, vpu);
        fqm<INL>(z, vmu, ct);
        fqm<INL>(u, cu, ct);
        r.t1 = cu; r.t2 = vpu; r.u = u; r.v = v; r.z = z;
    }
#endif
}
// p + n (sub = false) or p - n (sub = true); Z2 = nullptr means an affine-Niels operand (d = 2z).
// PRESWAPPED: the caller has already exchanged n_vpu / n_vmu for a subtraction (the scalar-mul cores do it by address
// when they read the window table: 16 SEL fewer per addition), so only the output side looks at `sub`.
template <bool INL, bool M1 = kM1MulDefault, bool PRESWAPPED = false>
JJ_DEVICE void point_add_core_t(ext_point& r, const ext_point& p, const fe& n_vpu, const fe& n_vmu, const fe* n_z,
                                const fe& n_t2d, bool sub) {
    fe a, b, c, d, t, n1, n2;
    fe_select(n1, n_vmu, n_vpu, PRESWAPPED ? false : sub);  // multiplies (v - u)
    fe_select(n2, n_vpu, n_vmu, PRESWAPPED ? false : sub);  // multiplies (v + u)
    fe_sub<FqP>(t, p.v, p.u);
    fqm<INL, M1>(a, t, n1);
#if defined(JJ_NO_LAZY_ADD)
    fe_add<FqP>(t, p.v, p.u);
    fqm<INL, M1>(b, t, n2);
#else
    fe_add_lazy<FqP>(t, p.v, p.u);  // v + u in [0, 2q): second operand of the product (the Niels field is canonical)
    fqm<INL, M1>(b, n2, t);
#endif
    fqm<INL, M1>(c, p.t1, p.t2);
    fqm<INL, M1>(c, c, n_t2d);
    if (n_z) {
#if defined(JJ_NO_LAZY_ADD)
        fqm<INL, M1>(d, p.z, *n_z);
        fe_dbl<FqP>(d, d);
#else
        fe_add_lazy<FqP>(d, p.z, p.z);  // d = (2z) * n.z with the doubling done before the product, unreduced
        fqm<INL, M1>(d, *n_z, d);
#endif
    } else {
        fe_dbl<FqP>(d, p.z);
    }
    fe cu, cv, dpc, dmc, cz, ct;
    fe_sub<FqP>(cu, b, a);
    fe_add<FqP>(cv, b, a);
    fe_add<FqP>(dpc, d, c);
    fe_sub<FqP>(dmc, d, c);
    fe_select(cz, dpc, dmc, sub);
    fe_select(ct, dmc, dpc, sub);
    into_extended_t<INL, M1>(r, cu, cv, cz, ct);
}
template <bool INL, bool PRESWAPPED = false>
JJ_DEVICE void point_add_niels_t(ext_point& r, const ext_point& p, const ext_niels& n, bool sub) {
    point_add_core_t<INL, kM1MulDefault, PRESWAPPED>(r, p, n.vpu, n.vmu, &n.z, n.t2d, sub);
}
template <bool INL, bool M1 = kM1MulDefault, bool PRESWAPPED = false>
JJ_DEVICE void point_add_aff_niels_t(ext_point& r, const ext_point& p, const aff_niels& n, bool sub) {
    point_add_core_t<INL, M1, PRESWAPPED>(r, p, n.vpu, n.vmu, nullptr, n.t2d, sub);
}
template <bool INL>
JJ_DEVICE void point_to_niels_t(ext_niels& n, const ext_point& p) {
    fe d2, t;
    JJ_LOAD_CONST(d2, Curve::D2);
    fe_add<FqP>(n.vpu, p.v, p.u);
    fe_sub<FqP>(n.vmu, p.v, p.u);
    n.z = p.z;
    fqm<INL>(t, p.t1, p.t2);
    fqm<INL>(n.t2d, t, d2);
}
template <bool INL>
JJ_DEVICE void affine_to_niels_t(aff_niels& n, const aff_point& p) {
    fe d2, t;
    JJ_LOAD_CONST(d2, Curve::D2);
    fe_add<FqP>(n.vpu, p.v, p.u);
    fe_sub<FqP>(n.vmu, p.v, p.u);
    fqm<INL>(t, p.u, p.v);
    fqm<INL>(n.t2d, t, d2);
}
// p + q with both extended: q.to_niels() then the 8M add (src/lib.rs:992-999).
template <bool INL>
JJ_DEVICE void point_add_t(ext_point& r, const ext_point& p, const ext_point& q, bool sub) {
    ext_niels n;
    point_to_niels_t<INL>(n, q);
    point_add_niels_t<INL>(r, p, n, sub);
}
// default instantiations used by the scalar-mul cores
JJ_DEVICE void point_double(ext_point& r, const ext_point& p) { point_double_t<kInlineDouble>(r, p); }
JJ_DEVICE void point_add_niels(ext_point& r, const ext_point& p, const ext_niels& n, bool sub) { point_add_niels_t<false>(r, p, n, sub); }
JJ_DEVICE void point_add_niels_preswapped(ext_point& r, const ext_point& p, const ext_niels& n, bool sub) { point_add_niels_t<false, true>(r, p, n, sub); }
JJ_DEVICE void point_add_aff_niels(ext_point& r, const ext_point& p, const aff_niels& n, bool sub) { point_add_aff_niels_t<false>(r, p, n, sub); }
JJ_DEVICE void point_to_niels(ext_niels& n, const ext_point& p) { point_to_niels_t<false>(n, p); }
JJ_DEVICE void affine_to_niels(aff_niels& n, const aff_point& p) { affine_to_niels_t<false>(n, p); }
JJ_DEVICE void point_add(ext_point& r, const ext_point& p, const ext_point& q, bool sub) { point_add_t<false>(r, p, q, sub); }

// AffinePoint::from_bytes_inner / batch_from_bytes (src/lib.rs:492-534, 541-627) for one encoding held as 8 LE words:
// u^2 = (v^2 - 1) / (1 + d v^2), u = sqrt(u^2) with the sign fixed from the parity of the canonical u, ZIP-216 rejection
// of the non-canonical encodings of (0, +-1).  The reference inverts the denominator (one batched inversion in
// batch_from_bytes, :596-600) and then takes the root; here the root of the quotient comes out of ONE power of
// num * den (fq_sqrt_ratio), so decoding needs no inversion, no scratch and no chain across encodings.
// Returns false (and leaves the zero point) for non-canonical v, off-curve v, or -- with zip216 --
// the non-canonical encodings of (0, +-1) whose sign bit is set (ZIP 216, src/lib.rs:527-531).
JJ_DEVICE bool point_from_bytes(aff_point& p, const fe& enc, bool zip216) {
    fe vraw = enc, v, v2, num, den, u, un, uc, one, d;
    const uint32_t sign = vraw.w[7] >> 31;
    vraw.w[7] &= 0x7fffffffu;
    fe_set_zero(p.u);
    fe_set_zero(p.v);
    if (!fe_is_canonical<FqP>(vraw)) return false;
    fe_from_raw<FqP>(v, vraw);
    fe_set_one<FqP>(one);
    JJ_LOAD_CONST(d, Curve::D);
    fq_sqr(v2, v);
    fe_sub<FqP>(num, v2, one);          // v^2 - 1
    fq_mul(den, d, v2);
    fe_add<FqP>(den, one, den);         // 1 + d v^2 (never zero: -1/d is a non-residue)
    if (!fq_sqrt_ratio(u, num, den)) return false;
    fe_to_canonical<FqP>(uc, u);
    const bool flip = ((uc.w[0] ^ sign) & 1u) != 0;
    fe_neg<FqP>(un, u);
    if (zip216 && fe_is_zero(u) && flip) return false;
    fe_select(p.u, u, un, flip);
    p.v = v;
    return true;
}
JJ_DEVICE bool point_is_identity(const ext_point& p) {  // u == 0 and v == z  src/lib.rs:691-696
    return fe_is_zero(p.u) && fe_eq(p.v, p.z);
}

}  // namespace jj
