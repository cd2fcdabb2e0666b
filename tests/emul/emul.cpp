// Host emulation harness -- TEST INFRASTRUCTURE ONLY.
// Compiles the *kernel arithmetic source* (jubjub_b200/csrc/*.cuh) with a plain C++
// compiler, the PTX carry-chain primitives replaced by C emulation (JJ_HOST_EMUL), so
// the limb-level algorithms can be unit-tested against the oracle without a GPU.
// Nothing in the product loads this; it is not a CPU fallback.
#define JJ_HOST_EMUL 1
#include <cstddef>
#include <cstring>
#include "../../jubjub_b200/csrc/fe.cuh"

using namespace jj;

template <class F>
static void fe_op(int op, const fe* a, const fe* b, fe* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        fe r;
        switch (op) {
            case 0: mont_mul<F>(r, a[i], b[i]); break;
            case 1: mont_sqr<F>(r, a[i]); break;
            case 2: fe_add<F>(r, a[i], b[i]); break;
            case 3: fe_sub<F>(r, a[i], b[i]); break;
            case 4: fe_neg<F>(r, a[i]); break;
            case 5: fe_dbl<F>(r, a[i]); break;
            case 6: fe_invert<F>(r, a[i]); break;
            case 7: fe_to_canonical<F>(r, a[i]); break;
            case 8: fe_from_raw<F>(r, a[i]); break;
            case 10: mont_mul<F, false>(r, a[i], b[i]); break;  // k*m1 on the ALU pipe (elementwise kernels)
            case 11: mont_sqr<F, false>(r, a[i]); break;
            default: fe_set_zero(r); r.w[0] = fe_is_canonical<F>(a[i]); break;
        }
        out[i] = r;
    }
}
extern "C" void emul_fe_op(int which, int op, const void* a, const void* b, void* out, size_t n) {
    if (which == 0) fe_op<FqP>(op, (const fe*)a, (const fe*)b, (fe*)out, n);
    else fe_op<FrP>(op, (const fe*)a, (const fe*)b, (fe*)out, n);
}

#include "../../jubjub_b200/csrc/slotmul.cuh"
#include "../../jubjub_b200/csrc/torsion.cuh"

// pairing-based is_torsion_free (torsion.cuh) on ExtendedPoint inputs
extern "C" void emul_is_torsion_free(const void* p_, uint8_t* out, size_t n) {
    const ext_point* p = (const ext_point*)p_;
    for (size_t i = 0; i < n; i++) out[i] = point_is_torsion_free(p[i].u, p[i].v, p[i].z) ? 1 : 0;
}

// op: 0 double, 1 add (ext+ext), 2 sub (ext-ext), 3 add ext-niels, 4 sub ext-niels,
//     5 add affine-niels, 6 sub affine-niels, 7 to_niels (ext), 8 to_niels (affine), 9 neg
extern "C" void emul_point_op(int op, const void* p_, const void* q_, void* out_, size_t n) {
    const ext_point* p = (const ext_point*)p_;
    for (size_t i = 0; i < n; i++) {
        ext_point r;
        switch (op) {
            case 0: point_double(r, p[i]); break;
            case 1: point_add(r, p[i], ((const ext_point*)q_)[i], false); break;
            case 2: point_add(r, p[i], ((const ext_point*)q_)[i], true); break;
            case 3: point_add_niels(r, p[i], ((const ext_niels*)q_)[i], false); break;
            case 4: point_add_niels(r, p[i], ((const ext_niels*)q_)[i], true); break;
            case 5: point_add_aff_niels(r, p[i], ((const aff_niels*)q_)[i], false); break;
            case 6: point_add_aff_niels(r, p[i], ((const aff_niels*)q_)[i], true); break;
            case 7: point_to_niels(((ext_niels*)out_)[i], p[i]); continue;
            case 8: affine_to_niels(((aff_niels*)out_)[i], ((const aff_point*)p_)[i]); continue;
            default: point_neg(r, p[i]); break;
        }
        ((ext_point*)out_)[i] = r;
    }
}
extern "C" void emul_scalar_mul(const void* p_, const void* k_, void* out_, size_t n) {
    for (size_t i = 0; i < n; i++) {
        LocalTable tbl;
        ext_point acc;
        scalar_mul_core(acc, ((const ext_point*)p_)[i], ((const uint32_t*)k_) + 8 * i, tbl);
        ((ext_point*)out_)[i] = acc;
    }
}
// the constant-time mode of the same core (table scan, selects, every addition executed)
extern "C" void emul_scalar_mul_ct(const void* p_, const void* k_, void* out_, size_t n) {
    for (size_t i = 0; i < n; i++) {
        LocalTable tbl;
        ext_point acc;
        scalar_mul_core<LocalTable, true>(acc, ((const ext_point*)p_)[i], ((const uint32_t*)k_) + 8 * i, tbl);
        ((ext_point*)out_)[i] = acc;
    }
}
// one scalar k (32 LE bytes) for all points: the width-5 NAF path of is_torsion_free
extern "C" int emul_scalar_mul_wnaf(const void* p_, const void* k_, void* out_, size_t n) {
    NafDigits naf;
    wnaf5_recode(naf, (const uint32_t*)k_);
    for (size_t i = 0; i < n; i++) {
        LocalTable tbl;
        ext_point acc;
        scalar_mul_wnaf_core(acc, ((const ext_point*)p_)[i], naf, tbl);
        ((ext_point*)out_)[i] = acc;
    }
    int nz = 0;
    for (int i = 0; i < 256; i++) nz += naf.d[i] != 0;
    return nz;
}
// fixed-base table for window width W (4 or 7), same construction as k_fixed_table_build
template <int W>
static void fixed_table(const void* base_affine, uint32_t* table, int first, int count) {
    using G = FixedGeom<W>;
    ext_point B;
    point_from_affine(B, *(const aff_point*)base_affine);
    for (int e = first; e < G::ENTRIES && e < first + count; e++) {
        const bool top = e == G::NW * G::PER;
        int i = top ? G::NW - 1 : e / G::PER, j = top ? 0 : e % G::PER;
        uint32_t k[8] = {0};
        const int extra = (!top && i == G::NW - 1) ? G::EXCESS : 0;
        const int bit = W * i - extra;
        uint64_t v = (uint64_t)(j + 1) << (bit & 31);
        k[bit >> 5] = (uint32_t)v;
        if ((bit >> 5) + 1 < 8) k[(bit >> 5) + 1] = (uint32_t)(v >> 32);
        LocalTable tbl;
        ext_point acc;
        scalar_mul_core(acc, B, k, tbl);
        for (int d = 0; d < (top ? W : extra); d++) point_double(acc, acc);
        fe zi;
        fe_invert<FqP>(zi, acc.z);
        aff_point a;
        mont_mul<FqP>(a.u, acc.u, zi);
        mont_mul<FqP>(a.v, acc.v, zi);
        aff_niels nn;
        affine_to_niels(nn, a);
        std::memcpy(table + (size_t)e * 24, &nn, 96);
    }
}
extern "C" int emul_fixed_table_words(int w) {
    return (w == 4 ? FixedGeom<4>::ENTRIES : w == 7 ? FixedGeom<7>::ENTRIES : w == 16 ? FixedGeom<16>::ENTRIES : FixedGeom<12>::ENTRIES) * 24;
}
// builds entries [first, first + count) of the table in place
extern "C" void emul_fixed_table(const void* base_affine, uint32_t* table, int w, int first, int count) {
    if (w == 4) fixed_table<4>(base_affine, table, first, count);
    else if (w == 7) fixed_table<7>(base_affine, table, first, count);
    else if (w == 16) fixed_table<16>(base_affine, table, first, count);
    else fixed_table<12>(base_affine, table, first, count);
}
extern "C" void emul_scalar_mul_fixed(const uint32_t* table, const void* k_, void* out_, size_t n, int w) {
    fixed_table_view v{table};
    for (size_t i = 0; i < n; i++) {
        ext_point acc;
        if (w == 4) scalar_mul_fixed_core<4, true>(acc, ((const uint32_t*)k_) + 8 * i, v);
        else if (w == 7) scalar_mul_fixed_core<7, true>(acc, ((const uint32_t*)k_) + 8 * i, v);
        else if (w == 16) scalar_mul_fixed_core<16, true>(acc, ((const uint32_t*)k_) + 8 * i, v);
        else scalar_mul_fixed_core<12, true>(acc, ((const uint32_t*)k_) + 8 * i, v);
        ((ext_point*)out_)[i] = acc;
    }
}
extern "C" void emul_from_bytes(const void* in32, void* out_affine, uint8_t* ok, int zip216, size_t n) {
    for (size_t i = 0; i < n; i++) {
        aff_point p;
        ok[i] = point_from_bytes(p, ((const fe*)in32)[i], zip216 != 0) ? 1 : 0;
        ((aff_point*)out_affine)[i] = p;
    }
}
extern "C" void emul_fq_sqrt(const void* a, void* out, uint8_t* ok, size_t n) {
    for (size_t i = 0; i < n; i++) {
        fe r;
        fe_set_zero(r);
        ok[i] = fq_sqrt(r, ((const fe*)a)[i]) ? 1 : 0;
        ((fe*)out)[i] = r;
    }
}

extern "C" void emul_fq_sqrt_ratio(const void* num, const void* den, void* out, uint8_t* ok, size_t n) {
    for (size_t i = 0; i < n; i++) {
        fe r;
        ok[i] = fq_sqrt_ratio(r, ((const fe*)num)[i], ((const fe*)den)[i]) ? 1 : 0;
        ((fe*)out)[i] = r;
    }
}

extern "C" void emul_fq_sqrt_ts(const void* a, void* out, uint8_t* ok, size_t n) {  // the loop form, for cross-checks
    for (size_t i = 0; i < n; i++) {
        fe r;
        fe_set_zero(r);
        ok[i] = fq_sqrt_ts(r, ((const fe*)a)[i]) ? 1 : 0;
        ((fe*)out)[i] = r;
    }
}
extern "C" int emul_fq_sqrt_tables_ok() {
    static FqSqrtTables t;
    return fq_sqrt_tables_build(t) ? 1 : 0;
}

extern "C" void emul_scalar_mul_slots(const void* p_, const void* k_, void* out_, size_t n) {
    for (size_t i = 0; i < n; i++) {
        uint32_t slots[S_COUNT * 8];
        SlotFile S{slots};
        const ext_point& P = ((const ext_point*)p_)[i];
        S.st(S_U, P.u); S.st(S_V, P.v); S.st(S_Z, P.z); S.st(S_T1, P.t1); S.st(S_T2, P.t2);
        LocalTable tbl;
        scalar_mul_slots(S, ((const uint32_t*)k_) + 8 * i, tbl);
        ext_point r;
        S.ld(r.u, S_U); S.ld(r.v, S_V); S.ld(r.z, S_Z); S.ld(r.t1, S_T1); S.ld(r.t2, S_T2);
        ((ext_point*)out_)[i] = r;
    }
}
