// slotmul.cuh -- variable-base scalar multiplication on a shared-memory "slot file".
//
// Second mapping of the same algorithm as scalarmul.cuh (signed radix-16 window, reference
// formulas).  There, the accumulator point lives in registers and the formulas are inlined; the
// unrolled loop body is ~50 KB and only 8-12 warps/SM can run before registers or the
// instruction cache give out, while calling shared Fq bodies costs ~20 register moves per call
// that ptxas schedules onto the FMA-heavy pipe -- the pipe IMAD.WIDE already saturates.
//
// Here every field element of the working set lives in a per-thread slot of shared memory
// ([slot][half][lane] 16-byte units: conflict-free for uniform AND for per-lane slot indices), and
// each field operation is one small noinline routine taking packed slot numbers: load operands
// (4 x LDS.128), compute in registers, store (2 x STS.128).  Code is ~12 KB, ~64 registers per
// thread, so 16+ warps/SM hide the multiplier latency; operand traffic rides the idle LSU pipe.
// Negative digits cost nothing: the Niels halves and the d+c / d-c outputs are swapped by
// choosing slot numbers per lane (the reference's subtraction formula, src/lib.rs:922-940).
#pragma once
#include "scalarmul.cuh"

namespace jj {

// Round 1's recoding, kept for this experimental mapping: K = ((k mod 2^252) + 0x0888...8) << 3, so that bit 255 of K is the top
// carry d63 and, after one more left shift, the top nibble is digit 62 (all 63 digits signed, four doublings in every pass).
JJ_DEVICE void recode_scalar_r1(uint32_t K[8], const uint32_t k[8]) {
    uint32_t t[8];
    add_cc(t[0], k[0], 0x88888888u);
#pragma unroll
    for (int i = 1; i < 7; i++) addc_cc(t[i], k[i], 0x88888888u);
    addc(t[7], k[7] & 0x0fffffffu, 0x08888888u);
#pragma unroll
    for (int i = 7; i > 0; i--) K[i] = (t[i] << 3) | (t[i - 1] >> 29);
    K[0] = t[0] << 3;
}

#if defined(JJ_HOST_EMUL)
// host emulation: one thread's slot file is a plain array of 8-word slots
struct SlotFile {
    uint32_t* base;
    JJ_DEVICE_SPEC void ld(fe& f, uint32_t s) const {
        for (int i = 0; i < 8; i++) f.w[i] = base[s * 8 + i];
    }
    JJ_DEVICE_SPEC void st(uint32_t s, const fe& f) const {
        for (int i = 0; i < 8; i++) base[s * 8 + i] = f.w[i];
    }
};
#define JJ_SLOT_FN static inline
#else
struct SlotFile {
    uint32_t base;  // shared-state-space byte address of (warp slab + lane * 16)
    __device__ __forceinline__ void ld(fe& f, uint32_t s) const {
        uint32_t a = base + s * 1024u;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(f.w[0]), "=r"(f.w[1]), "=r"(f.w[2]), "=r"(f.w[3]) : "r"(a));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+512];"
                     : "=r"(f.w[4]), "=r"(f.w[5]), "=r"(f.w[6]), "=r"(f.w[7]) : "r"(a));
    }
    __device__ __forceinline__ void st(uint32_t s, const fe& f) const {
        uint32_t a = base + s * 1024u;
        asm volatile("st.shared.v4.u32 [%4], {%0,%1,%2,%3};" ::"r"(f.w[0]), "r"(f.w[1]), "r"(f.w[2]), "r"(f.w[3]), "r"(a) : "memory");
        asm volatile("st.shared.v4.u32 [%4+512], {%0,%1,%2,%3};" ::"r"(f.w[4]), "r"(f.w[5]), "r"(f.w[6]), "r"(f.w[7]), "r"(a) : "memory");
    }
};
#define JJ_SLOT_FN static __device__ __noinline__
#endif

// slot numbers
enum Slot : uint32_t {
    S_U = 0, S_V, S_Z, S_T1, S_T2,   // accumulator (U, V, Z, T1, T2)
    S_A, S_B, S_C, S_D,              // temporaries
    S_N0, S_N1, S_N2, S_N3,          // Niels operand (v+u, v-u, z, t2d)
    S_COUNT
};
JJ_HD uint32_t ops3(uint32_t d, uint32_t a, uint32_t b) { return d | (a << 8) | (b << 16); }
JJ_HD uint32_t ops4(uint32_t d, uint32_t e, uint32_t a, uint32_t b) { return d | (a << 8) | (b << 16) | (e << 24); }

// d = a * b
JJ_SLOT_FN void s_mul(SlotFile S, uint32_t ops) {
    fe x, y, r;
    S.ld(x, (ops >> 8) & 0xff);
    S.ld(y, (ops >> 16) & 0xff);
    mont_mul<FqP>(r, x, y);
    S.st(ops & 0xff, r);
}
// d = a^2
JJ_SLOT_FN void s_sqr(SlotFile S, uint32_t ops) {
    fe x, r;
    S.ld(x, (ops >> 8) & 0xff);
    mont_sqr<FqP>(r, x);
    S.st(ops & 0xff, r);
}
// mode 0: d = a + b;  1: d = a - b;  2: d = 2a - b;  3: d = a + b and e = a - b;  4: d = 2a
JJ_SLOT_FN void s_lin(SlotFile S, uint32_t ops, int mode) {
    fe x, y, r;
    S.ld(x, (ops >> 8) & 0xff);
    if (mode != 4) S.ld(y, (ops >> 16) & 0xff);
    if (mode == 0) {
        fe_add<FqP>(r, x, y);
    } else if (mode == 1) {
        fe_sub<FqP>(r, x, y);
    } else if (mode == 2) {
        fe_dbl<FqP>(x, x);
        fe_sub<FqP>(r, x, y);
    } else if (mode == 3) {
        fe t;
        fe_sub<FqP>(t, x, y);
        S.st(ops >> 24, t);
        fe_add<FqP>(r, x, y);
    } else {
        fe_dbl<FqP>(r, x);
    }
    S.st(ops & 0xff, r);
}

// acc <- 2 * acc   (src/lib.rs:812-827 + into_extended :1052-1060), 4S + 3M
JJ_DEVICE void sp_double(SlotFile S) {
    s_sqr(S, ops3(S_A, S_U, 0));                 // A = uu
    s_sqr(S, ops3(S_B, S_V, 0));                 // B = vv
    s_sqr(S, ops3(S_C, S_Z, 0));                 // C = zz
    s_lin(S, ops3(S_D, S_U, S_V), 0);            // D = u + v
    s_sqr(S, ops3(S_D, S_D, 0));                 // D = uv2
    s_lin(S, ops4(S_T2, S_B, S_B, S_A), 3);      // T2 = vv + uu (completed v), B = vv - uu (completed z)
    s_lin(S, ops3(S_T1, S_D, S_T2), 1);          // T1 = uv2 - (vv + uu) (completed u)
    s_lin(S, ops3(S_C, S_C, S_B), 2);            // C = 2 zz - (vv - uu) (completed t)
    s_mul(S, ops3(S_U, S_T1, S_C));              // U = u * t
    s_mul(S, ops3(S_V, S_T2, S_B));              // V = v * z
    s_mul(S, ops3(S_Z, S_B, S_C));               // Z = z * t
}
// acc <- acc +/- N, N = slots (N0..N3) (src/lib.rs:905-918 / :927-938), 8M
JJ_DEVICE void sp_add_niels(SlotFile S, bool sub) {
    const uint32_t n_vmu = sub ? S_N0 : S_N1, n_vpu = sub ? S_N1 : S_N0;
    s_lin(S, ops4(S_A, S_B, S_V, S_U), 3);       // A = v + u, B = v - u
    s_mul(S, ops3(S_B, S_B, n_vmu));             // B = a = (v - u) * n.vmu
    s_mul(S, ops3(S_A, S_A, n_vpu));             // A = b = (v + u) * n.vpu
    s_mul(S, ops3(S_C, S_T1, S_T2));
    s_mul(S, ops3(S_C, S_C, S_N3));              // C = c = t1 t2 * n.t2d
    s_mul(S, ops3(S_D, S_Z, S_N2));
    s_lin(S, ops3(S_D, S_D, 0), 4);              // D = d = 2 z n.z
    s_lin(S, ops4(S_T2, S_T1, S_A, S_B), 3);     // T2 = b + a (completed v), T1 = b - a (completed u)
    s_lin(S, ops4(S_A, S_B, S_D, S_C), 3);       // A = d + c, B = d - c
    const uint32_t cz = sub ? S_B : S_A, ct = sub ? S_A : S_B;
    s_mul(S, ops3(S_U, S_T1, ct));               // U = u * t
    s_mul(S, ops3(S_V, S_T2, cz));               // V = v * z
    s_mul(S, ops3(S_Z, cz, ct));                 // Z = z * t
}
// (N0..N3) <- to_niels(acc)   (src/lib.rs:728-735)
JJ_DEVICE void sp_to_niels(SlotFile S) {
    fe d2;
    JJ_LOAD_CONST(d2, Curve::D2);
    S.st(S_D, d2);
    s_lin(S, ops4(S_N0, S_N1, S_V, S_U), 3);     // N0 = v + u, N1 = v - u
    fe z;
    S.ld(z, S_Z);
    S.st(S_N2, z);
    s_mul(S, ops3(S_N3, S_T1, S_T2));
    s_mul(S, ops3(S_N3, S_N3, S_D));
}

// acc slots <- [k] P.  P is read from the accumulator slots (U, V, Z, T1, T2); the window table
// goes through `tbl` (same policies as scalarmul.cuh).  Leaves the result in the accumulator slots.
template <class Table>
JJ_DEVICE void scalar_mul_slots(SlotFile S, const uint32_t k[8], Table& tbl) {
    ext_point P;
    S.ld(P.u, S_U); S.ld(P.v, S_V); S.ld(P.z, S_Z); S.ld(P.t1, S_T1); S.ld(P.t2, S_T2);
    // table: entry j = niels((j+1) P); the first entry stays in a spare copy for the additions
    ext_niels n1;
    sp_to_niels(S);
    S.ld(n1.vpu, S_N0); S.ld(n1.vmu, S_N1); S.ld(n1.z, S_N2); S.ld(n1.t2d, S_N3);
    tbl.store(0, n1);
#pragma unroll 1
    for (int j = 1; j < 8; j++) {
        S.st(S_N0, n1.vpu); S.st(S_N1, n1.vmu); S.st(S_N2, n1.z); S.st(S_N3, n1.t2d);
        sp_add_niels(S, false);       // acc = (j+1) P
        sp_to_niels(S);
        ext_niels nj;
        S.ld(nj.vpu, S_N0); S.ld(nj.vmu, S_N1); S.ld(nj.z, S_N2); S.ld(nj.t2d, S_N3);
        tbl.store(j, nj);
    }
    uint32_t K[8];
    recode_scalar_r1(K, k);
    {
        ext_point id;
        point_set_identity(id);
        bool top = (K[7] >> 31) != 0;
        fe t;
        fe_select(t, id.u, P.u, top);   S.st(S_U, t);
        fe_select(t, id.v, P.v, top);   S.st(S_V, t);
        fe_select(t, id.z, P.z, top);   S.st(S_Z, t);
        fe_select(t, id.t1, P.t1, top); S.st(S_T1, t);
        fe_select(t, id.t2, P.t2, top); S.st(S_T2, t);
        shl_256(K, 1);
    }
#pragma unroll 1
    for (int i = 62; i >= 0; i--) {
#pragma unroll 1
        for (int j = 0; j < 4; j++) sp_double(S);
        int d = (int)(K[7] >> 28) - 8;
        shl_256(K, 4);
        if (d != 0) {
            ext_niels n;
            tbl.load((d < 0 ? -d : d) - 1, n);
            S.st(S_N0, n.vpu); S.st(S_N1, n.vmu); S.st(S_N2, n.z); S.st(S_N3, n.t2d);
            sp_add_niels(S, d < 0);
        }
    }
}

}  // namespace jj
