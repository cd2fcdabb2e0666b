// kernels.cuh -- sm_100a kernels of the batched Jubjub engine.
//
// Mapping: one thread owns one unit (field op, point op or scalar-mul); units are
// independent (the reference's types are Copy values, src/lib.rs:80,138), so batches are
// grid-strided over a grid sized in multiples of the SM count.  Memory layout is the
// reference's AoS (32-byte field elements), which is exactly one 256-bit LDG/STG
// (LDG.E.ENL2.256) per field element per thread: a warp's access is 1 KB contiguous for
// field batches and 32-byte-sector exact for 160-byte points.
#pragma once
#include <cuda_runtime.h>

#include "scalarmul.cuh"
#include "torsion.cuh"
#if defined(JJ_EXPERIMENTS)
#include "slotmul.cuh"  // shared-memory slot-file mapping (measured slower): not in the default build
#endif

namespace jj {

// ---- 256-bit global accesses ------------------------------------------------------------
__device__ __forceinline__ void ld_fe(fe& r, const void* p) {
    asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]),
                   "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void st_fe(void* p, const fe& r) {
    asm volatile("st.global.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(r.w[0]), "r"(r.w[1]), "r"(r.w[2]),
                 "r"(r.w[3]), "r"(r.w[4]), "r"(r.w[5]), "r"(r.w[6]), "r"(r.w[7]), "l"(p)
                 : "memory");
}
// L2 eviction hints for the scalar-mul kernels: the per-thread window table is re-read 63 times and
// rewritten by every unit, so its lines are marked evict-last; the streamed inputs and outputs are
// marked evict-first so that 352 MB of them per launch do not push the table out to DRAM.
__device__ __forceinline__ void ld_fe_keep(fe& r, const void* p) {
    asm volatile("ld.global.L2::evict_last.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]),
                   "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void st_fe_keep(void* p, const fe& r) {
    asm volatile("st.global.L2::evict_last.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(r.w[0]), "r"(r.w[1]),
                 "r"(r.w[2]), "r"(r.w[3]), "r"(r.w[4]), "r"(r.w[5]), "r"(r.w[6]), "r"(r.w[7]), "l"(p)
                 : "memory");
}
__device__ __forceinline__ void ld_fe_stream(fe& r, const void* p) {
    asm volatile("ld.global.L2::evict_first.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]),
                   "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void st_fe_stream(void* p, const fe& r) {
    asm volatile("st.global.L2::evict_first.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(r.w[0]), "r"(r.w[1]),
                 "r"(r.w[2]), "r"(r.w[3]), "r"(r.w[4]), "r"(r.w[5]), "r"(r.w[6]), "r"(r.w[7]), "l"(p)
                 : "memory");
}
__device__ __forceinline__ void ld_ext_stream(ext_point& p, const void* base, size_t i) {
    const char* b = (const char*)base + i * 160;
    ld_fe_stream(p.u, b);
    ld_fe_stream(p.v, b + 32);
    ld_fe_stream(p.z, b + 64);
    ld_fe_stream(p.t1, b + 96);
    ld_fe_stream(p.t2, b + 128);
}
__device__ __forceinline__ void st_ext_stream(void* base, size_t i, const ext_point& p) {
    char* b = (char*)base + i * 160;
    st_fe_stream(b, p.u);
    st_fe_stream(b + 32, p.v);
    st_fe_stream(b + 64, p.z);
    st_fe_stream(b + 96, p.t1);
    st_fe_stream(b + 128, p.t2);
}
__device__ __forceinline__ void ld_ext(ext_point& p, const void* base, size_t i) {
    const char* b = (const char*)base + i * 160;
    ld_fe(p.u, b);
    ld_fe(p.v, b + 32);
    ld_fe(p.z, b + 64);
    ld_fe(p.t1, b + 96);
    ld_fe(p.t2, b + 128);
}
__device__ __forceinline__ void st_ext(void* base, size_t i, const ext_point& p) {
    char* b = (char*)base + i * 160;
    st_fe(b, p.u);
    st_fe(b + 32, p.v);
    st_fe(b + 64, p.z);
    st_fe(b + 96, p.t1);
    st_fe(b + 128, p.t2);
}

// ---- field batches ------------------------------------------------------------------------
enum FeOp {
    FE_MUL = 0, FE_SQR, FE_ADD, FE_SUB, FE_NEG, FE_DBL, FE_INV, FE_TO_BYTES, FE_FROM_BYTES,
    FE_FROM_WIDE, FE_STREAM, FE_SQRT
};

__device__ __forceinline__ uint64_t splitmix_at(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// d0*R2 + d1*R3 (src/fr.rs:325-343)
template <class F>
__device__ __forceinline__ void fe_from_wide(fe& r, const fe& d0, const fe& d1) {
    fe r2, r3, x, y;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r2.w[i] = F::R2(i);
        r3.w[i] = F::R3(i);
    }
    mont_mul<F>(x, r2, d0);
    mont_mul<F>(y, r3, d1);
    fe_add<F>(r, x, y);
}

// CANON: inputs are canonical integers (converted with from_raw, which also reduces values
// >= m), outputs are converted back with to_canonical.
template <class F, int OP, bool CANON>
__global__ void __launch_bounds__(256) k_fe_op(const char* __restrict__ a, const char* __restrict__ b,
                                               char* __restrict__ out, uint8_t* __restrict__ ok, size_t n,
                                               uint64_t seed, size_t first) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        fe x, y, r;
        bool flag = true;
        if (OP == FE_STREAM) {
            fe d0, d1;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint64_t lo = splitmix_at(seed, (uint64_t)(first + i) * 8 + k);
                uint64_t hi = splitmix_at(seed, (uint64_t)(first + i) * 8 + 4 + k);
                d0.w[2 * k] = (uint32_t)lo;
                d0.w[2 * k + 1] = (uint32_t)(lo >> 32);
                d1.w[2 * k] = (uint32_t)hi;
                d1.w[2 * k + 1] = (uint32_t)(hi >> 32);
            }
            fe_from_wide<F>(r, d0, d1);
        } else if (OP == FE_FROM_WIDE) {
            ld_fe(x, a + i * 64);
            ld_fe(y, a + i * 64 + 32);
            fe_from_wide<F>(r, x, y);
        } else {
            ld_fe(x, a + i * 32);
            if (OP == FE_MUL || OP == FE_ADD || OP == FE_SUB) ld_fe(y, b + i * 32);
            if (CANON && OP != FE_FROM_BYTES) {
                fe_from_raw<F>(x, x);
                if (OP == FE_MUL || OP == FE_ADD || OP == FE_SUB) fe_from_raw<F>(y, y);
            }
            switch (OP) {
                case FE_MUL: mont_mul<F, false>(r, x, y); break;  // 112 multiplier instructions: stays HBM-bound (measured)
                case FE_SQR: mont_sqr<F>(r, x); break;            // issue-bound at 64 B/op: the default form is faster
                case FE_ADD: fe_add<F>(r, x, y); break;
                case FE_SUB: fe_sub<F>(r, x, y); break;
                case FE_NEG: fe_neg<F>(r, x); break;
                case FE_DBL: fe_dbl<F>(r, x); break;
                case FE_INV:
                    flag = !fe_is_zero(x);
                    fe_invert<F>(r, x);
                    break;
                case FE_TO_BYTES: fe_to_canonical<F>(r, x); break;
                case FE_SQRT:
                    fe_set_zero(r);
                    flag = (F::M0 == FqP::M0) ? fq_sqrt(r, x) : fr_sqrt(r, x);
                    if (!flag) fe_set_zero(r);
                    break;
                default:  // FE_FROM_BYTES
                    flag = fe_is_canonical<F>(x);
                    fe_from_raw<F>(r, x);
                    break;
            }
        }
        if (CANON && OP != FE_TO_BYTES) fe_to_canonical<F>(r, r);
        st_fe(out + i * 32, r);
        if ((OP == FE_INV || OP == FE_FROM_BYTES || OP == FE_SQRT) && ok) ok[i] = flag ? 1 : 0;
    }
}

// Batched inversion (the reference's Fr::invert / Fq::invert per element, src/fr.rs:438-540, CtOption::none
// for 0 -> ok[i] = 0 and a zero result): Montgomery's trick along each thread's strided chain, exactly as
// ff::BatchInverter does for batch_normalize (src/lib.rs:849, 1086), so a chain of ~32 elements costs ONE
// Fermat inversion plus 3 products per element.  The inverse of a field element is unique, so the limbs
// are the ones the reference's addition chain produces.  `scratch` (n x 32 B) holds the prefix products;
// `out` may alias `a`.
template <class F, bool CANON>
__global__ void __launch_bounds__(128) k_fe_invert_batched(const char* __restrict__ a, char* out, uint8_t* __restrict__ ok,
                                                           char* __restrict__ scratch, size_t n) {
    const size_t T = (size_t)gridDim.x * blockDim.x, t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    fe acc, x;
    fe_set_one<F>(acc);
    size_t cnt = 0;
    for (size_t i = t; i < n; i += T, cnt++) {
        ld_fe(x, a + i * 32);
        if (CANON) fe_from_raw<F>(x, x);
        st_fe(scratch + i * 32, acc);
        if (!fe_is_zero(x)) mont_mul<F>(acc, acc, x);
    }
    fe_invert<F>(acc, acc);
    for (size_t c = cnt; c-- > 0;) {
        const size_t i = t + c * T;
        fe pre, r;
        ld_fe(x, a + i * 32);
        if (CANON) fe_from_raw<F>(x, x);
        ld_fe(pre, scratch + i * 32);
        const bool nz = !fe_is_zero(x);
        fe_set_zero(r);
        if (nz) {
            mont_mul<F>(r, pre, acc);
            mont_mul<F>(acc, acc, x);
        }
        if (CANON) fe_to_canonical<F>(r, r);
        st_fe(out + i * 32, r);
        if (ok) ok[i] = nz ? 1 : 0;
    }
}

// ---- point batches ---------------------------------------------------------------------------
enum PtOp { PT_DBL = 0, PT_ADD, PT_ADD_NIELS, PT_ADD_AFFINE_NIELS, PT_TO_NIELS, PT_AFFINE_TO_NIELS };

template <int OP>
__global__ void __launch_bounds__(128) k_point_op(const char* __restrict__ p, const char* __restrict__ q,
                                                  char* __restrict__ out, size_t n, bool sub) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (OP == PT_AFFINE_TO_NIELS) {
            aff_point a;
            aff_niels nn;
            ld_fe(a.u, p + i * 64);
            ld_fe(a.v, p + i * 64 + 32);
            affine_to_niels_t<true>(nn, a);
            st_fe(out + i * 96, nn.vpu);
            st_fe(out + i * 96 + 32, nn.vmu);
            st_fe(out + i * 96 + 64, nn.t2d);
            continue;
        }
        ext_point P, R;
        ld_ext(P, p, i);
        if (OP == PT_TO_NIELS) {
            ext_niels nn;
            point_to_niels_t<true>(nn, P);
            st_fe(out + i * 128, nn.vpu);
            st_fe(out + i * 128 + 32, nn.vmu);
            st_fe(out + i * 128 + 64, nn.z);
            st_fe(out + i * 128 + 96, nn.t2d);
            continue;
        }
        if (OP == PT_DBL) {
            point_double_t<true>(R, P);
        } else if (OP == PT_ADD) {
            ext_point Q;
            ld_ext(Q, q, i);
            point_add_t<true>(R, P, Q, sub);
        } else if (OP == PT_ADD_NIELS) {
            ext_niels nn;
            ld_fe(nn.vpu, q + i * 128);
            ld_fe(nn.vmu, q + i * 128 + 32);
            ld_fe(nn.z, q + i * 128 + 64);
            ld_fe(nn.t2d, q + i * 128 + 96);
            point_add_niels_t<true>(R, P, nn, sub);
        } else {
            aff_niels nn;
            ld_fe(nn.vpu, q + i * 96);
            ld_fe(nn.vmu, q + i * 96 + 32);
            ld_fe(nn.t2d, q + i * 96 + 64);
            point_add_aff_niels_t<true>(R, P, nn, sub);
        }
        st_ext(out, i, R);
    }
}

// flags: 0 is_identity (src/lib.rs:691-696), 1 is_small_order (:699-705)
template <int WHAT>
__global__ void __launch_bounds__(128) k_point_flag(const char* __restrict__ p, uint8_t* __restrict__ out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        ext_point P;
        ld_ext(P, p, i);
        if (WHAT == 1) {  // src/lib.rs:699-705
            point_double(P, P);
            point_double(P, P);
            out[i] = fe_is_zero(P.u) ? 1 : 0;
        } else {
            out[i] = point_is_identity(P) ? 1 : 0;
        }
    }
}
// Neg for ExtendedPoint (src/lib.rs:195-210): (-U, V, Z, -T1, T2), all 160 bytes as the reference's.
__global__ void __launch_bounds__(256) k_point_neg(const char* __restrict__ p, char* __restrict__ out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        ext_point P;
        ld_ext(P, p, i);
        fe_neg<FqP>(P.u, P.u);
        fe_neg<FqP>(P.t1, P.t1);
        st_ext(out, i, P);
    }
}
// ConstantTimeEq for ExtendedPoint (src/lib.rs:153-163): u z' == u' z and v z' == v' z.
__global__ void __launch_bounds__(128) k_point_eq(const char* __restrict__ p, const char* __restrict__ q, uint8_t* __restrict__ out,
                                                  size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        fe pu, pv, pz, qu, qv, qz, a, b;
        ld_fe(pu, p + i * 160);
        ld_fe(pv, p + i * 160 + 32);
        ld_fe(pz, p + i * 160 + 64);
        ld_fe(qu, q + i * 160);
        ld_fe(qv, q + i * 160 + 32);
        ld_fe(qz, q + i * 160 + 64);
        fq_mul(a, pu, qz);
        fq_mul(b, qu, pz);
        bool eq = fe_eq(a, b);
        fq_mul(a, pv, qz);
        fq_mul(b, qv, pz);
        out[i] = (eq && fe_eq(a, b)) ? 1 : 0;
    }
}
// From<AffinePoint> for ExtendedPoint (src/lib.rs:214-226): (u, v) -> (u, v, 1, u, v).
__global__ void __launch_bounds__(256) k_affine_to_extended(const char* __restrict__ p, char* __restrict__ out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        ext_point P;
        ld_fe(P.u, p + i * 64);
        ld_fe(P.v, p + i * 64 + 32);
        fe_set_one<FqP>(P.z);
        P.t1 = P.u;
        P.t2 = P.v;
        st_ext(out, i, P);
    }
}
// mul_by_cofactor (src/lib.rs:722-724): three doublings with the reference's formula sequence => all 160 bytes equal.
__global__ void __launch_bounds__(128) k_mul_by_cofactor(const char* __restrict__ p, char* __restrict__ out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        ext_point P;
        ld_ext(P, p, i);
        point_double_t<true>(P, P);
        point_double_t<true>(P, P);
        point_double_t<true>(P, P);
        st_ext(out, i, P);
    }
}
// Sum<ExtendedPoint> (src/lib.rs:183-193: `iter.fold(identity, |acc, item| acc + item)`), batched and grouped: the input is
// `groups` consecutive groups of `g` points; one pass replaces every group by ceil(g / F) partial sums (thread t of a group
// adds the F consecutive points [tF, tF + F) of that group with the reference's `&ExtendedPoint + &ExtendedPoint`,
// :992-999), and the host repeats the pass until one point per group is left.  The association order differs from the
// reference's left fold, so the projective coordinates do; the point (normalised / encoded) is the same.
template <int F>
__global__ void __launch_bounds__(128) k_point_sum_pass(const char* __restrict__ in, char* __restrict__ out, size_t groups,
                                                        size_t g) {
    const size_t per = (g + F - 1) / F;  // partial sums per group after this pass
    const size_t total = groups * per, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += stride) {
        const size_t grp = w / per, t = w % per;
        const size_t lo = t * F, hi = lo + F < g ? lo + F : g;
        ext_point acc, q;
        ld_ext(acc, in, grp * g + lo);
#pragma unroll 1
        for (size_t i = lo + 1; i < hi; i++) {
            ld_ext(q, in, grp * g + i);
            point_add(acc, acc, q, false);
        }
        st_ext(out, w, acc);
    }
}

// Subgroup membership by the order-8 Tate pairing (torsion.cuh) -- is_torsion_free (src/lib.rs:709-711) and,
// with PRIME_ORDER, is_prime_order (:717-719: torsion free and not the identity).  STRIDE is the byte size of one
// input point: 160 (ExtendedPoint) or 64 (AffinePoint, z = 1).  With AND_INTO the result is combined with the flag
// already in out[i] (the `ok` of a preceding decode: SubgroupPoint::from_bytes, src/lib.rs:1427-1429).
template <int STRIDE, bool PRIME_ORDER, bool AND_INTO>
__global__ void __launch_bounds__(128, 4) k_is_torsion_free(const char* __restrict__ p, uint8_t* __restrict__ out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        fe U, V, Z;
        ld_fe(U, p + i * STRIDE);
        ld_fe(V, p + i * STRIDE + 32);
        if (STRIDE == 160) ld_fe(Z, p + i * STRIDE + 64);
        else fe_set_one<FqP>(Z);
        bool f = point_is_torsion_free(U, V, Z);
        if (PRIME_ORDER) f = f && !(fe_is_zero(U) && fe_eq(V, Z));
        if (AND_INTO) f = f && out[i] != 0;
        out[i] = f ? 1 : 0;
    }
}

// ---- variable-base scalar multiplication -------------------------------------------------------
#if defined(JJ_EXPERIMENTS)
// Window table in shared memory: one 32 KB slab per warp, laid out [entry][fe][half][lane] in
// 16-byte units, so a lane reading *its own* entry index is bank-conflict free (every quarter-
// warp touches eight distinct 16-byte bank groups regardless of the entry each lane picks).
// Measured slower than the L2-resident table (occupancy capped at 7 warps/SM): experiments only.
struct SmemTable {
    uint4* slab;  // warp slab + lane
    __device__ __forceinline__ void put(int idx, const fe& f) {
        slab[(idx * 2) * 32] = make_uint4(f.w[0], f.w[1], f.w[2], f.w[3]);
        slab[(idx * 2 + 1) * 32] = make_uint4(f.w[4], f.w[5], f.w[6], f.w[7]);
    }
    __device__ __forceinline__ void get(int idx, fe& f) const {
        uint4 lo = slab[(idx * 2) * 32], hi = slab[(idx * 2 + 1) * 32];
        f.w[0] = lo.x; f.w[1] = lo.y; f.w[2] = lo.z; f.w[3] = lo.w;
        f.w[4] = hi.x; f.w[5] = hi.y; f.w[6] = hi.z; f.w[7] = hi.w;
    }
    __device__ __forceinline__ void store(int j, const ext_niels& n) {
        put(j * 4, n.vpu); put(j * 4 + 1, n.vmu); put(j * 4 + 2, n.z); put(j * 4 + 3, n.t2d);
    }
    __device__ __forceinline__ void load(int j, ext_niels& n) const {
        get(j * 4, n.vpu); get(j * 4 + 1, n.vmu); get(j * 4 + 2, n.z); get(j * 4 + 3, n.t2d);
    }
    __device__ __forceinline__ void load_signed(int j, bool neg, ext_niels& n) const {
        const int o = neg ? 1 : 0;
        get(j * 4 + o, n.vpu); get(j * 4 + 1 - o, n.vmu); get(j * 4 + 2, n.z); get(j * 4 + 3, n.t2d);
    }
};
#endif
// Window table in global scratch (L2-resident: resident threads x 1 KB), laid out
// [warp][entry][fe][lane][8 words]: table build stores are fully coalesced 256-bit stores,
// lookups read four whole 32-byte sectors.
struct GmemTable {
    char* base;  // warp slab + lane * 32
    __device__ __forceinline__ void store(int j, const ext_niels& n) {
        char* p = base + (size_t)j * 4 * 1024;
        st_fe_keep(p, n.vpu); st_fe_keep(p + 1024, n.vmu); st_fe_keep(p + 2048, n.z); st_fe_keep(p + 3072, n.t2d);
    }
    __device__ __forceinline__ void load(int j, ext_niels& n) const {
        const char* p = base + (size_t)j * 4 * 1024;
        ld_fe_keep(n.vpu, p); ld_fe_keep(n.vmu, p + 1024); ld_fe_keep(n.z, p + 2048); ld_fe_keep(n.t2d, p + 3072);
    }
    // entry j with its v+u / v-u halves exchanged by address when neg (scalarmul.cuh: LocalTable::load_signed)
    __device__ __forceinline__ void load_signed(int j, bool neg, ext_niels& n) const {
        const char* p = base + (size_t)j * 4 * 1024;
        const int o = neg ? 1024 : 0;
        ld_fe_keep(n.vpu, p + o); ld_fe_keep(n.vmu, p + 1024 - o); ld_fe_keep(n.z, p + 2048); ld_fe_keep(n.t2d, p + 3072);
    }
};

enum { TABLE_SMEM = 0, TABLE_GMEM = 1 };

// Fused all-gather: when n_peers > 0 every result is stored straight into each peer GPU's copy of
// the gathered output (peer-mapped pointers, NVLink P2P stores) at unit offset `base_unit + i`,
// instead of into `out`.  The collective rides the compute kernel's epilogue; no separate gather.
struct PeerOut {
    char* ptr[8];
    int n_peers;
    size_t base_unit;
};
struct SmulArgs {
    const char* points;
    const char* scalars;
    size_t scalar_stride;
    char* out;
    uint8_t* flag_out;
    size_t n;
    char* tbl_scratch;
    bool scalar_mont;
    bool in_affine;      // points are AffinePoint (64 B): (u, v) -> (u, v, 1, u, v), src/lib.rs:214-226
    int out_unit;        // NORM kernels: 64 = AffinePoint, 32 = encoding (src/lib.rs:455-464)
    char* norm_scratch;  // NORM kernels: n x 128 B (u, v, z, prefix product) + one 32 B running product per thread
    PeerOut peers;
};
__device__ __forceinline__ void ld_point_any(ext_point& P, const SmulArgs& a, size_t i) {
    if (a.in_affine) {
        aff_point q;
        ld_fe_stream(q.u, a.points + i * 64);
        ld_fe_stream(q.v, a.points + i * 64 + 32);
        point_from_affine(P, q);
    } else {
        ld_ext_stream(P, a.points, i);
    }
}
__device__ __forceinline__ void smul_store(const SmulArgs& a, size_t i, const ext_point& acc) {
    if (a.peers.n_peers > 0) {
#pragma unroll 1
        for (int r = 0; r < a.peers.n_peers; r++) st_ext_stream(a.peers.ptr[r], a.peers.base_unit + i, acc);
    } else {
        st_ext_stream(a.out, i, acc);
    }
}
// normalised result -> AffinePoint (64 B) or its encoding (32 B), into `out` or into every peer's gathered buffer
__device__ __forceinline__ void smul_store_norm(const SmulArgs& a, size_t i, const fe& x, const fe& y) {
    fe enc;
    if (a.out_unit == 32) {  // AffinePoint::to_bytes: canonical v, bit 255 = parity of canonical u
        fe xc;
        fe_to_canonical<FqP>(xc, x);
        fe_to_canonical<FqP>(enc, y);
        enc.w[7] |= (xc.w[0] & 1u) << 31;
    }
    const int np = a.peers.n_peers > 0 ? a.peers.n_peers : 1;
#pragma unroll 1
    for (int r = 0; r < np; r++) {
        char* base = a.peers.n_peers > 0 ? a.peers.ptr[r] : a.out;
        const size_t u = a.peers.n_peers > 0 ? a.peers.base_unit + i : i;
        if (a.out_unit == 32) {
            st_fe_stream(base + u * 32, enc);
        } else {
            st_fe_stream(base + u * 64, x);
            st_fe_stream(base + u * 64 + 32, y);
        }
    }
}

// NORM = false: results leave as ExtendedPoint (160 B).  NORM = true: batch_normalize (src/lib.rs:1084-1107) and, for
// out_unit = 32, AffinePoint::to_bytes (:455-464) are folded into the kernel -- every thread keeps (u, v, z, prefix
// product of its earlier z) of its own units in scratch, inverts the product of all its z once (Montgomery's trick along
// the thread's own units, as ff::BatchInverter does along the slice; z = 0 is skipped and yields (0, 0)) and walks back.
// The fused all-gather then moves 32-byte encodings instead of 160-byte points.  Worth it when a thread owns several
// units (device-resident batches); host-staged chunks of one round use the separate k_batch_normalize pass.
// CT: the constant-time mode of scalar_mul_core (JJ_CONST_TIME): table scan, selects, every addition executed.
template <int THREADS, int MIN_BLOCKS, int TABLE, bool NORM, bool CT = false>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) k_scalar_mul(const SmulArgs a) {
#if defined(JJ_EXPERIMENTS)
    extern __shared__ uint4 smem_tbl[];
#endif
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t stride = (size_t)gridDim.x * THREADS;
    // warp-major slot order: warp w of every block comes before warp w+1 of any block, so the last,
    // partial round of the batch leaves every SM with the same number of busy warps instead of
    // leaving whole SMs idle (the kernel is pipe-bound: fewer warps per SM finish proportionally sooner)
    const size_t slot = ((size_t)warp * gridDim.x + blockIdx.x) * 32 + lane;
    char* run_slot = NORM ? a.norm_scratch + a.n * 128 + slot * 32 : nullptr;
    if (NORM && slot < a.n) {
        fe one;
        fe_set_one<FqP>(one);
        st_fe(run_slot, one);
    }
    for (size_t i = slot; i < a.n; i += stride) {
        ext_point P, acc;
        fe k;
        ld_point_any(P, a, i);
        ld_fe_stream(k, a.scalars + i * a.scalar_stride);
        if (a.scalar_mont) fe_to_canonical<FrP>(k, k);  // Fr::to_bytes, src/lib.rs:877
#if defined(JJ_EXPERIMENTS)
        if (TABLE == TABLE_SMEM) {
            SmemTable t{smem_tbl + (size_t)warp * 2048 + lane};
            scalar_mul_core(acc, P, k.w, t);
        } else
#endif
        {
            size_t gwarp = (size_t)blockIdx.x * (THREADS / 32) + warp;
            GmemTable t{a.tbl_scratch + gwarp * 32768 + lane * 32};
            scalar_mul_core<GmemTable, CT>(acc, P, k.w, t);
        }
        if (NORM) {
            fe run;
            char* rec = a.norm_scratch + i * 128;
            ld_fe(run, run_slot);
            st_fe_stream(rec, acc.u);
            st_fe_stream(rec + 32, acc.v);
            st_fe_stream(rec + 64, acc.z);
            st_fe_stream(rec + 96, run);
            if (!fe_is_zero(acc.z)) {
                fq_mul(run, run, acc.z);
                st_fe(run_slot, run);
            }
        } else if (a.flag_out) {
            a.flag_out[i] = point_is_identity(acc) ? 1 : 0;
        } else {
            smul_store(a, i, acc);
        }
    }
    if (NORM && slot < a.n) {
        fe inv;
        ld_fe(inv, run_slot);
        fq_pow_const_shared<ExpInvert<FqP>>(inv, inv);
        const size_t cnt = (a.n - slot + stride - 1) / stride;
#pragma unroll 1
        for (size_t c = cnt; c-- > 0;) {
            const size_t i = slot + c * stride;
            const char* rec = a.norm_scratch + i * 128;
            fe u, v, z, pre, zi;
            ld_fe_stream(u, rec);
            ld_fe_stream(v, rec + 32);
            ld_fe_stream(z, rec + 64);
            ld_fe_stream(pre, rec + 96);
            if (fe_is_zero(z)) {
                fe_set_zero(zi);
            } else {
                fq_mul(zi, pre, inv);
                fq_mul(inv, inv, z);
            }
            fq_mul(u, u, zi);
            fq_mul(v, v, zi);
            smul_store_norm(a, i, u, v);
        }
    }
}

// One scalar for the whole batch, given as its width-5 NAF (scalarmul.cuh).  This is the reference's own
// is_torsion_free, [r]P == O (src/lib.rs:709-711), kept as the on-device cross-check of the pairing test
// (jj_is_torsion_free with JJ_TORSION_LADDER).  Same mapping and table scratch as k_scalar_mul; digits come
// from the kernel parameters (uniform loads), so the add / no-add branch is warp-uniform.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_scalar_mul_const(const SmulArgs a, const NafDigits naf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t stride = (size_t)gridDim.x * THREADS;
    const size_t slot = ((size_t)warp * gridDim.x + blockIdx.x) * 32 + lane;
    for (size_t i = slot; i < a.n; i += stride) {
        ext_point P, acc;
        ld_ext_stream(P, a.points, i);
        size_t gwarp = (size_t)blockIdx.x * (THREADS / 32) + warp;
        GmemTable t{a.tbl_scratch + gwarp * 32768 + lane * 32};
        scalar_mul_wnaf_core(acc, P, naf, t);
        if (a.flag_out) a.flag_out[i] = point_is_identity(acc) ? 1 : 0;
        else smul_store(a, i, acc);
    }
}

#if defined(JJ_EXPERIMENTS)
// Slot-file mapping (slotmul.cuh): the working set of each thread lives in 13 shared-memory slots
// (13 KB per warp), field operations are shared noinline routines.  Window table in global scratch.
template <int THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) k_scalar_mul_slots(const SmulArgs a) {
    extern __shared__ uint4 smem_slots[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SlotFile S{(uint32_t)__cvta_generic_to_shared(smem_slots) + (uint32_t)warp * (S_COUNT * 1024u) + (uint32_t)lane * 16u};
    const size_t stride = (size_t)gridDim.x * THREADS;
    const size_t gwarp = (size_t)blockIdx.x * (THREADS / 32) + warp;
    GmemTable t{a.tbl_scratch + gwarp * 32768 + lane * 32};
    for (size_t i = (size_t)blockIdx.x * THREADS + threadIdx.x; i < a.n; i += stride) {
        fe k, f;
        const char* b = a.points + i * 160;
        ld_fe(f, b);       S.st(S_U, f);
        ld_fe(f, b + 32);  S.st(S_V, f);
        ld_fe(f, b + 64);  S.st(S_Z, f);
        ld_fe(f, b + 96);  S.st(S_T1, f);
        ld_fe(f, b + 128); S.st(S_T2, f);
        ld_fe(k, a.scalars + i * a.scalar_stride);
        if (a.scalar_mont) fe_to_canonical<FrP>(k, k);
        scalar_mul_slots(S, k.w, t);
        ext_point acc;
        S.ld(acc.u, S_U); S.ld(acc.v, S_V); S.ld(acc.z, S_Z); S.ld(acc.t1, S_T1); S.ld(acc.t2, S_T2);
        if (a.flag_out) a.flag_out[i] = point_is_identity(acc) ? 1 : 0;
        else smul_store(a, i, acc);
    }
}
#endif

// ---- fixed-base scalar multiplication -----------------------------------------------------------
// Builds entry e of the fixed-base table (scalarmul.cuh): thread e runs the variable-base core on
// the small scalar (j+1) << (W*i) -- the top-carry entry on 1 << (W*(NW-1)) followed by W doublings --
// and normalises with one Fermat inversion.  Once per base.
template <int W>
__global__ void __launch_bounds__(64) k_fixed_table_build(const char* __restrict__ base_affine,
                                                          uint32_t* __restrict__ table, char* __restrict__ scratch) {
    using G = FixedGeom<W>;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    aff_point B;
    ld_fe(B.u, base_affine);
    ld_fe(B.v, base_affine + 32);
    // grid-stride over the entries: the window-table scratch (1 KB per thread) is sized by the grid, not by the table
    for (int e = tid; e < G::ENTRIES; e += nthreads) {
        const bool top = e == G::NW * G::PER;
        int i = top ? G::NW - 1 : e / G::PER, j = top ? 0 : e % G::PER;
        ext_point P, acc;
        point_from_affine(P, B);
        fe k;
        fe_set_zero(k);
        const int extra = (!top && i == G::NW - 1) ? G::EXCESS : 0;  // doublings owed by the top window (FixedGeom)
        {   // k = (j + 1) << (W * i - extra); spans at most two words
            const int bit = W * i - extra;
            uint64_t v = (uint64_t)(j + 1) << (bit & 31);
#pragma unroll
            for (int w = 0; w < 8; w++) {
                if (w == (bit >> 5)) k.w[w] = (uint32_t)v;
                if (w == (bit >> 5) + 1) k.w[w] = (uint32_t)(v >> 32);
            }
        }
        size_t gwarp = (size_t)tid >> 5;
        GmemTable t{scratch + gwarp * 32768 + (tid & 31) * 32};
        scalar_mul_core(acc, P, k.w, t);
        const int dbl = top ? W : extra;
#pragma unroll 1
        for (int d = 0; d < dbl; d++) point_double(acc, acc);
        fe zi;
        fe_invert<FqP>(zi, acc.z);
        aff_point a;
        aff_niels nn;
        mont_mul<FqP>(a.u, acc.u, zi);
        mont_mul<FqP>(a.v, acc.v, zi);
        affine_to_niels(nn, a);
        uint32_t* dst = table + (size_t)e * 24;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            dst[w] = nn.vpu.w[w];
            dst[8 + w] = nn.vmu.w[w];
            dst[16 + w] = nn.t2d.w[w];
        }
    }
}

// The shared AffineNiels window table (216 KB for W = 7) is staged into shared memory by ONE bulk asynchronous
// copy (TMA engine, `cp.async.bulk.shared::cluster.global` -> SASS UBLKCP) that signals an
// mbarrier with its byte count; the threads of the block wait on the barrier's phase and read
// their first scalars while the copy is in flight.
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* mbar) {
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst), bar = (uint32_t)__cvta_generic_to_shared(mbar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(gmem_src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
    uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar), done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
template <int THREADS, int W, bool INL>
__global__ void __launch_bounds__(THREADS)
    k_scalar_mul_fixed(const uint32_t* __restrict__ table, const char* __restrict__ scalars, char* __restrict__ out,
                       size_t n, bool scalar_mont) {
    extern __shared__ __align__(128) uint4 smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    uint32_t* stab = (uint32_t*)smem_raw;
    if (threadIdx.x == 0) {
        uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) tma_bulk_g2s(stab, table, FixedGeom<W>::BYTES, &mbar);
    fixed_table_view view{stab};
    const size_t stride = (size_t)gridDim.x * THREADS;
    bool staged = false;
    for (size_t i = (size_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride) {
        fe k;
        ext_point acc;
        ld_fe(k, scalars + i * 32);  // overlaps the table copy on the first iteration
        if (scalar_mont) fe_to_canonical<FrP>(k, k);
        if (!staged) {
            mbar_wait(&mbar, 0);
            staged = true;
        }
        scalar_mul_fixed_core<W, INL>(acc, k.w, view);
        st_ext(out, i, acc);
    }
    if (!staged) mbar_wait(&mbar, 0);  // never leave the block while the bulk copy is in flight
}

// Fixed-base multiplication with WIDE windows: the table stays in global memory.  With W = 12 it is 21 windows x 2 048
// entries x 96 B = 4.1 MB -- L2-resident and, because the warps of an SM walk the windows nearly in step, mostly L1-resident
// (one window's sub-table is 196 KB) -- and a scalar-mul is 22 mixed additions instead of the 37 of the 7-bit table that
// fits shared memory.  Entries are read with three 256-bit non-coherent loads (LDG.E.256.CONSTANT).
struct fixed_table_gview {
    const char* base;  // [entry][96 bytes]
    __device__ __forceinline__ void ld(fe& r, const char* p) const {
        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
                       "=r"(r.w[7])
                     : "l"(p));
    }
    __device__ __forceinline__ void load_signed(int entry, bool neg, aff_niels& n) const {
        const char* p = base + (size_t)entry * 96;
        const int o = neg ? 32 : 0;
        ld(n.vpu, p + o);
        ld(n.vmu, p + 32 - o);
        ld(n.t2d, p + 64);
    }
};
template <int THREADS, int W>
__global__ void __launch_bounds__(THREADS)
    k_scalar_mul_fixed_gmem(const uint32_t* __restrict__ table, const char* __restrict__ scalars, char* __restrict__ out,
                            size_t n, bool scalar_mont) {
    fixed_table_gview view{(const char*)table};
    const size_t stride = (size_t)gridDim.x * THREADS;
    for (size_t i = (size_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride) {
        fe k;
        ext_point acc;
        ld_fe(k, scalars + i * 32);
        if (scalar_mont) fe_to_canonical<FrP>(k, k);
        scalar_mul_fixed_core<W, true>(acc, k.w, view);
        st_ext(out, i, acc);
    }
}

// ---- normalisation / encoding -------------------------------------------------------------------
// batch_normalize (src/lib.rs:840-858, 1084-1107): Montgomery's trick along each thread's strided
// chain (elements t, t+T, t+2T, ...), one Fermat inversion per thread.  `scratch` (n x 32 B; for FMT = 64 it is
// out[i].u itself, as the reference uses q.u / q.v) holds the running products.  z == 0 is skipped and yields (0, 0).
// FMT: 64 = AffinePoint out (the iterator form, :840-858); 32 = the encoding of that point (AffinePoint::to_bytes fused,
// :455-464); 160 = the reference's in-place form (:1084-1107): the ExtendedPoint itself becomes (u/z, v/z, 1, u/z, v/z)
// and `out` may be `in`.
template <int FMT>
__global__ void __launch_bounds__(128) k_batch_normalize(const char* in, char* out, char* scratch, size_t n) {
    const size_t T = (size_t)gridDim.x * blockDim.x, t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const size_t sstride = FMT == 64 ? 64 : 32;
    fe acc, z;
    fe_set_one<FqP>(acc);
    size_t cnt = 0;
    for (size_t i = t; i < n; i += T, cnt++) {
        ld_fe(z, in + i * 160 + 64);
        st_fe(scratch + i * sstride, acc);
        if (!fe_is_zero(z)) mont_mul<FqP>(acc, acc, z);
    }
    fe_invert<FqP>(acc, acc);
    for (size_t c = cnt; c-- > 0;) {
        size_t i = t + c * T;
        fe u, v, zi, s;
        ld_fe(z, in + i * 160 + 64);
        ld_fe(u, in + i * 160);
        ld_fe(v, in + i * 160 + 32);
        ld_fe(s, scratch + i * sstride);
        if (fe_is_zero(z)) {
            fe_set_zero(zi);
        } else {
            mont_mul<FqP>(zi, s, acc);
            mont_mul<FqP>(acc, acc, z);
        }
        mont_mul<FqP>(u, u, zi);
        mont_mul<FqP>(v, v, zi);
        if (FMT == 64) {
            st_fe(out + i * 64, u);
            st_fe(out + i * 64 + 32, v);
        } else if (FMT == 32) {
            fe uc;
            fe_to_canonical<FqP>(uc, u);
            fe_to_canonical<FqP>(v, v);
            v.w[7] |= (uc.w[0] & 1u) << 31;
            st_fe(out + i * 32, v);
        } else {
            // the reference's loop sets z = one for every point; a skipped z = 0 therefore ends as (0, 0, 1, 0, 0)
            fe one;
            fe_set_one<FqP>(one);
            st_fe(out + i * 160, u);
            st_fe(out + i * 160 + 32, v);
            st_fe(out + i * 160 + 64, one);
            st_fe(out + i * 160 + 96, u);
            st_fe(out + i * 160 + 128, v);
        }
    }
}
// AffinePoint::to_bytes (src/lib.rs:455-464): canonical v, bit 255 = lsb of canonical u.
__global__ void __launch_bounds__(256) k_affine_to_bytes(const char* __restrict__ in, char* __restrict__ out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        fe u, v;
        ld_fe(u, in + i * 64);
        ld_fe(v, in + i * 64 + 32);
        fe_to_canonical<FqP>(u, u);
        fe_to_canonical<FqP>(v, v);
        v.w[7] |= (u.w[0] & 1u) << 31;
        st_fe(out + i * 32, v);
    }
}

// AffinePoint::batch_from_bytes (src/lib.rs:541-627).  The reference batches the inversion of the denominators
// 1 + d v^2 (ff::BatchInvert, :596-600) because an inversion costs as much as the square root that follows; here the
// root of the quotient comes out of one power of num * den (fq_sqrt_ratio, fe.cuh), so there is nothing left to batch:
// one thread per encoding, no scratch, no second pass.  ok[i] = 0 and (0, 0) for rejected encodings.
__global__ void __launch_bounds__(128, 4) k_from_bytes(const char* __restrict__ in, char* __restrict__ out,
                                                       uint8_t* __restrict__ ok, size_t n, bool zip216) {
    const size_t T = (size_t)gridDim.x * blockDim.x;
#pragma unroll 1
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += T) {
        fe enc;
        aff_point p;
        ld_fe(enc, in + i * 32);
        const bool good = point_from_bytes(p, enc, zip216);
        st_fe(out + i * 64, p.u);
        st_fe(out + i * 64 + 32, p.v);
        if (ok) ok[i] = good ? 1 : 0;
    }
}

// Builds the Fq square-root tables (fe.cuh) once per device: one thread, ~1 900 products.  status[0] = 1 when the
// subgroup hash came out perfect.
__global__ void k_fq_sqrt_init(uint32_t* status) {
    if (blockIdx.x == 0 && threadIdx.x == 0) status[0] = fq_sqrt_tables_build(g_fq_sqrt_tab) ? 1u : 0u;
}

// Integer-multiplier peak probes: register-only IMAD.WIDE.U32 (32x32+64 -> 64), the instruction the Montgomery kernels are
// made of and the unit the scalar-mul roofline is counted in.  The multiplicand is the chain's own
// previous low word, so ptxas cannot hoist or strength-reduce the products.
//   MODE 0: 8 independent accumulate chains per thread, all-register operands
//   MODE 1: the same with an immediate multiplier (the form of the reduction rows): one register read fewer per instruction
//   MODE 2: dependent chains of whole Fq products (119 IMAD.WIDE + ~50 other instructions each) -- the densest multiplier
//           stream the engine's own code can issue; measured the highest of the three (8.74e12 /s against 8.53e12 / 8.67e12)
// jj_measure_imad_peak reports the best of the three: the roofline denominator is what the chip was SEEN to sustain.
template <int MODE>
__global__ void __launch_bounds__(256) k_imad_peak(uint32_t* sink, uint32_t seed, int iters) {
    if (MODE == 2) {
        fe a, b;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            a.w[i] = (seed + threadIdx.x * 977u + i * 31u) & 0x3fffffffu;
            b.w[i] = (seed * 3u + blockIdx.x * 131u + i * 17u) & 0x3fffffffu;
        }
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
            mont_mul<FqP, true>(a, a, b);
            mont_mul<FqP, true>(b, b, a);
        }
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) x ^= a.w[i] ^ b.w[i];
        if (x == 0x1234567u) sink[0] = x;
        return;
    }
    uint32_t lo[8], hi[8];
    const uint32_t a = seed * 2654435761u + threadIdx.x * 40503u + blockIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        lo[k] = seed + 977u * k + threadIdx.x;
        hi[k] = seed ^ (k << 8);
    }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                uint32_t m = lo[k];
                if (MODE == 0) {
                    mad_lo_cc(lo[k], a, m, lo[k]);
                    madc_hi(hi[k], a, m, hi[k]);
                } else {
                    JJ_MAD_LO_CC_I(lo[k], m, 0x53bda402, lo[k]);
                    JJ_MADC_HI_I(hi[k], m, 0x53bda402, hi[k]);
                }
            }
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) x ^= lo[k] ^ hi[k];
    if (x == 0x1234567u) sink[0] = x;  // keeps the chains alive
}
// IMAD.WIDE instructions per product of MODE 2 (checked against the SASS by scripts/sass_summary.py: the shared Fq
// product body is 119 IMAD.WIDE)
constexpr int kImadWidePerFqMul = 119;

// Writes a buffer larger than L2 (bench.py's flush between timed iterations).
__global__ void k_fill(uint4* p, size_t n16, uint32_t v) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) p[i] = make_uint4(v, v, v, v);
}

}  // namespace jj
