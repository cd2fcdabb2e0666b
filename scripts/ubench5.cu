// Component probes for the scalar-mul kernel: how busy can the multiplier pipe get on each building block alone?
// Register-only dependent chains of (a) Fq products, (b) Fq squarings, (c) point doublings (the kernel's inlined form),
// (d) fe_add/fe_sub only, at 16 warps/SM like the kernel (512-thread blocks, <= 128 registers).  Prints IMAD.WIDE thread-ops/s
// per probe (static counts from SASS: product 119, squaring 91, doubling 721) next to the pure IMAD.WIDE probe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DJJ_DOUBLE_INLINE -o scripts/ubench5 scripts/ubench5.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../jubjub_b200/csrc/point.cuh"
using namespace jj;

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(uint32_t* sink, uint32_t s, int iters) {
    fe a, b;
    ext_point P;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a.w[i] = (s + threadIdx.x * 977u + i * 31u) & 0x3fffffffu;
        b.w[i] = (s * 3u + blockIdx.x * 131u + i * 17u) & 0x3fffffffu;
    }
    P.u = a; P.v = b; P.z = a; P.t1 = b; P.t2 = a;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) { mont_mul<FqP>(a, a, b); mont_mul<FqP>(b, b, a); }
        if (MODE == 1) { mont_sqr<FqP>(a, a); mont_sqr<FqP>(b, b); }
        if (MODE == 2) { point_double_t<true>(P, P); }
        if (MODE == 3) { fe_add<FqP>(a, a, b); fe_sub<FqP>(b, b, a); fe_add<FqP>(a, a, b); fe_sub<FqP>(b, b, a); }
        if (MODE == 4) {  // two independent product chains per thread (ILP 2)
            fe c = P.u, d = P.v;
            mont_mul<FqP>(a, a, b); mont_mul<FqP>(c, c, d);
            mont_mul<FqP>(b, b, a); mont_mul<FqP>(d, d, c);
            P.u = c; P.v = d;
        }
    }
    uint32_t z = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) z ^= a.w[i] ^ b.w[i] ^ P.u.w[i] ^ P.v.w[i] ^ P.z.w[i] ^ P.t1.w[i] ^ P.t2.w[i];
    if (z == 0x1234567u) sink[0] = z;
}
// pure IMAD.WIDE chains: MODE 0 all-register operands (jj_measure_imad_peak's form), 1 immediate multiplier (the form of the
// reduction rows), at W warps per SM
template <int MODE>
__global__ void __launch_bounds__(256) kw(uint32_t* sink, uint32_t seed, int iters) {
    uint32_t lo[8], hi[8];
    const uint32_t a = seed * 2654435761u + threadIdx.x * 40503u + blockIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) { lo[k] = seed + 977u * k + threadIdx.x; hi[k] = seed ^ (k << 8); }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                uint32_t m = lo[k];
                if (MODE == 0) { mad_lo_cc(lo[k], a, m, lo[k]); madc_hi(hi[k], a, m, hi[k]); }
                else { JJ_MAD_LO_CC_I(lo[k], m, 0x53bda402, lo[k]); JJ_MADC_HI_I(hi[k], m, 0x53bda402, hi[k]); }
            }
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) x ^= lo[k] ^ hi[k];
    if (x == 0x1234567u) sink[0] = x;
}
template <int MODE>
void gow(const char* name, int sms, uint32_t* sink, int blocks_per_sm) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        kw<MODE><<<sms * blocks_per_sm, 256>>>(sink, 7 + rep, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = ms < best ? ms : best;
    }
    printf("{\"probe\": \"%s, %d warps/SM\", \"ms\": %.3f, \"imad_wide_per_s\": %.4e}\n", name, blocks_per_sm * 8, best,
           (double)sms * blocks_per_sm * 256.0 * iters * 32.0 / (best * 1e-3));
}
template <int MODE>
void go(const char* name, int sms, uint32_t* sink, int iters, double wide_per_iter, double other_per_iter) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k<MODE><<<sms, 512>>>(sink, 7 + rep, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = ms < best ? ms : best;
    }
    double thr = (double)sms * 512 * iters;
    printf("{\"probe\": \"%s\", \"ms\": %.3f, \"imad_wide_per_s\": %.4e, \"other_per_s\": %.4e, \"iters_per_s\": %.4e}\n", name, best,
           thr * wide_per_iter / (best * 1e-3), thr * other_per_iter / (best * 1e-3), thr / (best * 1e-3));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    uint32_t* sink; cudaMalloc(&sink, 64);
    int sms = p.multiProcessorCount;
    for (int b : {1, 2, 4, 8}) gow<0>("pure IMAD.WIDE, register operands", sms, sink, b);
    for (int b : {1, 2, 4, 8}) gow<1>("pure IMAD.WIDE, immediate multiplier", sms, sink, b);
    go<0>("fq_mul chain x2 (119 IMAD.WIDE + ~50 other each)", sms, sink, 20000, 2 * 119, 2 * 50);
    go<4>("two independent fq_mul chains x2", sms, sink, 10000, 4 * 119, 4 * 50);
    go<1>("fq_sqr chain x2 (91 IMAD.WIDE + ~106 other each)", sms, sink, 20000, 2 * 91, 2 * 106);
    go<2>("point_double chain (721 IMAD.WIDE + ~742 other)", sms, sink, 5000, 721, 742);
    go<3>("fe_add/fe_sub only (84 ALU per iter)", sms, sink, 100000, 0, 84);
    return 0;
}
