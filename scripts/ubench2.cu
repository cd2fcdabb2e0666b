// Clean integer-pipe probes (operands are loop-carried so ptxas cannot hoist the products).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench2 scripts/ubench2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../jubjub_b200/csrc/ptx_ops.cuh"
using namespace jj;
#define ITERS 4096

__device__ __forceinline__ uint64_t madwide(uint32_t a, uint32_t b, uint64_t c) {
    uint64_t r; asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r;
}
// MODE 0: IMAD.WIDE.U32 only; 1: IMAD (32-bit lo); 2: IADD3 carry chains only;
// 3: IMAD.WIDE + IADD3 chains 1:1; 4: IMAD.WIDE + IADD3 1:2; 5: IMAD.HI
template <int MODE>
__global__ void __launch_bounds__(128) k(uint32_t* sink, uint32_t s) {
    uint64_t acc[8]; uint32_t x[8], r[16];
    uint32_t a = s + threadIdx.x;
    for (int k = 0; k < 8; k++) { acc[k] = (uint64_t)(k + 1) * s * 0x9E3779B97F4A7C15ull; x[k] = k * 77 + s; }
    for (int k = 0; k < 16; k++) r[k] = k * s + 1;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int rep = 0; rep < 2; rep++) {
            if (MODE == 0 || MODE == 3 || MODE == 4) {
#pragma unroll
                for (int k = 0; k < 8; k++) acc[k] = madwide(a, (uint32_t)acc[k], acc[k]);
            }
            if (MODE == 1) {
#pragma unroll
                for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(x[k]) : "r"(a));
            }
            if (MODE == 5) {
#pragma unroll
                for (int k = 0; k < 8; k++) asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(x[k]) : "r"(a));
            }
            if (MODE == 2 || MODE == 3 || MODE == 4) {
                add_cc(r[0], r[0], a);
#pragma unroll
                for (int k = 1; k < 7; k++) addc_cc(r[k], r[k], a);
                addc(r[7], r[7], a);
            }
            if (MODE == 4 || MODE == 2) {
                add_cc(r[8], r[8], a);
#pragma unroll
                for (int k = 9; k < 15; k++) addc_cc(r[k], r[k], a);
                addc(r[15], r[15], a);
            }
        }
    }
    uint32_t z = 0;
    for (int k = 0; k < 8; k++) z ^= (uint32_t)acc[k] ^ (uint32_t)(acc[k] >> 32) ^ x[k];
    for (int k = 0; k < 16; k++) z ^= r[k];
    if (z == 0x1234567) sink[0] = z;
}
template <int MODE> void go(const char* name, int sms, uint32_t* sink, double imads, double adds) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wps : {8, 16, 32}) {
        int blocks = sms * wps / 4; float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0); k<MODE><<<blocks, 128>>>(sink, 7 + rep); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best = ms < best ? ms : best;
        }
        double thr = (double)blocks * 128 * ITERS * 2;
        printf(" \"%s_w%d\": {\"imad_per_s\": %.4e, \"add_per_s\": %.4e},\n", name, wps, thr * imads / (best * 1e-3), thr * adds / (best * 1e-3));
    }
}
int main() {
    uint32_t* sink; cudaMalloc(&sink, 256);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    printf("{\n");
    go<0>("imad_wide", sms, sink, 8, 0);
    go<1>("imad_lo", sms, sink, 8, 0);
    go<5>("imad_hi", sms, sink, 8, 0);
    go<2>("iadd3_chain", sms, sink, 0, 16);
    go<3>("wide_plus_add_1to1", sms, sink, 8, 8);
    go<4>("wide_plus_add_1to2", sms, sink, 8, 16);
    printf(" \"sms\": %d}\n", sms);
    return 0;
}
