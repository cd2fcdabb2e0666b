// FP64 pipe probes: DFMA rate alone, and DFMA issued together with IMAD.WIDE.U32 chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench3 scripts/ubench3.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../jubjub_b200/csrc/ptx_ops.cuh"
using namespace jj;
#define ITERS 4096
// MODE 0: 16 independent DFMA chains; 1: 8 IMAD.WIDE accumulate chains; 2: both interleaved (16 DFMA + 8 IMAD.WIDE per rep)
template <int MODE>
__global__ void __launch_bounds__(128) k(double* sink, double s, uint32_t u) {
    double acc[16];
    uint32_t lo[8], hi[8];
    const double a = 1.0000001 + s * 1e-9, b = 0.9999999 - s * 1e-9;
    const uint32_t m = u * 2654435761u + threadIdx.x;
    for (int k2 = 0; k2 < 16; k2++) acc[k2] = s + k2;
    for (int k2 = 0; k2 < 8; k2++) { lo[k2] = u + 977u * k2 + threadIdx.x; hi[k2] = u ^ (k2 << 8); }
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int rep = 0; rep < 2; rep++) {
            if (MODE == 0 || MODE == 2) {
#pragma unroll
                for (int k2 = 0; k2 < 16; k2++) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(acc[k2]) : "d"(a), "d"(b));
            }
            if (MODE == 1 || MODE == 2) {
#pragma unroll
                for (int k2 = 0; k2 < 8; k2++) { uint32_t mm = lo[k2]; mad_lo_cc(lo[k2], m, mm, lo[k2]); madc_hi(hi[k2], m, mm, hi[k2]); }
            }
        }
    }
    double x = 0; for (int k2 = 0; k2 < 16; k2++) x += acc[k2];
    uint32_t z = 0; for (int k2 = 0; k2 < 8; k2++) z ^= lo[k2] ^ hi[k2];
    if (x == 1234.5 || z == 0x1234567u) sink[0] = x + z;
}
template <int MODE> void go(const char* name, int sms, double* sink, double dfma, double imad) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wps : {8, 16, 32}) {
        int blocks = sms * wps / 4; float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0); k<MODE><<<blocks, 128>>>(sink, 1.0 + rep, 7u + rep); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best = ms < best ? ms : best;
        }
        double thr = (double)blocks * 128 * ITERS * 2;
        printf(" \"%s_w%d\": {\"dfma_per_s\": %.4e, \"imad_wide_per_s\": %.4e},\n", name, wps, thr * dfma / (best * 1e-3), thr * imad / (best * 1e-3));
    }
}
int main() {
    double* sink; cudaMalloc(&sink, 256);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    printf("{\n");
    go<0>("dfma", sms, sink, 16, 0);
    go<1>("imad_wide", sms, sink, 0, 8);
    go<2>("dfma_plus_imad_wide", sms, sink, 16, 8);
    printf(" \"sms\": %d}\n", sms);
    return 0;
}
