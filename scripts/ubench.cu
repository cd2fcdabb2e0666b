// Integer-pipe microbenchmarks for the roofline denominators (run on the B200 box).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench scripts/ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../jubjub_b200/csrc/point.cuh"
using namespace jj;

#define ITERS 4096
// 1. independent IMAD.WIDE.U32 (no carries)
__global__ void __launch_bounds__(256) k_wide(uint64_t* sink, uint32_t s) {
    uint64_t acc[16]; uint32_t a = s + threadIdx.x, b = s * 2654435761u + blockIdx.x;
    for (int k = 0; k < 16; k++) acc[k] = k * s;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int k = 0; k < 16; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a), "r"(b));
    uint64_t x = 0; for (int k = 0; k < 16; k++) x ^= acc[k];
    if (x == 0x1234567) sink[0] = x;
}
// 2. carry-chained IMAD.WIDE.U32.X: 4 chains of 4 per iteration (mad.lo.cc / madc.hi.cc pairs)
__global__ void __launch_bounds__(256) k_chain(uint32_t* sink, uint32_t s) {
    uint32_t r[32]; uint32_t a = s + threadIdx.x, b = s * 2654435761u + blockIdx.x;
    for (int k = 0; k < 32; k++) r[k] = k * s;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t* e = r + 8 * c;
            mad_lo_cc(e[0], a, b, e[0]); madc_hi_cc(e[1], a, b, e[1]);
            madc_lo_cc(e[2], a, b, e[2]); madc_hi_cc(e[3], a, b, e[3]);
            madc_lo_cc(e[4], a, b, e[4]); madc_hi_cc(e[5], a, b, e[5]);
            madc_lo_cc(e[6], a, b, e[6]); madc_hi(e[7], a, b, e[7]);
        }
    }
    uint32_t x = 0; for (int k = 0; k < 32; k++) x ^= r[k];
    if (x == 0x1234567) sink[0] = x;
}
// 3. IADD3 carry chains only (8-word add chains)
__global__ void __launch_bounds__(256) k_add(uint32_t* sink, uint32_t s) {
    uint32_t r[32]; uint32_t a = s + threadIdx.x;
    for (int k = 0; k < 32; k++) r[k] = k * s;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t* e = r + 8 * c;
            add_cc(e[0], e[0], a);
            for (int k = 1; k < 7; k++) addc_cc(e[k], e[k], a);
            addc(e[7], e[7], a);
        }
    }
    uint32_t x = 0; for (int k = 0; k < 32; k++) x ^= r[k];
    if (x == 0x1234567) sink[0] = x;
}
// 4. 1:1 mix of independent IMAD.WIDE and LOP3 (dual-pipe issue)
__global__ void __launch_bounds__(256) k_mix(uint64_t* sink, uint32_t s) {
    uint64_t acc[8]; uint32_t l[8]; uint32_t a = s + threadIdx.x, b = s * 2654435761u + blockIdx.x;
    for (int k = 0; k < 8; k++) { acc[k] = k * s; l[k] = k ^ s; }
#pragma unroll 1
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int rep = 0; rep < 2; rep++)
#pragma unroll
        for (int k = 0; k < 8; k++) {
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a), "r"(b));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[k]) : "r"(a), "r"(b));
        }
    uint64_t x = 0; for (int k = 0; k < 8; k++) x ^= acc[k] ^ l[k];
    if (x == 0x1234567) sink[0] = x;
}
// 5. dependent mont_mul / mont_sqr / fe_add chains in registers
template <int WHAT>
__global__ void __launch_bounds__(256) k_field(uint32_t* sink, uint32_t s) {
    fe x, y;
    for (int k = 0; k < 8; k++) { x.w[k] = (s + k * 77 + threadIdx.x) & 0x0fffffff; y.w[k] = (s * 3 + k + blockIdx.x) & 0x0fffffff; }
#pragma unroll 1
    for (int it = 0; it < ITERS / 4; it++) {
        if (WHAT == 0) { mont_mul<FqP>(x, x, y); mont_mul<FqP>(y, y, x); }
        else if (WHAT == 1) { fe_add<FqP>(x, x, y); fe_sub<FqP>(y, y, x); }
        else { mont_mul<FrP>(x, x, y); mont_mul<FrP>(y, y, x); }
    }
    uint32_t z = 0; for (int k = 0; k < 8; k++) z ^= x.w[k] ^ y.w[k];
    if (z == 0x1234567) sink[0] = z;
}
// 6. point double / add throughput in registers
template <int WHAT>
__global__ void __launch_bounds__(128, 3) k_point(uint32_t* sink, uint32_t s) {
    ext_point p; ext_niels n;
    fe* f = (fe*)&p;
    for (int j = 0; j < 5; j++) for (int k = 0; k < 8; k++) f[j].w[k] = (s + j * 1000 + k * 77 + threadIdx.x) & 0x0fffffff;
    point_to_niels(n, p);
#pragma unroll 1
    for (int it = 0; it < ITERS / 16; it++) {
        if (WHAT == 0) point_double(p, p); else point_add_niels(p, p, n, (it & 1));
    }
    uint32_t z = 0; for (int j = 0; j < 5; j++) for (int k = 0; k < 8; k++) z ^= f[j].w[k];
    if (z == 0x1234567) sink[0] = z;
}

int main() {
    void* sink; cudaMalloc(&sink, 256);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    printf("{\"sms\": %d", sms);
    for (int wps : {4, 8, 16, 32}) {  // warps per SM via blocks of 128 threads
        int blocks = sms * wps / 4, thr = 128;
        float ms; double thr_total = (double)blocks * thr;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto t = [&](auto launch) { float best = 1e30f; for (int rep = 0; rep < 3; rep++) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float m; cudaEventElapsedTime(&m, e0, e1); if (rep) best = m < best ? m : best; } return best; };
        ms = t([&] { k_wide<<<blocks, thr>>>((uint64_t*)sink, 7); });
        printf(",\n \"wide_w%d\": %.4e", wps, thr_total * ITERS * 16 / (ms * 1e-3));
        ms = t([&] { k_chain<<<blocks, thr>>>((uint32_t*)sink, 7); });
        printf(", \"chainX_w%d\": %.4e", wps, thr_total * ITERS * 16 / (ms * 1e-3));
        ms = t([&] { k_add<<<blocks, thr>>>((uint32_t*)sink, 7); });
        printf(", \"iadd3_w%d\": %.4e", wps, thr_total * ITERS * 32 / (ms * 1e-3));
        ms = t([&] { k_mix<<<blocks, thr>>>((uint64_t*)sink, 7); });
        printf(", \"mix_imad_w%d\": %.4e", wps, thr_total * ITERS * 16 / (ms * 1e-3));
        ms = t([&] { k_field<0><<<blocks, thr>>>((uint32_t*)sink, 7); });
        printf(", \"fq_mul_w%d\": %.4e", wps, thr_total * (ITERS / 4) * 2 / (ms * 1e-3));
        ms = t([&] { k_field<2><<<blocks, thr>>>((uint32_t*)sink, 7); });
        printf(", \"fr_mul_w%d\": %.4e", wps, thr_total * (ITERS / 4) * 2 / (ms * 1e-3));
        ms = t([&] { k_field<1><<<blocks, thr>>>((uint32_t*)sink, 7); });
        printf(", \"fq_addsub_w%d\": %.4e", wps, thr_total * (ITERS / 4) * 2 / (ms * 1e-3));
        if (wps <= 12 || wps == 16) {
            int pb = sms * wps / 4;
            if (wps <= 12) {
                ms = t([&] { k_point<0><<<pb, 128>>>((uint32_t*)sink, 7); });
                printf(", \"pt_dbl_w%d\": %.4e", wps, (double)pb * 128 * (ITERS / 16) / (ms * 1e-3));
                ms = t([&] { k_point<1><<<pb, 128>>>((uint32_t*)sink, 7); });
                printf(", \"pt_add_w%d\": %.4e", wps, (double)pb * 128 * (ITERS / 16) / (ms * 1e-3));
            }
        }
    }
    printf("}\n");
    return 0;
}
