// jubjub_b200.cu -- C ABI (include/jubjub_b200.h) over the sm_100a kernels in kernels.cuh.
//
// Host side of the boundary: argument checking, host<->device staging in chunks on two
// streams (copy of chunk c+1 overlaps compute of chunk c), kernel launch geometry, the
// fixed-base table cache, and the NCCL all-gather of sharded results.  There is no CPU
// arithmetic path in this library: if no CUDA device is usable every call fails with
// JJ_ERR_NO_DEVICE / JJ_ERR_CUDA.
#include "../../include/jubjub_b200.h"

#include <dlfcn.h>

#include <algorithm>
#include <mutex>
#include <initializer_list>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <new>

#include "kernels.cuh"

using namespace jj;

struct Id128 {  // ncclUniqueId, passed by value
    char b[128];
};

namespace {

constexpr size_t kChunkUnits = 1u << 17;  // host-staging chunk (units)
constexpr int kStages = 2;

struct Staging {
    cudaStream_t stream = nullptr;
    char* buf[4] = {nullptr, nullptr, nullptr, nullptr};  // up to 3 inputs + 1 output (+ok)
    size_t cap[4] = {0, 0, 0, 0};
    char* tbl = nullptr;  // scalar-mul window-table scratch (gmem variant)
    size_t tbl_cap = 0;
    char* tmp2 = nullptr;  // affine scratch for JJ_OUT_BYTES
    size_t tmp2_cap = 0;
};

// Minimal NCCL surface, resolved at run time so the library loads without NCCL.
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

bool load_nccl() {
    if (g_nccl.handle) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        void* h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (!h) continue;
        g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
        g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(h, "ncclCommInitRank");
        g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
        g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
        g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
        g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
        if (g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllGather && g_nccl.CommDestroy) {
            g_nccl.handle = h;
            return true;
        }
        dlclose(h);
    }
    return false;
}
}  // namespace

struct jj_ctx {
    int device = 0;
    int sm_count = 0;
    int smul_variant = 0;
    cudaStream_t stream = nullptr;  // user-visible ordering stream (device-pointer calls, timer)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    Staging st[kStages];
    char* tbl = nullptr;  // window-table scratch for the main stream
    size_t tbl_cap = 0;
    char* tmp = nullptr;  // extended-point scratch for JJ_OUT_AFFINE / JJ_OUT_BYTES
    size_t tmp_cap = 0;
    char* tmp2 = nullptr;
    size_t tmp2_cap = 0;
    uint32_t* fixed_table = nullptr;  // 64*8*24 words
    char* fixed_base_dev = nullptr;   // 64 B
    char fixed_base_key[64];
    bool fixed_valid = false;
    int fixed_w = 0;
    char* flush = nullptr;
    size_t flush_bytes = 0;
    void* nccl_comm = nullptr;
    int nranks = 1, rank = 0;
    PeerOut peers{};  // peer-mapped gathered-output buffers (fused all-gather); n_peers = 0: off
    int* barrier_word = nullptr;
    uint64_t launches = 0;
    char err[512];
};

namespace {

int32_t fail(jj_ctx* c, int32_t code, const char* fmt, ...) {
    if (c) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(c->err, sizeof c->err, fmt, ap);
        va_end(ap);
    }
    return code;
}
#define CU(c, call)                                                                              \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return fail((c), e_ == cudaErrorMemoryAllocation ? JJ_ERR_OOM : JJ_ERR_CUDA, "%s: %s", #call, \
                        cudaGetErrorString(e_));                                                 \
    } while (0)

int32_t ensure(jj_ctx* c, char** buf, size_t* cap, size_t bytes) {
    if (*cap >= bytes && *buf) return JJ_OK;
    if (*buf) CU(c, cudaFree(*buf));
    *buf = nullptr;
    *cap = 0;
    CU(c, cudaMalloc((void**)buf, bytes));
    *cap = bytes;
    return JJ_OK;
}

int grid_for(const jj_ctx* c, size_t n, int threads, int blocks_per_sm) {
    size_t need = (n + threads - 1) / threads;
    size_t cap = (size_t)c->sm_count * blocks_per_sm;
    return (int)std::max<size_t>(1, std::min(need, cap));
}

// ---- scalar-mul variants (jj_set_scalar_mul_variant) -------------------------------------------
struct SmulVariant {
    int threads, min_blocks, table;
};
const SmulVariant kVariants[] = {
    {0, 0, 0},                 // 0: default -> kDefaultVariant
    {224, 1, TABLE_SMEM},      // 1: 7 warps/SM, table in 224 KB of shared memory
    {128, 2, TABLE_GMEM},      // 2: 8 warps/SM, table in L2-resident global scratch
    {128, 3, TABLE_GMEM},      // 3: 12 warps/SM
    {128, 4, TABLE_GMEM},      // 4: 16 warps/SM (<= 128 registers)
    {256, 1, TABLE_GMEM},      // 5: 8 warps/SM in one block
    {192, 1, TABLE_SMEM},      // 6: 6 warps/SM, shared memory
    {128, 1, TABLE_SMEM},      // 7: 4 warps/SM, shared memory (one block of 128 KB)
    {64, 3, TABLE_SMEM},       // 8: 3 blocks x 2 warps, shared memory
    {96, 4, TABLE_GMEM},       // 9: 12 warps/SM in 4 blocks
    {64, 6, TABLE_GMEM},       // 10: 12 warps/SM in 6 blocks
    {384, 1, TABLE_GMEM},      // 11: 12 warps/SM in one block
    {192, 2, TABLE_GMEM},      // 12: 12 warps/SM in two blocks
    {512, 1, TABLE_GMEM},      // 13: 16 warps/SM in one block (<= 128 registers)
    {320, 1, TABLE_GMEM},      // 14: 10 warps/SM
    {512, 1, 2},               // 15: slot-file mapping (slotmul.cuh), 16 warps/SM
    {384, 1, 2},               // 16: slot-file, 12 warps/SM
    {544, 1, 2},               // 17: slot-file, 17 warps/SM (226 KB of shared memory)
    {256, 2, 2},               // 18: slot-file, 2 x 8 warps/SM
    {448, 1, TABLE_GMEM},      // 19: 14 warps/SM (<= 146 registers)
    {480, 1, TABLE_GMEM},      // 20: 15 warps/SM (<= 136 registers)
    {256, 2, TABLE_GMEM},      // 21: 16 warps/SM in two blocks (<= 128 registers)
    {576, 1, TABLE_GMEM},      // 22: 18 warps/SM (<= 113 registers, spills)
    {640, 1, TABLE_GMEM},      // 23: 20 warps/SM (<= 96 registers: spill code only around the additions)
    {768, 1, TABLE_GMEM},      // 24: 24 warps/SM (<= 80 registers)
    {704, 1, TABLE_GMEM},      // 25: 22 warps/SM (<= 88 registers)
    {896, 1, TABLE_GMEM},      // 26: 28 warps/SM (<= 72 registers)
    {1024, 1, TABLE_GMEM},     // 27: 32 warps/SM (<= 64 registers)
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);
constexpr int kDefaultVariant = 13;

template <int T, int MB, int TAB>
int32_t launch_smul_t(jj_ctx* c, cudaStream_t s, SmulArgs a, char** tbl, size_t* tbl_cap) {
    auto kern = k_scalar_mul<T, MB, TAB>;
    size_t smem = TAB == TABLE_SMEM ? (size_t)(T / 32) * 32768 : 0;
    int grid = grid_for(c, a.n, T, MB);
    if (TAB == TABLE_SMEM) {
        CU(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    } else {
        int32_t rc = ensure(c, tbl, tbl_cap, (size_t)grid * (T / 32) * 32768);
        if (rc) return rc;
    }
    a.tbl_scratch = *tbl;
    kern<<<grid, T, smem, s>>>(a);
    c->launches++;
    CU(c, cudaGetLastError());
    return JJ_OK;
}
template <int T, int MB>
int32_t launch_smul_slots(jj_ctx* c, cudaStream_t s, SmulArgs a, char** tbl, size_t* tbl_cap) {
    auto kern = k_scalar_mul_slots<T, MB>;
    size_t smem = (size_t)(T / 32) * S_COUNT * 1024;
    int grid = grid_for(c, a.n, T, MB);
    CU(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int32_t rc = ensure(c, tbl, tbl_cap, (size_t)grid * (T / 32) * 32768);
    if (rc) return rc;
    a.tbl_scratch = *tbl;
    kern<<<grid, T, smem, s>>>(a);
    c->launches++;
    CU(c, cudaGetLastError());
    return JJ_OK;
}
int32_t launch_smul(jj_ctx* c, cudaStream_t s, const char* pts, const char* sc, size_t sc_stride, char* out,
                    uint8_t* flag_out, size_t n, char** tbl, size_t* tbl_cap, bool scalar_mont,
                    const PeerOut* peers = nullptr) {
    int v = c->smul_variant > 0 && c->smul_variant < kNumVariants ? c->smul_variant : kDefaultVariant;
    SmulArgs a{};
    a.points = pts;
    a.scalars = sc;
    a.scalar_stride = sc_stride;
    a.out = out;
    a.flag_out = flag_out;
    a.n = n;
    a.scalar_mont = scalar_mont;
    if (peers) a.peers = *peers;
#define V(ID, T, MB, TAB) \
    case ID: return launch_smul_t<T, MB, TAB>(c, s, a, tbl, tbl_cap)
    switch (v) {
        V(1, 224, 1, TABLE_SMEM);
        V(2, 128, 2, TABLE_GMEM);
        V(3, 128, 3, TABLE_GMEM);
        V(4, 128, 4, TABLE_GMEM);
        V(5, 256, 1, TABLE_GMEM);
        V(6, 192, 1, TABLE_SMEM);
        V(7, 128, 1, TABLE_SMEM);
        V(8, 64, 3, TABLE_SMEM);
        V(9, 96, 4, TABLE_GMEM);
        V(10, 64, 6, TABLE_GMEM);
        V(11, 384, 1, TABLE_GMEM);
        V(12, 192, 2, TABLE_GMEM);
        V(13, 512, 1, TABLE_GMEM);
        V(14, 320, 1, TABLE_GMEM);
        V(19, 448, 1, TABLE_GMEM);
        V(20, 480, 1, TABLE_GMEM);
        V(21, 256, 2, TABLE_GMEM);
        V(22, 576, 1, TABLE_GMEM);
        V(23, 640, 1, TABLE_GMEM);
        V(24, 768, 1, TABLE_GMEM);
        V(25, 704, 1, TABLE_GMEM);
        V(26, 896, 1, TABLE_GMEM);
        V(27, 1024, 1, TABLE_GMEM);
        case 15: return launch_smul_slots<512, 1>(c, s, a, tbl, tbl_cap);
        case 16: return launch_smul_slots<384, 1>(c, s, a, tbl, tbl_cap);
        case 17: return launch_smul_slots<544, 1>(c, s, a, tbl, tbl_cap);
        case 18: return launch_smul_slots<256, 2>(c, s, a, tbl, tbl_cap);
    }
#undef V
    return fail(c, JJ_ERR_INVALID_ARG, "bad scalar-mul variant %d", v);
}

// ---- generic batched dispatch ---------------------------------------------------------------------
struct In {
    const void* p;
    size_t unit;  // bytes per unit; 0 = absent
};
struct Out {
    void* p;
    size_t unit;
};
// Launch(stream, din[3], dout[2], count, staging_or_null) -> status
// chunk_units (<= kChunkUnits): units per staged chunk of a host-pointer call.
template <class Launch>
int32_t run_batch(jj_ctx* c, uint32_t flags, size_t n, const In (&ins)[3], const Out (&outs)[2], Launch launch,
                  size_t chunk_units = kChunkUnits) {
    if (!c) return JJ_ERR_INVALID_ARG;
    for (const In& i : ins)
        if (i.unit && !i.p && n) return fail(c, JJ_ERR_INVALID_ARG, "null input pointer");
    if (outs[0].unit && !outs[0].p && n) return fail(c, JJ_ERR_INVALID_ARG, "null output pointer");
    CU(c, cudaSetDevice(c->device));
    if (n == 0) return JJ_OK;
    if (flags & JJ_DEVICE_PTRS) {
        const char* din[3];
        char* dout[2];
        for (int k = 0; k < 3; k++) {
            din[k] = (const char*)ins[k].p;
            if (ins[k].unit && ((uintptr_t)din[k] & 31)) return fail(c, JJ_ERR_INVALID_ARG, "device pointer not 32-byte aligned");
        }
        for (int k = 0; k < 2; k++) dout[k] = (char*)outs[k].p;
        if (outs[0].unit >= 32 && ((uintptr_t)dout[0] & 31)) return fail(c, JJ_ERR_INVALID_ARG, "device pointer not 32-byte aligned");
        int32_t rc = launch(c->stream, din, dout, n, (Staging*)nullptr);
        if (rc) return rc;
        if (!(flags & JJ_ASYNC)) CU(c, cudaStreamSynchronize(c->stream));
        return JJ_OK;
    }
    // host pointers: chunked, double-buffered staging
    size_t done = 0;
    int stage = 0;
    while (done < n) {
        size_t cnt = std::min(chunk_units, n - done);
        Staging& S = c->st[stage];
        const char* din[3] = {nullptr, nullptr, nullptr};
        char* dout[2] = {nullptr, nullptr};
        // the stream's previous chunk must have left the staging buffers
        CU(c, cudaStreamSynchronize(S.stream));
        size_t out_off[2] = {0, 0};
        for (int k = 0; k < 3; k++) {
            if (!ins[k].unit) continue;
            int32_t rc = ensure(c, &S.buf[k], &S.cap[k], kChunkUnits * ins[k].unit);
            if (rc) return rc;
            CU(c, cudaMemcpyAsync(S.buf[k], (const char*)ins[k].p + done * ins[k].unit, cnt * ins[k].unit,
                                  cudaMemcpyHostToDevice, S.stream));
            din[k] = S.buf[k];
        }
        {
            size_t need = 0;
            for (int k = 0; k < 2; k++) {
                out_off[k] = need;
                need += (kChunkUnits * outs[k].unit + 255) & ~(size_t)255;
            }
            int32_t rc = ensure(c, &S.buf[3], &S.cap[3], need);
            if (rc) return rc;
            for (int k = 0; k < 2; k++)
                if (outs[k].unit && outs[k].p) dout[k] = S.buf[3] + out_off[k];
        }
        int32_t rc = launch(S.stream, din, dout, cnt, &S);
        if (rc) return rc;
        for (int k = 0; k < 2; k++)
            if (outs[k].unit && outs[k].p)
                CU(c, cudaMemcpyAsync((char*)outs[k].p + done * outs[k].unit, dout[k], cnt * outs[k].unit,
                                      cudaMemcpyDeviceToHost, S.stream));
        done += cnt;
        stage = (stage + 1) % kStages;
    }
    for (int k = 0; k < kStages; k++) CU(c, cudaStreamSynchronize(c->st[k].stream));
    return JJ_OK;
}

template <class F, int OP>
int32_t fe_launch(jj_ctx* c, cudaStream_t s, bool canon, const char* a, const char* b, char* out, uint8_t* ok, size_t n,
                  uint64_t seed, size_t first) {
    int grid = grid_for(c, n, 256, (OP == FE_INV || OP == FE_SQRT) ? 4 : 8);
    if (canon)
        k_fe_op<F, OP, true><<<grid, 256, 0, s>>>(a, b, out, ok, n, seed, first);
    else
        k_fe_op<F, OP, false><<<grid, 256, 0, s>>>(a, b, out, ok, n, seed, first);
    c->launches++;
    CU(c, cudaGetLastError());
    return JJ_OK;
}
template <class F, int OP>
int32_t fe_binary(jj_ctx* c, const void* a, const void* b, void* out, size_t n, uint32_t flags) {
    In ins[3] = {{a, 32}, {b, b ? (size_t)32 : 0}, {nullptr, 0}};
    Out outs[2] = {{out, 32}, {nullptr, 0}};
    bool canon = flags & JJ_CANON;
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging*) {
        return fe_launch<F, OP>(c, s, canon, din[0], din[1], dout[0], nullptr, cnt, 0, 0);
    });
}
// Batched inversion: chains of ~32 elements per thread share one Fermat inversion (k_fe_invert_batched).
template <class F>
int32_t fe_invert_batch(jj_ctx* c, const void* a, void* out, uint8_t* ok, size_t n, uint32_t flags) {
    In ins[3] = {{a, 32}, {nullptr, 0}, {nullptr, 0}};
    Out outs[2] = {{out, 32}, {ok, 1}};
    const bool canon = flags & JJ_CANON;
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging* S) -> int32_t {
        char** scr = S ? &S->tmp2 : &c->tmp2;
        size_t* cap = S ? &S->tmp2_cap : &c->tmp2_cap;
        int32_t rc = ensure(c, scr, cap, cnt * 32);
        if (rc) return rc;
        size_t blocks = (cnt + 128 * 32 - 1) / (128 * 32);
        size_t per_wave = (size_t)c->sm_count * 2;
        blocks = std::max<size_t>(1, (blocks + per_wave - 1) / per_wave * per_wave);
        blocks = std::min(blocks, (cnt + 127) / 128);
        if (canon)
            k_fe_invert_batched<F, true><<<(int)blocks, 128, 0, s>>>(din[0], dout[0], (uint8_t*)dout[1], *scr, cnt);
        else
            k_fe_invert_batched<F, false><<<(int)blocks, 128, 0, s>>>(din[0], dout[0], (uint8_t*)dout[1], *scr, cnt);
        c->launches++;
        CU(c, cudaGetLastError());
        return JJ_OK;
    });
}
template <class F, int OP>
int32_t fe_with_ok(jj_ctx* c, const void* a, void* out, uint8_t* ok, size_t n, uint32_t flags) {
    In ins[3] = {{a, 32}, {nullptr, 0}, {nullptr, 0}};
    Out outs[2] = {{out, 32}, {ok, 1}};
    bool canon = (flags & JJ_CANON) && OP != FE_FROM_BYTES;
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging*) {
        return fe_launch<F, OP>(c, s, canon, din[0], nullptr, dout[0], (uint8_t*)dout[1], cnt, 0, 0);
    });
}

template <int OP>
int32_t pt_launch(jj_ctx* c, cudaStream_t s, const char* p, const char* q, char* out, size_t n, bool sub) {
    int grid = grid_for(c, n, 128, 4);
    k_point_op<OP><<<grid, 128, 0, s>>>(p, q, out, n, sub);
    c->launches++;
    CU(c, cudaGetLastError());
    return JJ_OK;
}
template <int OP>
int32_t pt_binary(jj_ctx* c, const void* p, size_t pu, const void* q, size_t qu, void* out, size_t ou, size_t n,
                  uint32_t flags) {
    if (flags & JJ_CANON) return fail(c, JJ_ERR_INVALID_ARG, "point entry points take Montgomery-form coordinates");
    In ins[3] = {{p, pu}, {q, qu}, {nullptr, 0}};
    Out outs[2] = {{out, ou}, {nullptr, 0}};
    bool sub = flags & JJ_SUBTRACT;
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging*) {
        return pt_launch<OP>(c, s, din[0], din[1], dout[0], cnt, sub);
    });
}

int32_t normalize_launch(jj_ctx* c, cudaStream_t s, const char* in, char* out, size_t n) {
    // one Fermat inversion per thread amortised over its strided chain: ~32 points per thread for
    // large batches, grid sized in whole multiples of the SM count (2 x 128-thread blocks each)
    size_t blocks = (n + 128 * 32 - 1) / (128 * 32);
    size_t per_wave = (size_t)c->sm_count * 2;
    blocks = std::max<size_t>(1, (blocks + per_wave - 1) / per_wave * per_wave);
    blocks = std::min(blocks, (n + 127) / 128);
    k_batch_normalize<<<(int)blocks, 128, 0, s>>>(in, out, n);
    c->launches++;
    CU(c, cudaGetLastError());
    return JJ_OK;
}
int32_t from_bytes_launch(jj_ctx* c, cudaStream_t s, const char* in, char* out, uint8_t* ok, size_t n, bool zip216) {
    // chains of ~8 encodings per thread: the Fermat inversion is amortised 8x while 4 x 128-thread blocks per
    // SM stay resident; grid in whole multiples of the SM count
    size_t blocks = (n + 128 * 8 - 1) / (128 * 8);
    size_t per_wave = (size_t)c->sm_count * 4;
    blocks = std::max<size_t>(1, (blocks + per_wave - 1) / per_wave * per_wave);
    blocks = std::min(blocks, (n + 127) / 128);
    k_from_bytes<<<(int)blocks, 128, 0, s>>>(in, out, ok, n, zip216);
    c->launches++;
    CU(c, cudaGetLastError());
    return JJ_OK;
}
int32_t to_bytes_launch(jj_ctx* c, cudaStream_t s, const char* in, char* out, size_t n) {
    k_affine_to_bytes<<<grid_for(c, n, 256, 8), 256, 0, s>>>(in, out, n);
    c->launches++;
    CU(c, cudaGetLastError());
    return JJ_OK;
}
// extended (device) -> requested output format (device)
int32_t finish_output(jj_ctx* c, cudaStream_t s, const char* ext, char* out, size_t n, uint32_t flags, char** tmp2,
                      size_t* tmp2_cap) {
    if (flags & JJ_OUT_BYTES) {
        int32_t rc = ensure(c, tmp2, tmp2_cap, n * 64);
        if (rc) return rc;
        rc = normalize_launch(c, s, ext, *tmp2, n);
        if (rc) return rc;
        return to_bytes_launch(c, s, *tmp2, out, n);
    }
    return normalize_launch(c, s, ext, out, n);
}
size_t out_unit(uint32_t flags) { return (flags & JJ_OUT_BYTES) ? 32 : (flags & JJ_OUT_AFFINE) ? 64 : 160; }

// fixed-base window width: 7 (216 KB table, the default) or 4 (47 KB table, variant 100)
int fixed_w(const jj_ctx* c) { return c->smul_variant == 100 ? 4 : 7; }

int32_t build_fixed_table(jj_ctx* c, const void* base_affine, uint32_t flags) {
    char key[64];
    if (flags & JJ_DEVICE_PTRS)
        CU(c, cudaMemcpy(key, base_affine, 64, cudaMemcpyDeviceToHost));
    else
        memcpy(key, base_affine, 64);
    const int w = fixed_w(c);
    if (c->fixed_valid && c->fixed_w == w && memcmp(key, c->fixed_base_key, 64) == 0) return JJ_OK;
    if (!c->fixed_table) CU(c, cudaMalloc((void**)&c->fixed_table, FixedGeom<7>::BYTES));
    if (!c->fixed_base_dev) CU(c, cudaMalloc((void**)&c->fixed_base_dev, 64));
    CU(c, cudaMemcpy(c->fixed_base_dev, key, 64, cudaMemcpyHostToDevice));
    const int entries = w == 4 ? FixedGeom<4>::ENTRIES : FixedGeom<7>::ENTRIES;
    const int blocks = (entries + 63) / 64;
    int32_t rc = ensure(c, &c->tbl, &c->tbl_cap, (size_t)(blocks * 2) * 32768);
    if (rc) return rc;
    if (w == 4)
        k_fixed_table_build<4><<<blocks, 64, 0, c->stream>>>(c->fixed_base_dev, c->fixed_table, c->tbl);
    else
        k_fixed_table_build<7><<<blocks, 64, 0, c->stream>>>(c->fixed_base_dev, c->fixed_table, c->tbl);
    c->launches++;
    CU(c, cudaGetLastError());
    CU(c, cudaStreamSynchronize(c->stream));
    memcpy(c->fixed_base_key, key, 64);
    c->fixed_valid = true;
    c->fixed_w = w;
    return JJ_OK;
}

template <int T, int W, bool INL>
int32_t launch_fixed(jj_ctx* c, cudaStream_t s, const char* scalars, char* dst, size_t cnt, bool smont) {
    auto kern = k_scalar_mul_fixed<T, W, INL>;
    size_t smem = FixedGeom<W>::BYTES;
    // static mbarrier + dynamic table: opt in explicitly
    CU(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem + 1024));
    const int per_sm = smem > 100000 ? 1 : 2;
    kern<<<grid_for(c, cnt, T, per_sm), T, smem, s>>>(c->fixed_table, scalars, dst, cnt, smont);
    c->launches++;
    CU(c, cudaGetLastError());
    return JJ_OK;
}

}  // namespace

// =================================================================================== C ABI
extern "C" {

const char* jj_version(void) { return "jubjub_b200 0.1.0 (sm_100a)"; }

int32_t jj_init(int device, jj_ctx** out) {
    if (!out) return JJ_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return JJ_ERR_NO_DEVICE;
    if (device < 0 || device >= count) return JJ_ERR_INVALID_ARG;
    jj_ctx* c = new (std::nothrow) jj_ctx();
    if (!c) return JJ_ERR_OOM;
    c->err[0] = 0;
    c->device = device;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete c;
        return JJ_ERR_CUDA;
    }
    if (prop.major < 10) {
        delete c;
        return JJ_ERR_NO_DEVICE;  // sm_100a code only
    }
    c->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int k = 0; k < kStages && ok; k++) ok = cudaStreamCreateWithFlags(&c->st[k].stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreate(&c->ev0) == cudaSuccess && cudaEventCreate(&c->ev1) == cudaSuccess;
    if (!ok) {
        jj_destroy(c);
        return JJ_ERR_CUDA;
    }
    {   // device-resident tables of the table-driven Fq square root (decode path): built once per device -- a
        // second context must not rewrite them under a kernel of the first.  Whether they are there is read
        // from the device itself (the `ready` word of the table), so a cudaDeviceReset in between is noticed.
        static std::mutex mu;
        std::lock_guard<std::mutex> lock(mu);
        uint32_t ready = 0;
        ok = cudaMemcpyFromSymbol(&ready, g_fq_sqrt_tab, sizeof(ready), offsetof(FqSqrtTables, ready)) == cudaSuccess;
        if (ok && ready != 1u) {
            uint32_t* st = nullptr;
            uint32_t host = 0;
            ok = cudaMalloc(&st, sizeof(uint32_t)) == cudaSuccess;
            if (ok) {
                k_fq_sqrt_init<<<1, 32, 0, c->stream>>>(st);
                ok = cudaMemcpyAsync(&host, st, sizeof(host), cudaMemcpyDeviceToHost, c->stream) == cudaSuccess &&
                     cudaStreamSynchronize(c->stream) == cudaSuccess && host == 1u;
                cudaFree(st);
            }
        }
        if (!ok) {
            jj_destroy(c);
            return JJ_ERR_CUDA;
        }
    }
    *out = c;
    return JJ_OK;
}

int32_t jj_destroy(jj_ctx* c) {
    if (!c) return JJ_ERR_INVALID_ARG;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
    if (c->barrier_word) cudaFree(c->barrier_word);
    for (int k = 0; k < kStages; k++) {
        for (char* b : c->st[k].buf)
            if (b) cudaFree(b);
        if (c->st[k].tbl) cudaFree(c->st[k].tbl);
        if (c->st[k].tmp2) cudaFree(c->st[k].tmp2);
        if (c->st[k].stream) cudaStreamDestroy(c->st[k].stream);
    }
    for (void* p : {(void*)c->tbl, (void*)c->tmp, (void*)c->tmp2, (void*)c->fixed_table, (void*)c->fixed_base_dev,
                    (void*)c->flush})
        if (p) cudaFree(p);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return JJ_OK;
}

int32_t jj_sync(jj_ctx* c) {
    if (!c) return JJ_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    return JJ_OK;
}
const char* jj_last_error(const jj_ctx* c) { return c ? c->err : "null context"; }
uint64_t jj_launch_count(const jj_ctx* c) { return c ? c->launches : 0; }
int32_t jj_set_scalar_mul_variant(jj_ctx* c, int32_t v) {
    if (!c || v < 0 || (v >= kNumVariants && v != 100)) return JJ_ERR_INVALID_ARG;  // 100: fixed-base kernel with shared Fq bodies
    c->smul_variant = v;
    return JJ_OK;
}
int32_t jj_device_info(jj_ctx* c, int32_t* sm_count, int32_t* sm_clock_khz, uint64_t* hbm_bytes) {
    if (!c) return JJ_ERR_INVALID_ARG;
    cudaDeviceProp prop;
    CU(c, cudaGetDeviceProperties(&prop, c->device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (sm_clock_khz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
        *sm_clock_khz = khz;
    }
    if (hbm_bytes) *hbm_bytes = prop.totalGlobalMem;
    return JJ_OK;
}

int32_t jj_malloc(jj_ctx* c, size_t bytes, void** dptr) {
    if (!c || !dptr) return JJ_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaMalloc(dptr, bytes ? bytes : 1));
    return JJ_OK;
}
int32_t jj_free(jj_ctx* c, void* dptr) {
    if (!c) return JJ_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaFree(dptr));
    return JJ_OK;
}
int32_t jj_host_alloc(jj_ctx* c, size_t bytes, void** hptr) {
    if (!c || !hptr) return JJ_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return JJ_OK;
}
int32_t jj_host_free(jj_ctx* c, void* hptr) {
    if (!c) return JJ_ERR_INVALID_ARG;
    CU(c, cudaFreeHost(hptr));
    return JJ_OK;
}
int32_t jj_memcpy_h2d(jj_ctx* c, void* dptr, const void* hptr, size_t bytes) {
    if (!c || (bytes && (!dptr || !hptr))) return JJ_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return JJ_OK;
}
int32_t jj_memcpy_d2h(jj_ctx* c, void* hptr, const void* dptr, size_t bytes) {
    if (!c || (bytes && (!dptr || !hptr))) return JJ_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return JJ_OK;
}
int32_t jj_timer_start(jj_ctx* c) {
    if (!c) return JJ_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaEventRecord(c->ev0, c->stream));
    return JJ_OK;
}
int32_t jj_timer_stop(jj_ctx* c, float* ms) {
    if (!c || !ms) return JJ_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaEventRecord(c->ev1, c->stream));
    CU(c, cudaEventSynchronize(c->ev1));
    CU(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return JJ_OK;
}
int32_t jj_flush_l2(jj_ctx* c) {
    if (!c) return JJ_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    if (!c->flush) {
        c->flush_bytes = (size_t)256 << 20;  // > 126 MB L2
        CU(c, cudaMalloc((void**)&c->flush, c->flush_bytes));
    }
    k_fill<<<c->sm_count * 8, 256, 0, c->stream>>>((uint4*)c->flush, c->flush_bytes / 16, (uint32_t)c->launches);
    CU(c, cudaGetLastError());
    return JJ_OK;
}

// ---- CUDA graphs: capture a sequence of JJ_DEVICE_PTRS | JJ_ASYNC calls once, replay it with one launch ----
int32_t jj_graph_begin(jj_ctx* c) {
    if (!c) return JJ_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    return JJ_OK;
}
int32_t jj_graph_end(jj_ctx* c, void** graph_exec) {
    if (!c || !graph_exec) return JJ_ERR_INVALID_ARG;
    cudaGraph_t g = nullptr;
    CU(c, cudaStreamEndCapture(c->stream, &g));
    cudaGraphExec_t e = nullptr;
    cudaError_t rc = cudaGraphInstantiate(&e, g, 0);
    cudaGraphDestroy(g);
    if (rc != cudaSuccess) return fail(c, JJ_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(rc));
    *graph_exec = e;
    return JJ_OK;
}
int32_t jj_graph_launch(jj_ctx* c, void* graph_exec) {
    if (!c || !graph_exec) return JJ_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaGraphLaunch((cudaGraphExec_t)graph_exec, c->stream));
    c->launches++;
    return JJ_OK;
}
int32_t jj_graph_destroy(jj_ctx* c, void* graph_exec) {
    if (!c || !graph_exec) return JJ_ERR_INVALID_ARG;
    CU(c, cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
    return JJ_OK;
}

int32_t jj_measure_imad_peak(jj_ctx* c, double* imad_per_sec) {
    if (!c || !imad_per_sec) return JJ_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    int32_t rc = ensure(c, &c->tmp2, &c->tmp2_cap, 64);
    if (rc) return rc;
    const int iters = 20000, blocks = c->sm_count * 4;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {  // first pass is warm-up
        CU(c, cudaEventRecord(c->ev0, c->stream));
        k_imad_peak<<<blocks, 256, 0, c->stream>>>((uint32_t*)c->tmp2, 12345u + rep, iters);
        CU(c, cudaEventRecord(c->ev1, c->stream));
        CU(c, cudaEventSynchronize(c->ev1));
        float ms = 0;
        CU(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (rep > 0) best = std::min(best, ms);
    }
    *imad_per_sec = (double)blocks * 256.0 * iters * 32.0 / (best * 1e-3);
    return JJ_OK;
}

// ---- field -----------------------------------------------------------------------------------------
#define FE_BIN(NAME, OP)                                                                                      \
    int32_t jj_fq_##NAME(jj_ctx* c, const void* a, const void* b, void* out, size_t n, uint32_t flags) {      \
        if (!b && n) return fail(c, JJ_ERR_INVALID_ARG, "null input pointer");                                \
        return fe_binary<FqP, OP>(c, a, b, out, n, flags);                                                    \
    }                                                                                                         \
    int32_t jj_fr_##NAME(jj_ctx* c, const void* a, const void* b, void* out, size_t n, uint32_t flags) {      \
        if (!b && n) return fail(c, JJ_ERR_INVALID_ARG, "null input pointer");                                \
        return fe_binary<FrP, OP>(c, a, b, out, n, flags);                                                    \
    }
#define FE_UN(NAME, OP)                                                                        \
    int32_t jj_fq_##NAME(jj_ctx* c, const void* a, void* out, size_t n, uint32_t flags) {      \
        return fe_binary<FqP, OP>(c, a, nullptr, out, n, flags);                               \
    }                                                                                          \
    int32_t jj_fr_##NAME(jj_ctx* c, const void* a, void* out, size_t n, uint32_t flags) {      \
        return fe_binary<FrP, OP>(c, a, nullptr, out, n, flags);                               \
    }
FE_BIN(mul, FE_MUL)
FE_BIN(add, FE_ADD)
FE_BIN(sub, FE_SUB)
FE_UN(square, FE_SQR)
FE_UN(neg, FE_NEG)
FE_UN(double, FE_DBL)

int32_t jj_fq_invert(jj_ctx* c, const void* a, void* out, uint8_t* ok, size_t n, uint32_t flags) {
    return fe_invert_batch<FqP>(c, a, out, ok, n, flags);
}
int32_t jj_fr_invert(jj_ctx* c, const void* a, void* out, uint8_t* ok, size_t n, uint32_t flags) {
    return fe_invert_batch<FrP>(c, a, out, ok, n, flags);
}
int32_t jj_fq_sqrt(jj_ctx* c, const void* a, void* out, uint8_t* ok, size_t n, uint32_t flags) {
    return fe_with_ok<FqP, FE_SQRT>(c, a, out, ok, n, flags);
}
int32_t jj_fr_sqrt(jj_ctx* c, const void* a, void* out, uint8_t* ok, size_t n, uint32_t flags) {
    return fe_with_ok<FrP, FE_SQRT>(c, a, out, ok, n, flags);
}
int32_t jj_fq_to_bytes(jj_ctx* c, const void* a, void* out, size_t n, uint32_t flags) {
    return fe_binary<FqP, FE_TO_BYTES>(c, a, nullptr, out, n, flags & ~JJ_CANON);
}
int32_t jj_fr_to_bytes(jj_ctx* c, const void* a, void* out, size_t n, uint32_t flags) {
    return fe_binary<FrP, FE_TO_BYTES>(c, a, nullptr, out, n, flags & ~JJ_CANON);
}
int32_t jj_fq_from_bytes(jj_ctx* c, const void* in, void* out, uint8_t* ok, size_t n, uint32_t flags) {
    return fe_with_ok<FqP, FE_FROM_BYTES>(c, in, out, ok, n, flags & ~JJ_CANON);
}
int32_t jj_fr_from_bytes(jj_ctx* c, const void* in, void* out, uint8_t* ok, size_t n, uint32_t flags) {
    return fe_with_ok<FrP, FE_FROM_BYTES>(c, in, out, ok, n, flags & ~JJ_CANON);
}
}  // extern "C"
template <class F>
static int32_t from_wide(jj_ctx* c, const void* in, void* out, size_t n, uint32_t flags) {
    In ins[3] = {{in, 64}, {nullptr, 0}, {nullptr, 0}};
    Out outs[2] = {{out, 32}, {nullptr, 0}};
    bool canon = flags & JJ_CANON;
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging*) {
        return fe_launch<F, FE_FROM_WIDE>(c, s, canon, din[0], nullptr, dout[0], nullptr, cnt, 0, 0);
    });
}

extern "C" {int32_t jj_fq_from_bytes_wide(jj_ctx* c, const void* in, void* out, size_t n, uint32_t flags) {
    return from_wide<FqP>(c, in, out, n, flags);
}
int32_t jj_fr_from_bytes_wide(jj_ctx* c, const void* in, void* out, size_t n, uint32_t flags) {
    return from_wide<FrP>(c, in, out, n, flags);
}
}  // extern "C"
template <class F>
static int32_t stream_gen(jj_ctx* c, uint64_t seed, size_t first, void* out, size_t n, uint32_t flags) {
    In ins[3] = {{nullptr, 0}, {nullptr, 0}, {nullptr, 0}};
    Out outs[2] = {{out, 32}, {nullptr, 0}};
    bool canon = flags & JJ_CANON;
    size_t base = first;
    size_t* progress = new size_t(0);
    int32_t rc = run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char**, char** dout, size_t cnt, Staging*) {
        size_t off = *progress;
        *progress += cnt;
        return fe_launch<F, FE_STREAM>(c, s, canon, nullptr, nullptr, dout[0], nullptr, cnt, seed, base + off);
    });
    delete progress;
    return rc;
}

extern "C" {int32_t jj_fq_stream(jj_ctx* c, uint64_t seed, size_t first, void* out, size_t n, uint32_t flags) {
    return stream_gen<FqP>(c, seed, first, out, n, flags);
}
int32_t jj_fr_stream(jj_ctx* c, uint64_t seed, size_t first, void* out, size_t n, uint32_t flags) {
    return stream_gen<FrP>(c, seed, first, out, n, flags);
}

// ---- points ------------------------------------------------------------------------------------------
int32_t jj_point_double(jj_ctx* c, const void* p, void* out, size_t n, uint32_t flags) {
    return pt_binary<PT_DBL>(c, p, 160, nullptr, 0, out, 160, n, flags);
}
int32_t jj_point_add(jj_ctx* c, const void* p, const void* q, void* out, size_t n, uint32_t flags) {
    if (!q && n) return fail(c, JJ_ERR_INVALID_ARG, "null input pointer");
    return pt_binary<PT_ADD>(c, p, 160, q, 160, out, 160, n, flags);
}
int32_t jj_point_add_niels(jj_ctx* c, const void* p, const void* q, void* out, size_t n, uint32_t flags) {
    if (!q && n) return fail(c, JJ_ERR_INVALID_ARG, "null input pointer");
    return pt_binary<PT_ADD_NIELS>(c, p, 160, q, 128, out, 160, n, flags);
}
int32_t jj_point_add_affine_niels(jj_ctx* c, const void* p, const void* q, void* out, size_t n, uint32_t flags) {
    if (!q && n) return fail(c, JJ_ERR_INVALID_ARG, "null input pointer");
    return pt_binary<PT_ADD_AFFINE_NIELS>(c, p, 160, q, 96, out, 160, n, flags);
}
int32_t jj_point_to_niels(jj_ctx* c, const void* p, void* out, size_t n, uint32_t flags) {
    return pt_binary<PT_TO_NIELS>(c, p, 160, nullptr, 0, out, 128, n, flags);
}
int32_t jj_affine_to_niels(jj_ctx* c, const void* p, void* out, size_t n, uint32_t flags) {
    return pt_binary<PT_AFFINE_TO_NIELS>(c, p, 64, nullptr, 0, out, 96, n, flags);
}

int32_t jj_scalar_mul(jj_ctx* c, const void* points, const void* scalars, void* out, size_t n, uint32_t flags) {
    if (!c) return JJ_ERR_INVALID_ARG;
    if (flags & JJ_CANON) return fail(c, JJ_ERR_INVALID_ARG, "point entry points take Montgomery-form coordinates");
    In ins[3] = {{points, 160}, {scalars, 32}, {nullptr, 0}};
    Out outs[2] = {{out, out_unit(flags)}, {nullptr, 0}};
    bool smont = flags & JJ_SCALAR_MONT, conv = flags & (JJ_OUT_AFFINE | JJ_OUT_BYTES);
    // host batches are staged in chunks of whole "rounds" (one unit per resident thread): a chunk that ends in a
    // partly filled round leaves the multiplier pipe under-occupied for that round
    size_t chunk = kChunkUnits;
    {
        int v = c->smul_variant > 0 && c->smul_variant < kNumVariants ? c->smul_variant : kDefaultVariant;
        size_t round = (size_t)c->sm_count * kVariants[v].threads * kVariants[v].min_blocks;
        if (round && round <= kChunkUnits) chunk = kChunkUnits / round * round;
    }
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging* S) -> int32_t {
        char** tbl = S ? &S->tbl : &c->tbl;
        size_t* tcap = S ? &S->tbl_cap : &c->tbl_cap;
        if (!conv) return launch_smul(c, s, din[0], din[1], 32, dout[0], nullptr, cnt, tbl, tcap, smont);
        // extended results go to scratch, then normalise (and encode) into the caller's buffer
        char** tmp = S ? &S->buf[2] : &c->tmp;
        size_t* tmpcap = S ? &S->cap[2] : &c->tmp_cap;
        int32_t rc = ensure(c, tmp, tmpcap, std::max(cnt, S ? kChunkUnits : cnt) * 160);
        if (rc) return rc;
        rc = launch_smul(c, s, din[0], din[1], 32, *tmp, nullptr, cnt, tbl, tcap, smont);
        if (rc) return rc;
        return finish_output(c, s, *tmp, dout[0], cnt, flags, S ? &S->tmp2 : &c->tmp2, S ? &S->tmp2_cap : &c->tmp2_cap);
    }, chunk);
}

int32_t jj_scalar_mul_encoded(jj_ctx* c, const void* points32, const void* scalars, void* out, uint8_t* ok, size_t n,
                              uint32_t flags) {
    if (!c) return JJ_ERR_INVALID_ARG;
    if (flags & JJ_CANON) return fail(c, JJ_ERR_INVALID_ARG, "point entry points take Montgomery-form coordinates");
    In ins[3] = {{points32, 32}, {scalars, 32}, {nullptr, 0}};
    Out outs[2] = {{out, out_unit(flags)}, {ok, 1}};
    const bool smont = flags & JJ_SCALAR_MONT, conv = flags & (JJ_OUT_AFFINE | JJ_OUT_BYTES), zip216 = !(flags & JJ_PRE_ZIP216);
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging* S) -> int32_t {
        // decode -> affine scratch -> extended scratch -> scalar-mul (in place when the output is converted)
        char** tbl = S ? &S->tbl : &c->tbl;
        size_t* tcap = S ? &S->tbl_cap : &c->tbl_cap;
        char** ext = S ? &S->buf[2] : &c->tmp;
        size_t* extcap = S ? &S->cap[2] : &c->tmp_cap;
        char** aff = S ? &S->tmp2 : &c->tmp2;
        size_t* affcap = S ? &S->tmp2_cap : &c->tmp2_cap;
        const size_t cap_units = std::max(cnt, S ? kChunkUnits : cnt);
        int32_t rc = ensure(c, ext, extcap, cap_units * 160);
        if (rc) return rc;
        rc = ensure(c, aff, affcap, cap_units * 64);
        if (rc) return rc;
        rc = from_bytes_launch(c, s, din[0], *aff, (uint8_t*)dout[1], cnt, zip216);
        if (rc) return rc;
        k_affine_to_extended<<<grid_for(c, cnt, 256, 8), 256, 0, s>>>(*aff, *ext, cnt);
        c->launches++;
        CU(c, cudaGetLastError());
        rc = launch_smul(c, s, *ext, din[1], 32, conv ? *ext : dout[0], nullptr, cnt, tbl, tcap, smont);
        if (rc || !conv) return rc;
        return finish_output(c, s, *ext, dout[0], cnt, flags, aff, affcap);
    });
}

int32_t jj_scalar_mul_fixed(jj_ctx* c, const void* base_affine, const void* scalars, void* out, size_t n, uint32_t flags) {
    if (!c) return JJ_ERR_INVALID_ARG;
    if (!base_affine) return fail(c, JJ_ERR_INVALID_ARG, "null base point");
    if (flags & JJ_CANON) return fail(c, JJ_ERR_INVALID_ARG, "point entry points take Montgomery-form coordinates");
    CU(c, cudaSetDevice(c->device));
    int32_t rc0 = build_fixed_table(c, base_affine, flags);
    if (rc0) return rc0;
    In ins[3] = {{scalars, 32}, {nullptr, 0}, {nullptr, 0}};
    Out outs[2] = {{out, out_unit(flags)}, {nullptr, 0}};
    bool smont = flags & JJ_SCALAR_MONT, conv = flags & (JJ_OUT_AFFINE | JJ_OUT_BYTES);
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging* S) -> int32_t {
        char* dst = dout[0];
        char** tmp = S ? &S->buf[2] : &c->tmp;
        size_t* tmpcap = S ? &S->cap[2] : &c->tmp_cap;
        if (conv) {
            int32_t rc = ensure(c, tmp, tmpcap, std::max(cnt, S ? kChunkUnits : cnt) * 160);
            if (rc) return rc;
            dst = *tmp;
        }
        int32_t rc = fixed_w(c) == 4 ? launch_fixed<256, 4, false>(c, s, din[0], dst, cnt, smont)
                                     : launch_fixed<512, 7, true>(c, s, din[0], dst, cnt, smont);
        if (rc) return rc;
        if (conv) return finish_output(c, s, dst, dout[0], cnt, flags, S ? &S->tmp2 : &c->tmp2, S ? &S->tmp2_cap : &c->tmp2_cap);
        return JJ_OK;
    });
}

int32_t jj_batch_normalize(jj_ctx* c, const void* in, void* out, size_t n, uint32_t flags) {
    if (!c) return JJ_ERR_INVALID_ARG;
    if (in == out && n) return fail(c, JJ_ERR_INVALID_ARG, "batch_normalize output must not alias its input");
    In ins[3] = {{in, 160}, {nullptr, 0}, {nullptr, 0}};
    Out outs[2] = {{out, 64}, {nullptr, 0}};
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging*) {
        return normalize_launch(c, s, din[0], dout[0], cnt);
    });
}
int32_t jj_affine_to_bytes(jj_ctx* c, const void* in, void* out, size_t n, uint32_t flags) {
    In ins[3] = {{in, 64}, {nullptr, 0}, {nullptr, 0}};
    Out outs[2] = {{out, 32}, {nullptr, 0}};
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging*) {
        return to_bytes_launch(c, s, din[0], dout[0], cnt);
    });
}

int32_t jj_batch_from_bytes(jj_ctx* c, const void* in, void* out, uint8_t* ok, size_t n, uint32_t flags) {
    if (!c) return JJ_ERR_INVALID_ARG;
    In ins[3] = {{in, 32}, {nullptr, 0}, {nullptr, 0}};
    Out outs[2] = {{out, 64}, {ok, 1}};
    bool zip216 = !(flags & JJ_PRE_ZIP216);
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging*) -> int32_t {
        return from_bytes_launch(c, s, din[0], dout[0], (uint8_t*)dout[1], cnt, zip216);
    });
}

int32_t jj_is_torsion_free(jj_ctx* c, const void* p, uint8_t* flags_out, size_t n, uint32_t flags) {
    if (!c) return JJ_ERR_INVALID_ARG;
    if (!flags_out && n) return fail(c, JJ_ERR_INVALID_ARG, "null output pointer");
    CU(c, cudaSetDevice(c->device));
    // [r]P == identity with r = FR_MODULUS_BYTES (src/lib.rs:73-76, 709-711); r is the same for every unit, so
    // the batch shares its width-5 NAF (42 additions instead of the 58 of the per-unit signed radix-16 windows)
    static const NafDigits naf = [] {
        const uint32_t r_words[8] = {FrP::M0, FrP::M1, FrP::M2, FrP::M3, FrP::M4, FrP::M5, FrP::M6, FrP::M7};
        NafDigits d;
        wnaf5_recode(d, r_words);
        return d;
    }();
    In ins[3] = {{p, 160}, {nullptr, 0}, {nullptr, 0}};
    Out outs[2] = {{flags_out, 1}, {nullptr, 0}};
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging* S) -> int32_t {
        char** tbl = S ? &S->tbl : &c->tbl;
        size_t* tcap = S ? &S->tbl_cap : &c->tbl_cap;
        constexpr int T = 512;
        int grid = grid_for(c, cnt, T, 1);
        int32_t rc = ensure(c, tbl, tcap, (size_t)grid * (T / 32) * 32768);
        if (rc) return rc;
        SmulArgs a{};
        a.points = din[0];
        a.flag_out = (uint8_t*)dout[0];
        a.n = cnt;
        a.tbl_scratch = *tbl;
        k_scalar_mul_const<T><<<grid, T, 0, s>>>(a, naf);
        c->launches++;
        CU(c, cudaGetLastError());
        return JJ_OK;
    });
}
}  // extern "C"
template <int WHAT>
static int32_t point_flag(jj_ctx* c, const void* p, uint8_t* flags_out, size_t n, uint32_t flags) {
    if (!c) return JJ_ERR_INVALID_ARG;
    if (!flags_out && n) return fail(c, JJ_ERR_INVALID_ARG, "null output pointer");
    In ins[3] = {{p, 160}, {nullptr, 0}, {nullptr, 0}};
    Out outs[2] = {{flags_out, 1}, {nullptr, 0}};
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging*) -> int32_t {
        k_point_flag<WHAT><<<grid_for(c, cnt, 128, 4), 128, 0, s>>>(din[0], (uint8_t*)dout[0], cnt);
        c->launches++;
        CU(c, cudaGetLastError());
        return JJ_OK;
    });
}

extern "C" {int32_t jj_is_identity(jj_ctx* c, const void* p, uint8_t* f, size_t n, uint32_t flags) { return point_flag<0>(c, p, f, n, flags); }
int32_t jj_is_small_order(jj_ctx* c, const void* p, uint8_t* f, size_t n, uint32_t flags) { return point_flag<1>(c, p, f, n, flags); }

// ---- multi-GPU -----------------------------------------------------------------------------------------
int32_t jj_comm_unique_id(void* id128) {
    if (!id128) return JJ_ERR_INVALID_ARG;
    if (!load_nccl()) return JJ_ERR_NCCL;
    return g_nccl.GetUniqueId(id128) == 0 ? JJ_OK : JJ_ERR_NCCL;
}
int32_t jj_comm_init(jj_ctx* c, int32_t nranks, int32_t rank, const void* id128) {
    if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return JJ_ERR_INVALID_ARG;
    if (!load_nccl()) return fail(c, JJ_ERR_NCCL, "libnccl.so.2 not found");
    CU(c, cudaSetDevice(c->device));
    Id128 id;
    memcpy(id.b, id128, 128);
    int rc = g_nccl.CommInitRank(&c->nccl_comm, nranks, id, rank);
    if (rc != 0) return fail(c, JJ_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    c->nranks = nranks;
    c->rank = rank;
    return JJ_OK;
}
int32_t jj_comm_destroy(jj_ctx* c) {
    if (!c) return JJ_ERR_INVALID_ARG;
    if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
    c->nccl_comm = nullptr;
    c->nranks = 1;
    c->rank = 0;
    return JJ_OK;
}
int32_t jj_scalar_mul_sharded(jj_ctx* c, const void* points_local, const void* scalars_local, void* out_all,
                              size_t n_local, uint32_t flags) {
    if (!c || !out_all) return JJ_ERR_INVALID_ARG;
    if (!(flags & JJ_DEVICE_PTRS)) return fail(c, JJ_ERR_INVALID_ARG, "jj_scalar_mul_sharded takes device pointers");
    CU(c, cudaSetDevice(c->device));
    size_t unit = out_unit(flags);
    bool fused = c->peers.n_peers == c->nranks && c->nranks > 1 && unit == 160;
    if (fused) {
        // compute + all-gather in ONE kernel: results are stored into every rank's gathered buffer
        if (c->peers.ptr[c->rank] != (char*)out_all) return fail(c, JJ_ERR_INVALID_ARG, "out_all is not the registered peer buffer");
        if (((uintptr_t)points_local | (uintptr_t)scalars_local | (uintptr_t)out_all) & 31)
            return fail(c, JJ_ERR_INVALID_ARG, "device pointer not 32-byte aligned");
        PeerOut po = c->peers;
        po.base_unit = (size_t)c->rank * n_local;
        int32_t rc = launch_smul(c, c->stream, (const char*)points_local, (const char*)scalars_local, 32, nullptr, nullptr,
                                 n_local, &c->tbl, &c->tbl_cap, flags & JJ_SCALAR_MONT, &po);
        if (rc) return rc;
        // stream-ordered rendezvous: once this 4-byte all-reduce completes here, every peer's kernel
        // (enqueued before its own all-reduce) has finished storing into this rank's buffer
        if (!c->barrier_word) {
            CU(c, cudaMalloc((void**)&c->barrier_word, 256));
            CU(c, cudaMemset(c->barrier_word, 0, 256));
        }
        int nrc = g_nccl.AllReduce(c->barrier_word, c->barrier_word + 32, 1, /*ncclInt32*/ 2, /*ncclSum*/ 0, c->nccl_comm, c->stream);
        if (nrc != 0) return fail(c, JJ_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "?");
    } else {
        char* mine = (char*)out_all + (size_t)c->rank * n_local * unit;
        int32_t rc = jj_scalar_mul(c, points_local, scalars_local, mine, n_local, flags | JJ_ASYNC);
        if (rc) return rc;
        if (c->nranks > 1) {
            if (!c->nccl_comm) return fail(c, JJ_ERR_NCCL, "jj_comm_init was not called");
            // in-place all-gather: every rank's block already sits at its own offset
            int nrc = g_nccl.AllGather(mine, out_all, n_local * unit, /*ncclInt8*/ 0, c->nccl_comm, c->stream);
            if (nrc != 0) return fail(c, JJ_ERR_NCCL, "ncclAllGather: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "?");
        }
    }
    if (!(flags & JJ_ASYNC)) CU(c, cudaStreamSynchronize(c->stream));
    return JJ_OK;
}

int32_t jj_ipc_export(jj_ctx* c, const void* dptr, void* handle64) {
    if (!c || !dptr || !handle64) return JJ_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    CU(c, cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, (void*)dptr));
    return JJ_OK;
}
int32_t jj_ipc_open(jj_ctx* c, const void* handle64, void** dptr) {
    if (!c || !dptr || !handle64) return JJ_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CU(c, cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return JJ_OK;
}
int32_t jj_ipc_close(jj_ctx* c, void* dptr) {
    if (!c || !dptr) return JJ_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaIpcCloseMemHandle(dptr));
    return JJ_OK;
}
int32_t jj_comm_set_peer_outputs(jj_ctx* c, void* const* peer_out_all, int32_t count) {
    if (!c) return JJ_ERR_INVALID_ARG;
    if (!peer_out_all || count == 0) {
        c->peers.n_peers = 0;
        return JJ_OK;
    }
    if (count != c->nranks || count > 8) return fail(c, JJ_ERR_INVALID_ARG, "need one pointer per rank (<= 8)");
    for (int r = 0; r < count; r++) {
        if (!peer_out_all[r] || ((uintptr_t)peer_out_all[r] & 31)) return fail(c, JJ_ERR_INVALID_ARG, "bad peer pointer");
        c->peers.ptr[r] = (char*)peer_out_all[r];
    }
    c->peers.n_peers = count;
    return JJ_OK;
}

}  // extern "C"
