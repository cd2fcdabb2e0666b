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
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "kernels.cuh"

using namespace jj;

struct Id128 {  // ncclUniqueId, passed by value
    char b[128];
};

namespace {

constexpr size_t kChunkUnits = 1u << 17;  // default host-staging chunk (units); staging buffers grow to the chunk in use
constexpr int kStages = 2;

struct Staging {
    cudaStream_t stream = nullptr;
    cudaStream_t aux = nullptr;   // second copy stream: an input the first kernel of the chunk does not read (BatchOpts::late_in)
    cudaEvent_t late = nullptr;   // ... has arrived
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
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
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
        g_nccl.Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclBroadcast");
        g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
        g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
        g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
        g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
        // every entry point the library calls must resolve: a partial libnccl is treated as absent
        if (g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllGather && g_nccl.AllReduce && g_nccl.Broadcast &&
            g_nccl.GroupStart && g_nccl.GroupEnd && g_nccl.CommDestroy) {
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
    char* norm = nullptr;  // scratch of the fused normalise epilogue (k_scalar_mul<.., NORM>)
    size_t norm_cap = 0;
    int live_graphs = 0;  // graphs captured on this context and not yet destroyed: their nodes hold scratch pointers
    cudaEvent_t ev_fork = nullptr;
    cudaEvent_t ev_join[kStages] = {nullptr, nullptr};
    uint32_t* fixed_table = nullptr;  // fixed-base window table of the cached base (FixedGeom<fixed_w>)
    size_t fixed_table_cap = 0;
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

bool capturing(jj_ctx* c) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    return cudaStreamIsCapturing(c->stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone;
}
// Grows a scratch buffer.  Buffers used by calls on the context's main stream can be baked into captured
// graphs (jj_graph_*): they are never reallocated while a capture is in progress (cudaMalloc / cudaFree are
// illegal there) or while a captured graph is alive (its kernel nodes would write through the stale pointer).
int32_t ensure(jj_ctx* c, char** buf, size_t* cap, size_t bytes, bool main_scratch = true) {
    if (*cap >= bytes && *buf) return JJ_OK;
    if (!main_scratch) {  // staging-stream buffers: host-pointer calls are never captured
        if (*buf) CU(c, cudaFree(*buf));
        *buf = nullptr;
        *cap = 0;
        CU(c, cudaMalloc((void**)buf, bytes));
        *cap = bytes;
        return JJ_OK;
    }
    if (capturing(c))
        return fail(c, JJ_ERR_INVALID_ARG, "scratch would have to grow during graph capture: run the sequence once eagerly first");
    if (c->live_graphs > 0)
        return fail(c, JJ_ERR_INVALID_ARG, "scratch would be reallocated while %d captured graph(s) still refer to it: "
                    "destroy them (jj_graph_destroy) or run the larger batch before capturing", c->live_graphs);
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

// ---- scalar-mul mappings (jj_set_scalar_mul_variant) --------------------------------------------
// The default build ships the default mapping and two A/B mappings; the experiments of round 1 (shared-memory
// tables, the slot-file kernel, other block shapes; all measured slower) compile only
// with -DJJ_EXPERIMENTS.
struct SmulVariant {
    int id, threads, min_blocks, table;
};
const SmulVariant kVariants[] = {
    {13, 512, 1, TABLE_GMEM},  // default: 16 warps/SM in one block (<= 128 registers), table in L2-resident scratch
    {24, 768, 1, TABLE_GMEM},  // 24 warps/SM (<= 80 registers): +1.3 % on resident batches, but its tables leave L2
    {5, 256, 1, TABLE_GMEM},   // 8 warps/SM (<= 255 registers)
#if defined(JJ_EXPERIMENTS)
    {1, 224, 1, TABLE_SMEM},  {2, 128, 2, TABLE_GMEM},  {3, 128, 3, TABLE_GMEM},  {4, 128, 4, TABLE_GMEM},
    {6, 192, 1, TABLE_SMEM},  {7, 128, 1, TABLE_SMEM},  {8, 64, 3, TABLE_SMEM},   {9, 96, 4, TABLE_GMEM},
    {10, 64, 6, TABLE_GMEM},  {11, 384, 1, TABLE_GMEM}, {12, 192, 2, TABLE_GMEM}, {14, 320, 1, TABLE_GMEM},
    {15, 512, 1, 2},          {16, 384, 1, 2},          {17, 544, 1, 2},          {18, 256, 2, 2},
    {19, 448, 1, TABLE_GMEM}, {20, 480, 1, TABLE_GMEM}, {21, 256, 2, TABLE_GMEM}, {22, 576, 1, TABLE_GMEM},
    {23, 640, 1, TABLE_GMEM}, {25, 704, 1, TABLE_GMEM}, {26, 896, 1, TABLE_GMEM}, {27, 1024, 1, TABLE_GMEM},
#endif
};
constexpr int kDefaultVariant = 13;
// knob values outside the mapping table
constexpr int kFixedW4 = 100;         // fixed-base kernel with 4-bit windows (47 KB table in shared memory)
constexpr int kFixedW7 = 107;         // fixed-base kernel with 7-bit windows (216 KB table in shared memory, TMA-staged)
constexpr int kFixedW16 = 116;        // fixed-base kernel with 16-bit windows (50 MB table in global memory: for long-lived bases)
constexpr int kFixedWide = 12;        // default: 12-bit windows, 4.1 MB table in global memory (L2 / L1 resident)
constexpr int kFusedNormOn = 200;     // converted outputs: always use the fused normalise epilogue (k_scalar_mul<.., NORM>)
constexpr int kFusedNormOff = 201;    // ... never (separate k_batch_normalize pass)

const SmulVariant& smul_variant(const jj_ctx* c) {
    for (const SmulVariant& v : kVariants)
        if (v.id == c->smul_variant) return v;
    return kVariants[0];
}
// units one launch keeps resident (one per thread): host batches are staged in whole multiples of it
size_t smul_round(const jj_ctx* c) {
    const SmulVariant& v = smul_variant(c);
    return (size_t)c->sm_count * v.threads * v.min_blocks;
}

template <int T, int MB, int TAB, bool NORM, bool CT = false>
int32_t launch_smul_t(jj_ctx* c, cudaStream_t s, SmulArgs a, char** tbl, size_t* tbl_cap) {
    auto kern = k_scalar_mul<T, MB, TAB, NORM, CT>;
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
#if defined(JJ_EXPERIMENTS)
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
#endif
// a.out_unit = 160: ExtendedPoint results.  64 / 32: the fused normalise epilogue (default mapping only); the caller
// provides a.norm_scratch (norm_scratch_bytes()).
size_t norm_scratch_bytes(const jj_ctx* c, size_t n) { return n * 128 + smul_round(c) * 32; }
int32_t launch_smul(jj_ctx* c, cudaStream_t s, SmulArgs a, char** tbl, size_t* tbl_cap, bool const_time = false) {
    const SmulVariant& v = smul_variant(c);
    if (const_time) {
        if (a.out_unit != 160 || v.id != kDefaultVariant)
            return fail(c, JJ_ERR_INVALID_ARG, "JJ_CONST_TIME runs the default mapping with ExtendedPoint results (convert afterwards)");
        return launch_smul_t<512, 1, TABLE_GMEM, false, true>(c, s, a, tbl, tbl_cap);
    }
    if (a.out_unit != 160) {
        if (v.id != kDefaultVariant) return fail(c, JJ_ERR_INVALID_ARG, "the fused normalise epilogue exists for the default mapping only");
        return launch_smul_t<512, 1, TABLE_GMEM, true>(c, s, a, tbl, tbl_cap);
    }
#define V(ID, T, MB, TAB) \
    case ID: return launch_smul_t<T, MB, TAB, false>(c, s, a, tbl, tbl_cap)
    switch (v.id) {
        V(13, 512, 1, TABLE_GMEM);
        V(24, 768, 1, TABLE_GMEM);
        V(5, 256, 1, TABLE_GMEM);
#if defined(JJ_EXPERIMENTS)
        V(1, 224, 1, TABLE_SMEM);
        V(2, 128, 2, TABLE_GMEM);
        V(3, 128, 3, TABLE_GMEM);
        V(4, 128, 4, TABLE_GMEM);
        V(6, 192, 1, TABLE_SMEM);
        V(7, 128, 1, TABLE_SMEM);
        V(8, 64, 3, TABLE_SMEM);
        V(9, 96, 4, TABLE_GMEM);
        V(10, 64, 6, TABLE_GMEM);
        V(11, 384, 1, TABLE_GMEM);
        V(12, 192, 2, TABLE_GMEM);
        V(14, 320, 1, TABLE_GMEM);
        V(19, 448, 1, TABLE_GMEM);
        V(20, 480, 1, TABLE_GMEM);
        V(21, 256, 2, TABLE_GMEM);
        V(22, 576, 1, TABLE_GMEM);
        V(23, 640, 1, TABLE_GMEM);
        V(25, 704, 1, TABLE_GMEM);
        V(26, 896, 1, TABLE_GMEM);
        V(27, 1024, 1, TABLE_GMEM);
        case 15: return launch_smul_slots<512, 1>(c, s, a, tbl, tbl_cap);
        case 16: return launch_smul_slots<384, 1>(c, s, a, tbl, tbl_cap);
        case 17: return launch_smul_slots<544, 1>(c, s, a, tbl, tbl_cap);
        case 18: return launch_smul_slots<256, 2>(c, s, a, tbl, tbl_cap);
#endif
    }
#undef V
    return fail(c, JJ_ERR_INVALID_ARG, "bad scalar-mul variant %d", v.id);
}
SmulArgs smul_args(const char* pts, bool in_affine, const char* sc, char* out, size_t n, bool scalar_mont) {
    SmulArgs a{};
    a.points = pts;
    a.in_affine = in_affine;
    a.scalars = sc;
    a.scalar_stride = 32;
    a.out = out;
    a.n = n;
    a.scalar_mont = scalar_mont;
    a.out_unit = 160;
    return a;
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
// Device alias of a pinned, mapped host range (cudaHostAlloc / cudaHostRegister under unified addressing), or nullptr.
char* mapped_alias(const void* p, size_t bytes) {
    if (!p || !bytes || ((uintptr_t)p & 31)) return nullptr;
    cudaPointerAttributes a0{}, a1{};
    if (cudaPointerGetAttributes(&a0, p) != cudaSuccess || cudaPointerGetAttributes(&a1, (const char*)p + bytes - 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (a0.type != cudaMemoryTypeHost || a1.type != cudaMemoryTypeHost || !a0.devicePointer) return nullptr;
    return (char*)a0.devicePointer;
}
// How a host-pointer call is staged.
struct BatchOpts {
    size_t chunk_units = kChunkUnits;  // units per staged chunk (the staging buffers grow to it)
    bool direct_in0 = false;   // ins[0]: when the caller's buffer is pinned, the kernel reads it in place over PCIe (coalesced
                               // 32-byte units, a compute-bound consumer) instead of waiting for an upload
    bool direct_out0 = false;  // outs[0]: ... the producing kernel stores into it in place (write-only 32-byte units)
    int late_in = -1;          // this input is uploaded on the stage's second stream while the chunk's first kernel runs; the
                               // launch callback waits for Staging::late before the kernel that reads it
};
BatchOpts chunked(size_t units) {
    BatchOpts o;
    o.chunk_units = units;
    return o;
}
// Launch(stream, din[3], dout[2], count, staging_or_null) -> status
template <class Launch>
int32_t run_batch(jj_ctx* c, uint32_t flags, size_t n, const In (&ins)[3], const Out (&outs)[2], Launch launch,
                  const BatchOpts& opts = BatchOpts()) {
    const size_t chunk_units = opts.chunk_units;
    if (!c) return JJ_ERR_INVALID_ARG;
    for (const In& i : ins)
        if (i.unit && !i.p && n) return fail(c, JJ_ERR_INVALID_ARG, "null input pointer");
    if (outs[0].unit && !outs[0].p && n) return fail(c, JJ_ERR_INVALID_ARG, "null output pointer");
    CU(c, cudaSetDevice(c->device));
    if (capturing(c) && (flags & (JJ_DEVICE_PTRS | JJ_ASYNC)) != (JJ_DEVICE_PTRS | JJ_ASYNC))
        return fail(c, JJ_ERR_INVALID_ARG, "only JJ_DEVICE_PTRS | JJ_ASYNC calls can be captured into a graph");
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
    char* const alias_in0 = opts.direct_in0 && ins[0].unit ? mapped_alias(ins[0].p, n * ins[0].unit) : nullptr;
    char* const alias_out0 = opts.direct_out0 && outs[0].unit && outs[0].p ? mapped_alias(outs[0].p, n * outs[0].unit) : nullptr;
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
            if (k == 0 && alias_in0) {
                din[0] = alias_in0 + done * ins[0].unit;
                continue;
            }
            int32_t rc = ensure(c, &S.buf[k], &S.cap[k], chunk_units * ins[k].unit, false);
            if (rc) return rc;
            const bool late = k == opts.late_in;
            CU(c, cudaMemcpyAsync(S.buf[k], (const char*)ins[k].p + done * ins[k].unit, cnt * ins[k].unit,
                                  cudaMemcpyHostToDevice, late ? S.aux : S.stream));
            if (late) CU(c, cudaEventRecord(S.late, S.aux));
            din[k] = S.buf[k];
        }
        {
            size_t need = 0;
            for (int k = 0; k < 2; k++) {
                out_off[k] = need;
                if (k == 0 && alias_out0) continue;
                need += (chunk_units * outs[k].unit + 255) & ~(size_t)255;
            }
            int32_t rc = ensure(c, &S.buf[3], &S.cap[3], std::max<size_t>(need, 256), false);
            if (rc) return rc;
            for (int k = 0; k < 2; k++)
                if (outs[k].unit && outs[k].p) dout[k] = S.buf[3] + out_off[k];
            if (alias_out0) dout[0] = alias_out0 + done * outs[0].unit;
        }
        int32_t rc = launch(S.stream, din, dout, cnt, &S);
        if (rc) return rc;
        for (int k = 0; k < 2; k++)
            if (outs[k].unit && outs[k].p && !(k == 0 && alias_out0))
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
// grid of the Montgomery-trick kernels: chains of `per_thread` elements per thread for large batches, in whole
// multiples of `blocks_per_sm` x SM count blocks of 128 threads
int chain_grid(const jj_ctx* c, size_t n, int per_thread, int blocks_per_sm) {
    size_t blocks = (n + 128 * (size_t)per_thread - 1) / (128 * (size_t)per_thread);
    size_t per_wave = (size_t)c->sm_count * blocks_per_sm;
    blocks = std::max<size_t>(1, (blocks + per_wave - 1) / per_wave * per_wave);
    return (int)std::min(blocks, (n + 127) / 128);
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
        int32_t rc = ensure(c, scr, cap, cnt * 32, !S);
        if (rc) return rc;
        const int blocks = chain_grid(c, cnt, 32, 2);
        if (canon)
            k_fe_invert_batched<F, true><<<blocks, 128, 0, s>>>(din[0], dout[0], (uint8_t*)dout[1], *scr, cnt);
        else
            k_fe_invert_batched<F, false><<<blocks, 128, 0, s>>>(din[0], dout[0], (uint8_t*)dout[1], *scr, cnt);
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

// ExtendedPoint (device) -> AffinePoint (fmt 64), encoding (fmt 32) or normalised ExtendedPoint (fmt 160, may be in
// place).  One Fermat inversion per thread amortised over its strided chain (~32 points per thread for large batches).
// fmt 32 / 160 need n x 32 B of scratch for the running products (fmt 64 keeps them in the output).
int32_t normalize_launch(jj_ctx* c, cudaStream_t s, const char* in, char* out, size_t n, int fmt, char** scr, size_t* scr_cap,
                         bool main_scratch) {
    const int blocks = chain_grid(c, n, 32, 2);
    if (fmt == 64) {
        k_batch_normalize<64><<<blocks, 128, 0, s>>>(in, out, out, n);
    } else {
        int32_t rc = ensure(c, scr, scr_cap, n * 32, main_scratch);
        if (rc) return rc;
        if (fmt == 32) k_batch_normalize<32><<<blocks, 128, 0, s>>>(in, out, *scr, n);
        else k_batch_normalize<160><<<blocks, 128, 0, s>>>(in, out, *scr, n);
    }
    c->launches++;
    CU(c, cudaGetLastError());
    return JJ_OK;
}
int32_t from_bytes_launch(jj_ctx* c, cudaStream_t s, const char* in, char* out, uint8_t* ok, size_t n, bool zip216) {
    // one thread per encoding (no inversion chain: the root of the quotient comes out of one power, fe.cuh); plain blocks
    // of 128, so the tail of the batch is balanced by the block scheduler
    const size_t blocks = std::min<size_t>((n + 127) / 128, (size_t)1 << 24);
    k_from_bytes<<<(unsigned)std::max<size_t>(1, blocks), 128, 0, s>>>(in, out, ok, n, zip216);
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
size_t out_unit(uint32_t flags) { return (flags & JJ_OUT_BYTES) ? 32 : (flags & JJ_OUT_AFFINE) ? 64 : 160; }

// Variable-base scalar multiplication of `cnt` device-resident units into `dst` in the format the flags ask for.
// ExtendedPoint output: one kernel.  AffinePoint / encoding output: either the kernel's fused normalise epilogue
// (when every thread owns several units, i.e. device-resident batches of >= 4 rounds) or the kernel into an
// ExtendedPoint scratch followed by one k_batch_normalize pass that writes the final format.
int32_t smul_any(jj_ctx* c, cudaStream_t s, Staging* S, const char* pts, bool in_affine, const char* sc, char* dst, size_t cnt,
                 uint32_t flags, const PeerOut* peers = nullptr) {
    char** tbl = S ? &S->tbl : &c->tbl;
    size_t* tcap = S ? &S->tbl_cap : &c->tbl_cap;
    const int unit = (int)out_unit(flags);
    const bool ct = flags & JJ_CONST_TIME;
    SmulArgs a = smul_args(pts, in_affine, sc, dst, cnt, flags & JJ_SCALAR_MONT);
    if (peers) a.peers = *peers;
    if (unit == 160) return launch_smul(c, s, a, tbl, tcap, ct);
    // The fused epilogue measured 1 % SLOWER than the separate pass on one GPU (34.63 vs 34.29 ms per 2^20, bench_ops r02a:
    // 75 776 Fermat inversions, one per resident thread, against 32 768 in k_batch_normalize), so it is used where it
    // pays -- the fused all-gather, which then moves 32-byte encodings instead of 160-byte points -- or on request.
    const bool default_map = smul_variant(c).id == kDefaultVariant;
    bool fused = c->smul_variant == kFusedNormOn && !S;
    if (ct) fused = false;  // constant-time mode: ExtendedPoint kernel, then the separate normalise pass
    if (peers && peers->n_peers > 0) {
        if (S || !default_map || ct)
            return fail(c, JJ_ERR_INVALID_ARG, "fused gather of converted outputs needs the default, variable-time mapping");
        fused = true;
    }
    if (fused) {
        int32_t rc = ensure(c, &c->norm, &c->norm_cap, norm_scratch_bytes(c, cnt));
        if (rc) return rc;
        a.out_unit = unit;
        a.norm_scratch = c->norm;
        return launch_smul(c, s, a, tbl, tcap);
    }
    // extended results go to scratch, then one pass normalises (and encodes) into the caller's buffer
    char** tmp = S ? &S->buf[2] : &c->tmp;
    size_t* tmpcap = S ? &S->cap[2] : &c->tmp_cap;
    int32_t rc = ensure(c, tmp, tmpcap, cnt * 160, !S);
    if (rc) return rc;
    a.out = *tmp;
    rc = launch_smul(c, s, a, tbl, tcap, ct);
    if (rc) return rc;
    // the running products of the normalise pass for 32-byte outputs live in the (dead) scalar-mul table scratch
    return normalize_launch(c, s, *tmp, dst, cnt, unit, tbl, tcap, !S);
}
// whole rounds per staged chunk: a chunk that ends in a partly filled round leaves the multiplier pipe
// under-occupied for that round.  `rounds` = 1 keeps the exposed first upload / last download small (the 352 B/unit
// ExtendedPoint path); the wire-format path moves 65 B/unit and wants chunks of several rounds instead, so that the
// normalise pass between two scalar-mul launches amortises its Fermat inversion over a chain of points (wire_rounds()).
size_t smul_chunk(const jj_ctx* c, int rounds = 1) {
    const size_t round = smul_round(c);
    return round ? round * rounds : kChunkUnits;
}

// Rounds per staged chunk of the wire-format path.  Measured per 2^20 units (pinned host buffers, scripts/wire_sweep.py)
// with the inversion-free decode: 1 round 42.0 ms, 2: 40.5, 4: 39.9, 8: 39.3-39.4 -- against 38.0 ms device-resident.  What a
// small chunk pays is the normalise pass (one Fermat inversion per thread for a chain of two points when the chunk is one
// round) between two scalar-mul launches that each fill the chip.  Chunks that ramp 1, 2, 4, 4, 2, 1 rounds (small first
// upload, small last download) were measured and lose for the same reason: 39.8 ms.
// JJ_WIRE_ROUNDS overrides (experiments).
int env_rounds(const char* name, int dflt) {
    const char* e = getenv(name);
    int v = e ? atoi(e) : 0;
    return v >= 1 && v <= 64 ? v : dflt;
}
int wire_rounds() {
    static const int r = env_rounds("JJ_WIRE_ROUNDS", 8);
    return r;
}
// JJ_WIRE_DIRECT=0: stage pinned buffers like pageable ones (A/B)
bool wire_direct() {
    static const bool d = [] {
        const char* e = getenv("JJ_WIRE_DIRECT");
        return !e || atoi(e) != 0;
    }();
    return d;
}


// fixed-base window width: 7 (216 KB table, the default) or 4 (47 KB table, variant 100)
int fixed_w(const jj_ctx* c) {
    return c->smul_variant == kFixedW4 ? 4 : c->smul_variant == kFixedW7 ? 7 : c->smul_variant == kFixedW16 ? 16 : kFixedWide;
}
size_t fixed_table_bytes(int w) {
    return w == 4 ? FixedGeom<4>::BYTES : w == 7 ? FixedGeom<7>::BYTES : w == 16 ? FixedGeom<16>::BYTES : FixedGeom<kFixedWide>::BYTES;
}

int32_t build_fixed_table(jj_ctx* c, const void* base_affine, uint32_t flags) {
    if (capturing(c))
        return fail(c, JJ_ERR_INVALID_ARG, "jj_scalar_mul_fixed cannot be captured into a graph (the table cache is checked on the host)");
    char key[64];
    if (flags & JJ_DEVICE_PTRS)
        CU(c, cudaMemcpy(key, base_affine, 64, cudaMemcpyDeviceToHost));
    else
        memcpy(key, base_affine, 64);
    const int w = fixed_w(c);
    if (c->fixed_valid && c->fixed_w == w && memcmp(key, c->fixed_base_key, 64) == 0) return JJ_OK;
    if (c->fixed_table && c->fixed_table_cap < fixed_table_bytes(w)) {
        CU(c, cudaFree(c->fixed_table));
        c->fixed_table = nullptr;
        c->fixed_valid = false;
    }
    if (!c->fixed_table) {
        CU(c, cudaMalloc((void**)&c->fixed_table, fixed_table_bytes(w)));
        c->fixed_table_cap = fixed_table_bytes(w);
    }
    if (!c->fixed_base_dev) CU(c, cudaMalloc((void**)&c->fixed_base_dev, 64));
    CU(c, cudaMemcpy(c->fixed_base_dev, key, 64, cudaMemcpyHostToDevice));
    const int entries = w == 4 ? FixedGeom<4>::ENTRIES : w == 7 ? FixedGeom<7>::ENTRIES : w == 16 ? FixedGeom<16>::ENTRIES
                                                                                                   : FixedGeom<kFixedWide>::ENTRIES;
    // one thread per entry up to 8 blocks of 64 threads per SM (the kernel grid-strides over larger tables)
    const int blocks = std::min((entries + 63) / 64, c->sm_count * 8);
    int32_t rc = ensure(c, &c->tbl, &c->tbl_cap, (size_t)(blocks * 2) * 32768);
    if (rc) return rc;
    if (w == 4)
        k_fixed_table_build<4><<<blocks, 64, 0, c->stream>>>(c->fixed_base_dev, c->fixed_table, c->tbl);
    else if (w == 7)
        k_fixed_table_build<7><<<blocks, 64, 0, c->stream>>>(c->fixed_base_dev, c->fixed_table, c->tbl);
    else if (w == 16)
        k_fixed_table_build<16><<<blocks, 64, 0, c->stream>>>(c->fixed_base_dev, c->fixed_table, c->tbl);
    else
        k_fixed_table_build<kFixedWide><<<blocks, 64, 0, c->stream>>>(c->fixed_base_dev, c->fixed_table, c->tbl);
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

template <int W>
int32_t launch_fixed_wide(jj_ctx* c, cudaStream_t s, const char* scalars, char* dst, size_t cnt, bool smont) {
    constexpr int T = 512;
    k_scalar_mul_fixed_gmem<T, W><<<grid_for(c, cnt, T, 1), T, 0, s>>>(c->fixed_table, scalars, dst, cnt, smont);
    c->launches++;
    CU(c, cudaGetLastError());
    return JJ_OK;
}

int32_t nccl_fail(jj_ctx* c, const char* what, int rc) {
    return fail(c, JJ_ERR_NCCL, "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
}
// 4-byte all-reduce on the context's stream: a stream-ordered rendezvous of all ranks
int32_t rendezvous(jj_ctx* c) {
    if (!c->barrier_word) {
        CU(c, cudaMalloc((void**)&c->barrier_word, 256));
        CU(c, cudaMemset(c->barrier_word, 0, 256));
    }
    int rc = g_nccl.AllReduce(c->barrier_word, c->barrier_word + 32, 1, /*ncclInt32*/ 2, /*ncclSum*/ 0, c->nccl_comm, c->stream);
    return rc == 0 ? JJ_OK : nccl_fail(c, "ncclAllReduce", rc);
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
    if (prop.major != 10 || prop.minor != 0) {
        delete c;
        return JJ_ERR_NO_DEVICE;  // the library holds sm_100a SASS only (no PTX): nothing else can run it
    }
    c->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int k = 0; k < kStages && ok; k++)
        ok = cudaStreamCreateWithFlags(&c->st[k].stream, cudaStreamNonBlocking) == cudaSuccess &&
             cudaStreamCreateWithFlags(&c->st[k].aux, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->st[k].late, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreate(&c->ev0) == cudaSuccess && cudaEventCreate(&c->ev1) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; k < kStages && ok; k++) ok = cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming) == cudaSuccess;
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
        if (c->st[k].aux) cudaStreamDestroy(c->st[k].aux);
        if (c->st[k].late) cudaEventDestroy(c->st[k].late);
    }
    for (void* p : {(void*)c->tbl, (void*)c->tmp, (void*)c->tmp2, (void*)c->norm, (void*)c->fixed_table,
                    (void*)c->fixed_base_dev, (void*)c->flush})
        if (p) cudaFree(p);
    for (cudaEvent_t e : {c->ev0, c->ev1, c->ev_fork, c->ev_join[0], c->ev_join[1]})
        if (e) cudaEventDestroy(e);
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
    if (!c) return JJ_ERR_INVALID_ARG;
    bool known = v == 0 || v == kFixedW4 || v == kFixedW7 || v == kFixedW16 || v == kFusedNormOn || v == kFusedNormOff;
    for (const SmulVariant& k : kVariants) known = known || k.id == v;
    if (!known) return fail(c, JJ_ERR_INVALID_ARG, "unknown scalar-mul variant %d (experimental mappings need a -DJJ_EXPERIMENTS build)", v);
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
    if (capturing(c)) return fail(c, JJ_ERR_INVALID_ARG, "a capture is already in progress");
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
    c->live_graphs++;  // scratch buffers stay where they are until the graph is destroyed (ensure())
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
    if (c->live_graphs > 0) c->live_graphs--;
    return JJ_OK;
}

int32_t jj_measure_imad_peak(jj_ctx* c, double* imad_per_sec) {
    if (!c || !imad_per_sec) return JJ_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    if (capturing(c)) return fail(c, JJ_ERR_INVALID_ARG, "not capturable");
    int32_t rc = ensure(c, &c->tmp2, &c->tmp2_cap, 64);
    if (rc) return rc;
    // best of three register-only probes (kernels.cuh): 64 warps/SM of IMAD.WIDE chains with register / immediate operands,
    // 16 warps/SM of dependent Fq-product chains
    double best_rate = 0;
    for (int mode = 0; mode < 3; mode++) {
        const int iters = mode == 2 ? 4000 : 20000, blocks = c->sm_count * (mode == 2 ? 2 : 8);
        const double per_thread_iter = mode == 2 ? 2.0 * kImadWidePerFqMul : 32.0;
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {  // first pass is warm-up
            CU(c, cudaEventRecord(c->ev0, c->stream));
            if (mode == 0) k_imad_peak<0><<<blocks, 256, 0, c->stream>>>((uint32_t*)c->tmp2, 12345u + rep, iters);
            else if (mode == 1) k_imad_peak<1><<<blocks, 256, 0, c->stream>>>((uint32_t*)c->tmp2, 12345u + rep, iters);
            else k_imad_peak<2><<<blocks, 256, 0, c->stream>>>((uint32_t*)c->tmp2, 12345u + rep, iters);
            c->launches++;
            CU(c, cudaGetLastError());
            CU(c, cudaEventRecord(c->ev1, c->stream));
            CU(c, cudaEventSynchronize(c->ev1));
            float ms = 0;
            CU(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
            if (rep > 0) best = std::min(best, ms);
        }
        best_rate = std::max(best_rate, (double)blocks * 256.0 * iters * per_thread_iter / (best * 1e-3));
    }
    *imad_per_sec = best_rate;
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

int32_t jj_point_neg(jj_ctx* c, const void* p, void* out, size_t n, uint32_t flags) {
    if (!c) return JJ_ERR_INVALID_ARG;
    if (flags & JJ_CANON) return fail(c, JJ_ERR_INVALID_ARG, "point entry points take Montgomery-form coordinates");
    In ins[3] = {{p, 160}, {nullptr, 0}, {nullptr, 0}};
    Out outs[2] = {{out, 160}, {nullptr, 0}};
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging*) -> int32_t {
        k_point_neg<<<grid_for(c, cnt, 256, 8), 256, 0, s>>>(din[0], dout[0], cnt);
        c->launches++;
        CU(c, cudaGetLastError());
        return JJ_OK;
    });
}
int32_t jj_point_eq(jj_ctx* c, const void* p, const void* q, uint8_t* flags_out, size_t n, uint32_t flags) {
    if (!c) return JJ_ERR_INVALID_ARG;
    if (flags & JJ_CANON) return fail(c, JJ_ERR_INVALID_ARG, "point entry points take Montgomery-form coordinates");
    In ins[3] = {{p, 160}, {q, 160}, {nullptr, 0}};
    Out outs[2] = {{flags_out, 1}, {nullptr, 0}};
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging*) -> int32_t {
        k_point_eq<<<grid_for(c, cnt, 128, 8), 128, 0, s>>>(din[0], din[1], (uint8_t*)dout[0], cnt);
        c->launches++;
        CU(c, cudaGetLastError());
        return JJ_OK;
    });
}
int32_t jj_affine_to_extended(jj_ctx* c, const void* p, void* out, size_t n, uint32_t flags) {
    if (!c) return JJ_ERR_INVALID_ARG;
    if (flags & JJ_CANON) return fail(c, JJ_ERR_INVALID_ARG, "point entry points take Montgomery-form coordinates");
    In ins[3] = {{p, 64}, {nullptr, 0}, {nullptr, 0}};
    Out outs[2] = {{out, 160}, {nullptr, 0}};
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging*) -> int32_t {
        k_affine_to_extended<<<grid_for(c, cnt, 256, 8), 256, 0, s>>>(din[0], dout[0], cnt);
        c->launches++;
        CU(c, cudaGetLastError());
        return JJ_OK;
    });
}

int32_t jj_mul_by_cofactor(jj_ctx* c, const void* p, void* out, size_t n, uint32_t flags) {
    if (!c) return JJ_ERR_INVALID_ARG;
    if (flags & JJ_CANON) return fail(c, JJ_ERR_INVALID_ARG, "point entry points take Montgomery-form coordinates");
    In ins[3] = {{p, 160}, {nullptr, 0}, {nullptr, 0}};
    Out outs[2] = {{out, 160}, {nullptr, 0}};
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging*) -> int32_t {
        k_mul_by_cofactor<<<grid_for(c, cnt, 128, 4), 128, 0, s>>>(din[0], dout[0], cnt);
        c->launches++;
        CU(c, cudaGetLastError());
        return JJ_OK;
    });
}

int32_t jj_point_sum(jj_ctx* c, const void* points, void* out, size_t groups, size_t group_size, uint32_t flags) {
    if (!c) return JJ_ERR_INVALID_ARG;
    if (flags & JJ_CANON) return fail(c, JJ_ERR_INVALID_ARG, "point entry points take Montgomery-form coordinates");
    if (groups && (!points || !out)) return fail(c, JJ_ERR_INVALID_ARG, "null pointer");
    CU(c, cudaSetDevice(c->device));
    if (capturing(c)) return fail(c, JJ_ERR_INVALID_ARG, "jj_point_sum cannot be captured into a graph");
    if (groups == 0) return JJ_OK;
    const bool dev = flags & JJ_DEVICE_PTRS;
    const size_t unit = out_unit(flags);
    cudaStream_t s = c->stream;
    if (group_size == 0) {  // empty sums are the identity (the fold's initial value)
        std::vector<uint64_t> id(groups * 20, 0);
        const uint64_t one[4] = {0x00000001fffffffeull, 0x5884b7fa00034802ull, 0x998c4fefecbc4ff5ull, 0x1824b159acc5056full};
        for (size_t gidx = 0; gidx < groups; gidx++)
            for (int w = 0; w < 4; w++) id[gidx * 20 + 4 + w] = id[gidx * 20 + 8 + w] = one[w];
        if (unit != 160) return fail(c, JJ_ERR_INVALID_ARG, "empty groups are supported for ExtendedPoint output only");
        if (dev) CU(c, cudaMemcpy(out, id.data(), groups * 160, cudaMemcpyHostToDevice));
        else memcpy(out, id.data(), groups * 160);
        return JJ_OK;
    }
    constexpr int F = 16;
    const size_t n = groups * group_size;
    // device-resident input (uploaded in one piece for host callers: 160 B per point, the kernels are HBM-bound)
    const char* cur = (const char*)points;
    if (!dev) {
        int32_t rc = ensure(c, &c->tmp, &c->tmp_cap, n * 160);
        if (rc) return rc;
        CU(c, cudaMemcpyAsync(c->tmp, points, n * 160, cudaMemcpyHostToDevice, s));
        cur = c->tmp;
    } else if ((uintptr_t)points & 31 || (uintptr_t)out & 31) {
        return fail(c, JJ_ERR_INVALID_ARG, "device pointer not 32-byte aligned");
    }
    // ping-pong partial sums in the normalise scratch
    const size_t per1 = (group_size + F - 1) / F;
    int32_t rc = ensure(c, &c->norm, &c->norm_cap, 2 * groups * per1 * 160 + 64);
    if (rc) return rc;
    char* buf[2] = {c->norm, c->norm + ((groups * per1 * 160 + 31) & ~(size_t)31)};
    size_t g = group_size;
    int which = 0;
    while (g > 1) {
        const size_t per = (g + F - 1) / F;
        k_point_sum_pass<F><<<grid_for(c, groups * per, 128, 4), 128, 0, s>>>(cur, buf[which], groups, g);
        c->launches++;
        CU(c, cudaGetLastError());
        cur = buf[which];
        which ^= 1;
        g = per;
    }
    // cur: one ExtendedPoint per group -> requested output format
    char* dst = dev ? (char*)out : nullptr;
    if (!dev) {
        rc = ensure(c, &c->tmp2, &c->tmp2_cap, groups * 160 + groups * 32);
        if (rc) return rc;
        dst = c->tmp2;
    }
    if (unit == 160) {
        CU(c, cudaMemcpyAsync(dst, cur, groups * 160, cudaMemcpyDeviceToDevice, s));
    } else {
        rc = normalize_launch(c, s, cur, dst, groups, (int)unit, &c->tbl, &c->tbl_cap, true);
        if (rc) return rc;
    }
    if (!dev) CU(c, cudaMemcpyAsync(out, dst, groups * unit, cudaMemcpyDeviceToHost, s));
    if (!dev || !(flags & JJ_ASYNC)) CU(c, cudaStreamSynchronize(s));
    return JJ_OK;
}

int32_t jj_scalar_mul(jj_ctx* c, const void* points, const void* scalars, void* out, size_t n, uint32_t flags) {
    if (!c) return JJ_ERR_INVALID_ARG;
    if (flags & JJ_CANON) return fail(c, JJ_ERR_INVALID_ARG, "point entry points take Montgomery-form coordinates");
    In ins[3] = {{points, 160}, {scalars, 32}, {nullptr, 0}};
    Out outs[2] = {{out, out_unit(flags)}, {nullptr, 0}};
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging* S) -> int32_t {
        return smul_any(c, s, S, din[0], false, din[1], dout[0], cnt, flags);
    }, chunked(smul_chunk(c)));
}

int32_t jj_scalar_mul_encoded(jj_ctx* c, const void* points32, const void* scalars, void* out, uint8_t* ok, size_t n,
                              uint32_t flags) {
    if (!c) return JJ_ERR_INVALID_ARG;
    if (flags & JJ_CANON) return fail(c, JJ_ERR_INVALID_ARG, "point entry points take Montgomery-form coordinates");
    if ((flags & JJ_CHECK_SUBGROUP) && !ok && n) return fail(c, JJ_ERR_INVALID_ARG, "JJ_CHECK_SUBGROUP needs the ok[] array");
    In ins[3] = {{points32, 32}, {scalars, 32}, {nullptr, 0}};
    Out outs[2] = {{out, out_unit(flags)}, {ok, 1}};
    const bool zip216 = !(flags & JJ_PRE_ZIP216), subgroup = flags & JJ_CHECK_SUBGROUP;
    // Host buffers: when they are pinned, the decode kernel reads the encodings in place and the normalise + encode pass
    // stores the 32-byte results in place (both coalesced, both compute-bound kernels), and the scalars -- which only the
    // scalar-mul kernel reads -- are uploaded while the chunk is being decoded: no upload in front of the first kernel, no
    // download behind the last one.  Pageable buffers are staged as everywhere else.
    BatchOpts opts = chunked(smul_chunk(c, wire_rounds()));
    opts.direct_in0 = wire_direct();
    opts.direct_out0 = wire_direct() && out_unit(flags) == 32;
    opts.late_in = wire_direct() ? 1 : -1;
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging* S) -> int32_t {
        // decode -> AffinePoint scratch (64 B) [-> subgroup test on it] -> scalar-mul reading the affine points directly
        char** aff = S ? &S->tmp2 : &c->tmp2;
        size_t* affcap = S ? &S->tmp2_cap : &c->tmp2_cap;
        int32_t rc = ensure(c, aff, affcap, cnt * 64, !S);
        if (rc) return rc;
        rc = from_bytes_launch(c, s, din[0], *aff, (uint8_t*)dout[1], cnt, zip216);
        if (rc) return rc;
        if (subgroup) {  // SubgroupPoint::from_bytes (src/lib.rs:1427-1429): decoded AND torsion free
            k_is_torsion_free<64, false, true><<<grid_for(c, cnt, 128, 4), 128, 0, s>>>(*aff, (uint8_t*)dout[1], cnt);
            c->launches++;
            CU(c, cudaGetLastError());
        }
        if (S && opts.late_in == 1) CU(c, cudaStreamWaitEvent(s, S->late, 0));  // the scalars have arrived
        return smul_any(c, s, S, *aff, true, din[1], dout[0], cnt, flags);
    }, opts);
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
    const bool smont = flags & JJ_SCALAR_MONT;
    const int unit = (int)out_unit(flags);
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging* S) -> int32_t {
        char* dst = dout[0];
        char** tmp = S ? &S->buf[2] : &c->tmp;
        size_t* tmpcap = S ? &S->cap[2] : &c->tmp_cap;
        if (unit != 160) {
            int32_t rc = ensure(c, tmp, tmpcap, cnt * 160, !S);
            if (rc) return rc;
            dst = *tmp;
        }
        const int w = fixed_w(c);
        int32_t rc = w == 4   ? launch_fixed<256, 4, false>(c, s, din[0], dst, cnt, smont)
                     : w == 7 ? launch_fixed<512, 7, true>(c, s, din[0], dst, cnt, smont)
                     : w == 16 ? launch_fixed_wide<16>(c, s, din[0], dst, cnt, smont)
                               : launch_fixed_wide<kFixedWide>(c, s, din[0], dst, cnt, smont);
        if (rc || unit == 160) return rc;
        return normalize_launch(c, s, dst, dout[0], cnt, unit, S ? &S->tmp2 : &c->tmp2, S ? &S->tmp2_cap : &c->tmp2_cap, !S);
    });
}

int32_t jj_batch_normalize(jj_ctx* c, const void* in, void* out, size_t n, uint32_t flags) {
    if (!c) return JJ_ERR_INVALID_ARG;
    if (in == out && n) return fail(c, JJ_ERR_INVALID_ARG, "batch_normalize output must not alias its input");
    // JJ_OUT_BYTES: normalise and encode in one pass -- GroupEncoding::to_bytes for ExtendedPoint (src/lib.rs:1419-1421)
    const int unit = (flags & JJ_OUT_BYTES) ? 32 : 64;
    In ins[3] = {{in, 160}, {nullptr, 0}, {nullptr, 0}};
    Out outs[2] = {{out, (size_t)unit}, {nullptr, 0}};
    BatchOpts opts;
    opts.direct_out0 = unit == 32;  // 32-byte results are stored into a page-locked caller buffer in place
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging* S) {
        return normalize_launch(c, s, din[0], dout[0], cnt, unit, S ? &S->tmp2 : &c->tmp2, S ? &S->tmp2_cap : &c->tmp2_cap, !S);
    }, opts);
}
int32_t jj_batch_normalize_extended(jj_ctx* c, const void* in, void* out, size_t n, uint32_t flags) {
    if (!c) return JJ_ERR_INVALID_ARG;
    In ins[3] = {{in, 160}, {nullptr, 0}, {nullptr, 0}};
    Out outs[2] = {{out, 160}, {nullptr, 0}};
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging* S) {
        return normalize_launch(c, s, din[0], dout[0], cnt, 160, S ? &S->tmp2 : &c->tmp2, S ? &S->tmp2_cap : &c->tmp2_cap, !S);
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
}  // extern "C"

// [r]P == identity by the reference's own route (src/lib.rs:709-711): the batch shares r, so its width-5 NAF (42
// additions instead of the 58 of the per-unit signed radix-16 windows) drives one scalar-mul kernel.  JJ_TORSION_LADDER.
static int32_t torsion_ladder(jj_ctx* c, const void* p, uint8_t* flags_out, size_t n, uint32_t flags) {
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
        int32_t rc = ensure(c, tbl, tcap, (size_t)grid * (T / 32) * 32768, !S);
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
// Subgroup tests by the order-8 Tate pairing (torsion.cuh): one 223-bit power instead of a scalar multiplication.
template <bool PRIME_ORDER>
static int32_t torsion_pairing(jj_ctx* c, const void* p, uint8_t* flags_out, size_t n, uint32_t flags) {
    In ins[3] = {{p, 160}, {nullptr, 0}, {nullptr, 0}};
    Out outs[2] = {{flags_out, 1}, {nullptr, 0}};
    return run_batch(c, flags, n, ins, outs, [=](cudaStream_t s, const char** din, char** dout, size_t cnt, Staging*) -> int32_t {
        k_is_torsion_free<160, PRIME_ORDER, false><<<grid_for(c, cnt, 128, 4), 128, 0, s>>>(din[0], (uint8_t*)dout[0], cnt);
        c->launches++;
        CU(c, cudaGetLastError());
        return JJ_OK;
    });
}
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

extern "C" {
int32_t jj_is_torsion_free(jj_ctx* c, const void* p, uint8_t* flags_out, size_t n, uint32_t flags) {
    if (!c) return JJ_ERR_INVALID_ARG;
    if (!flags_out && n) return fail(c, JJ_ERR_INVALID_ARG, "null output pointer");
    if (flags & JJ_TORSION_LADDER) return torsion_ladder(c, p, flags_out, n, flags);
    return torsion_pairing<false>(c, p, flags_out, n, flags);
}
int32_t jj_is_prime_order(jj_ctx* c, const void* p, uint8_t* flags_out, size_t n, uint32_t flags) {
    if (!c) return JJ_ERR_INVALID_ARG;
    if (!flags_out && n) return fail(c, JJ_ERR_INVALID_ARG, "null output pointer");
    return torsion_pairing<true>(c, p, flags_out, n, flags);
}
int32_t jj_is_identity(jj_ctx* c, const void* p, uint8_t* f, size_t n, uint32_t flags) { return point_flag<0>(c, p, f, n, flags); }
int32_t jj_is_small_order(jj_ctx* c, const void* p, uint8_t* f, size_t n, uint32_t flags) { return point_flag<1>(c, p, f, n, flags); }

// ---- multi-GPU -----------------------------------------------------------------------------------------
int32_t jj_comm_unique_id(void* id128) {
    if (!id128) return JJ_ERR_INVALID_ARG;
    if (!load_nccl()) return JJ_ERR_NCCL;
    return g_nccl.GetUniqueId(id128) == 0 ? JJ_OK : JJ_ERR_NCCL;
}
int32_t jj_comm_init(jj_ctx* c, int32_t nranks, int32_t rank, const void* id128) {
    if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return JJ_ERR_INVALID_ARG;
    if (!load_nccl()) return fail(c, JJ_ERR_NCCL, "libnccl.so.2 not found (or it lacks an entry point this library calls)");
    CU(c, cudaSetDevice(c->device));
    Id128 id;
    memcpy(id.b, id128, 128);
    int rc = g_nccl.CommInitRank(&c->nccl_comm, nranks, id, rank);
    if (rc != 0) return nccl_fail(c, "ncclCommInitRank", rc);
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
    c->peers.n_peers = 0;
    return JJ_OK;
}
}  // extern "C"

// Block partition of SURVEY.md section 8e: rank g of G owns [g*n/G, (g+1)*n/G) -- blocks differ by at most one unit.
static size_t shard_begin(size_t n_total, int rank, int nranks) { return (size_t)((unsigned __int128)n_total * rank / nranks); }

// All ranks' blocks of out_all <- every rank's own block (already in place).  Equal blocks: one in-place
// ncclAllGather; ragged blocks: one ncclBroadcast per rank inside a group (the usual all-gather-v).
static int32_t gather_blocks(jj_ctx* c, char* out_all, size_t n_total, size_t unit) {
    if (c->nranks == 1) return JJ_OK;
    if (!c->nccl_comm) return fail(c, JJ_ERR_NCCL, "jj_comm_init was not called");
    if (n_total % c->nranks == 0) {
        const size_t blk = n_total / c->nranks * unit;
        int rc = g_nccl.AllGather(out_all + (size_t)c->rank * blk, out_all, blk, /*ncclInt8*/ 0, c->nccl_comm, c->stream);
        return rc == 0 ? JJ_OK : nccl_fail(c, "ncclAllGather", rc);
    }
    int rc = g_nccl.GroupStart();
    for (int r = 0; r < c->nranks && rc == 0; r++) {
        const size_t lo = shard_begin(n_total, r, c->nranks), hi = shard_begin(n_total, r + 1, c->nranks);
        char* blk = out_all + lo * unit;
        if (hi > lo) rc = g_nccl.Broadcast(blk, blk, (hi - lo) * unit, /*ncclInt8*/ 0, r, c->nccl_comm, c->stream);
    }
    int rc2 = g_nccl.GroupEnd();
    if (rc != 0 || rc2 != 0) return nccl_fail(c, "ncclBroadcast group", rc ? rc : rc2);
    return JJ_OK;
}

extern "C" {
int32_t jj_scalar_mul_sharded_n(jj_ctx* c, const void* points_local, const void* scalars_local, void* out_all,
                                void* out_local_host, size_t n_total, uint32_t flags) {
    if (!c || !out_all) return JJ_ERR_INVALID_ARG;
    if (flags & JJ_CANON) return fail(c, JJ_ERR_INVALID_ARG, "point entry points take Montgomery-form coordinates");
    CU(c, cudaSetDevice(c->device));
    if (capturing(c)) return fail(c, JJ_ERR_INVALID_ARG, "jj_scalar_mul_sharded cannot be captured into a graph");
    const size_t unit = out_unit(flags);
    const size_t lo = shard_begin(n_total, c->rank, c->nranks), n_local = shard_begin(n_total, c->rank + 1, c->nranks) - lo;
    const bool host_in = !(flags & JJ_DEVICE_PTRS);
    if (n_local && (!points_local || !scalars_local)) return fail(c, JJ_ERR_INVALID_ARG, "null input pointer");
    if ((uintptr_t)out_all & 31) return fail(c, JJ_ERR_INVALID_ARG, "device pointer not 32-byte aligned");
    if (!host_in && (((uintptr_t)points_local | (uintptr_t)scalars_local) & 31))
        return fail(c, JJ_ERR_INVALID_ARG, "device pointer not 32-byte aligned");
    if (c->nranks > 1 && !c->nccl_comm) return fail(c, JJ_ERR_NCCL, "jj_comm_init was not called");
    // (host-staged chunks with converted outputs go through the ExtendedPoint scratch + normalise pass of one rank:
    // they use the NCCL gather)
    const bool fused = c->peers.n_peers == c->nranks && c->nranks > 1 && !(host_in && unit != 160);
    if (fused && c->peers.ptr[c->rank] != (char*)out_all) return fail(c, JJ_ERR_INVALID_ARG, "out_all is not the registered peer buffer");
    PeerOut po = c->peers;
    po.base_unit = lo;
    if (fused) {
        // Leading rendezvous: a peer may still be reading the previous contents of ITS copy of out_all (work it
        // ordered on its context's stream, or finished on the host, before entering this call); nobody stores into
        // anybody's buffer before every rank has arrived here.
        int32_t rc = rendezvous(c);
        if (rc) return rc;
    }
    char* mine = (char*)out_all + lo * unit;
    if (!host_in) {
        int32_t rc = JJ_OK;
        if (n_local)
            rc = smul_any(c, c->stream, nullptr, (const char*)points_local, false, (const char*)scalars_local,
                          fused ? nullptr : mine, n_local, flags, fused ? &po : nullptr);
        if (rc) return rc;
        if (out_local_host && n_local)
            CU(c, cudaMemcpyAsync(out_local_host, mine, n_local * unit, cudaMemcpyDeviceToHost, c->stream));
    } else {
        // host inputs: this rank's block is staged in round-sized chunks on the two staging streams (upload of chunk
        // k+1 overlaps the kernel of chunk k); every chunk's kernel writes its results into place -- all ranks' buffers
        // when fused -- and the chunk is read back to the host from this rank's own copy.
        CU(c, cudaEventRecord(c->ev_fork, c->stream));
        static const int shard_rounds = env_rounds("JJ_SHARD_ROUNDS", 1);  // experiments: rounds per staged chunk
        const size_t chunk = smul_chunk(c, shard_rounds);
        size_t done = 0;
        int stage = 0, used = 0;
        while (done < n_local) {
            const size_t cnt = std::min(chunk, n_local - done);
            Staging& S = c->st[stage];
            CU(c, cudaStreamSynchronize(S.stream));
            int32_t rc = ensure(c, &S.buf[0], &S.cap[0], chunk * 160, false);
            if (!rc) rc = ensure(c, &S.buf[1], &S.cap[1], chunk * 32, false);
            if (rc) return rc;
            CU(c, cudaMemcpyAsync(S.buf[0], (const char*)points_local + done * 160, cnt * 160, cudaMemcpyHostToDevice, S.stream));
            CU(c, cudaMemcpyAsync(S.buf[1], (const char*)scalars_local + done * 32, cnt * 32, cudaMemcpyHostToDevice, S.stream));
            if (used < kStages) {
                // the uploads above touch this rank's staging buffers only: they need not wait for the leading rendezvous
                // (a rank that arrives early uploads its first chunks while it waits for the others); the kernels do
                CU(c, cudaStreamWaitEvent(S.stream, c->ev_fork, 0));
                used++;
            }
            PeerOut pc = po;
            pc.base_unit = lo + done;
            rc = smul_any(c, S.stream, &S, S.buf[0], false, S.buf[1], fused ? nullptr : mine + done * unit, cnt, flags,
                          fused ? &pc : nullptr);
            if (rc) return rc;
            if (out_local_host)
                CU(c, cudaMemcpyAsync((char*)out_local_host + done * unit, mine + done * unit, cnt * unit,
                                      cudaMemcpyDeviceToHost, S.stream));
            done += cnt;
            stage = (stage + 1) % kStages;
        }
        for (int k = 0; k < used; k++) {  // join: the main stream continues after every chunk
            CU(c, cudaEventRecord(c->ev_join[k], c->st[k].stream));
            CU(c, cudaStreamWaitEvent(c->stream, c->ev_join[k], 0));
        }
    }
    if (fused) {
        // Trailing rendezvous: once this 4-byte all-reduce completes here, every peer's kernel (enqueued before its
        // own all-reduce) has finished storing into this rank's buffer.
        int32_t rc = rendezvous(c);
        if (rc) return rc;
    } else {
        int32_t rc = gather_blocks(c, (char*)out_all, n_total, unit);
        if (rc) return rc;
    }
    if (!(flags & JJ_ASYNC) || host_in) CU(c, cudaStreamSynchronize(c->stream));
    return JJ_OK;
}
int32_t jj_scalar_mul_sharded(jj_ctx* c, const void* points_local, const void* scalars_local, void* out_all,
                              size_t n_local, uint32_t flags) {
    if (!c) return JJ_ERR_INVALID_ARG;
    if (!(flags & JJ_DEVICE_PTRS)) return fail(c, JJ_ERR_INVALID_ARG, "jj_scalar_mul_sharded takes device pointers");
    return jj_scalar_mul_sharded_n(c, points_local, scalars_local, out_all, nullptr, n_local * (size_t)c->nranks, flags);
}

// Sum over a batch that is spread over the ranks: every rank sums its own points (jj_point_sum), the G partial sums
// (160 B each) are all-gathered and every rank adds them up in rank order -- the one step of this engine where ranks
// really exchange data they need for their result.  out: one ExtendedPoint (or JJ_OUT_AFFINE / JJ_OUT_BYTES), the same on
// every rank.
int32_t jj_point_sum_sharded(jj_ctx* c, const void* points_local, void* out, size_t n_local, uint32_t flags) {
    if (!c || !out) return JJ_ERR_INVALID_ARG;
    if (!(flags & JJ_DEVICE_PTRS)) return fail(c, JJ_ERR_INVALID_ARG, "jj_point_sum_sharded takes device pointers");
    CU(c, cudaSetDevice(c->device));
    if (c->nranks > 1 && !c->nccl_comm) return fail(c, JJ_ERR_NCCL, "jj_comm_init was not called");
    if (!c->barrier_word) {
        CU(c, cudaMalloc((void**)&c->barrier_word, 256));
        CU(c, cudaMemset(c->barrier_word, 0, 256));
    }
    int32_t rc = ensure(c, &c->tmp2, &c->tmp2_cap, (size_t)(c->nranks + 1) * 160 + 64);
    if (rc) return rc;
    char* partials = c->tmp2;  // nranks x 160 B, this rank's at index rank
    rc = jj_point_sum(c, points_local, partials + (size_t)c->rank * 160, 1, n_local, JJ_DEVICE_PTRS | JJ_ASYNC);
    if (rc) return rc;
    if (c->nranks > 1) {
        int nrc = g_nccl.AllGather(partials + (size_t)c->rank * 160, partials, 160, /*ncclInt8*/ 0, c->nccl_comm, c->stream);
        if (nrc != 0) return nccl_fail(c, "ncclAllGather", nrc);
    }
    // the gathered partials are consumed from a second scratch (jj_point_sum uses tmp2 only for host outputs)
    rc = ensure(c, &c->tmp, &c->tmp_cap, (size_t)c->nranks * 160);
    if (rc) return rc;
    CU(c, cudaMemcpyAsync(c->tmp, partials, (size_t)c->nranks * 160, cudaMemcpyDeviceToDevice, c->stream));
    return jj_point_sum(c, c->tmp, out, 1, (size_t)c->nranks, flags);
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
