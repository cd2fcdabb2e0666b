/*
 * jubjub_b200.h -- C ABI of the B200-native batched Jubjub engine (libjubjub_b200.so).
 *
 * The reference (zkcrypto/jubjub 0.10.0) is a leaf Rust library with no FFI of its own; the
 * boundary below is what a Rust shim's `extern "C"` block binds to add batch entry points
 * (`batch_mul`, `batch_add`, `Fq::batch_mul`, ...) behind the unchanged jubjub:: types
 * (INTEGRATION.md shows that shim).  Each entry point names the reference item it replaces
 * (paths relative to /root/reference).  Plain pointers and sizes only; no C++/torch types.
 *
 * Data layout (all little-endian):
 *   field element     32 B = 4 x u64 limbs.  Default: the reference's internal Montgomery
 *                     form, i.e. what `Fr(pub(crate) [u64; 4])` (src/fr.rs:23) / bls12_381::Scalar
 *                     hold.  With JJ_CANON: canonical integers, i.e. `to_bytes()` form
 *                     (src/fr.rs:296-308); inputs >= m are reduced mod m as `from_raw` does (:347-349).
 *   ExtendedPoint     160 B = (u, v, z, t1, t2)            src/lib.rs:139-145
 *   AffinePoint        64 B = (u, v)                       src/lib.rs:81-84
 *   ExtendedNielsPoint 128 B = (v+u, v-u, z, t2d)          src/lib.rs:327-332
 *   AffineNielsPoint   96 B = (v+u, v-u, t2d)              src/lib.rs:255-259
 *   scalar             32 B canonical LE; bits 252..255 are ignored exactly as
 *                      `multiply_bits` does (src/lib.rs:381-385).  With JJ_SCALAR_MONT the 32 B
 *                      are Fr Montgomery limbs and are converted on the device (Fr::to_bytes,
 *                      src/fr.rs:296-308, as `Mul<&Fr>` does at src/lib.rs:877).
 *
 * Inputs must satisfy the reference's type invariants (field elements < m, z != 0).
 * Buffers are caller-owned and never retained.  Pointers are host memory unless
 * JJ_DEVICE_PTRS is set (then: memory of ctx's device, 32-byte aligned).  `out` may alias
 * an input of the same shape (the reference's `*Assign` operators, src/util.rs:126-152).
 * Calls are synchronous unless JJ_ASYNC is set together with JJ_DEVICE_PTRS; then they are
 * ordered on the context's stream and jj_sync() waits for them.
 * A context is not thread-safe; use one per thread or per GPU.
 *
 * Every function returns JJ_OK (0) or a negative error; nothing throws or aborts.
 * Per-element failures that the reference reports as CtOption::none (invert of zero,
 * src/fr.rs:539; non-canonical bytes, src/fr.rs:291; off-curve encodings, src/lib.rs:492-534)
 * are reported in a `uint8_t ok[n]` side array, not as an error.
 *
 * The batch entry points are variable-time in their data (the reference's constant-time policy is src/lib.rs:12-17).
 * The variable-base scalar multiplication has a constant-time-in-the-scalar mode, JJ_CONST_TIME; everything else
 * (fixed-base tables, decoding, inversion chains) is for public data.
 */
#ifndef JUBJUB_B200_H
#define JUBJUB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jj_ctx jj_ctx;

enum {
    JJ_OK = 0,
    JJ_ERR_INVALID_ARG = -1, /* null pointer, misaligned device pointer, bad flag, length mismatch
                                (the reference panics: assert_eq! at src/lib.rs:841) */
    JJ_ERR_CUDA = -2,
    JJ_ERR_NCCL = -3,
    JJ_ERR_OOM = -4,
    JJ_ERR_NO_DEVICE = -5
};

enum {
    JJ_MONT = 0u,          /* field elements are Montgomery limbs (default)                   */
    JJ_CANON = 1u << 0,    /* field elements (in and out) are canonical integers             */
    JJ_DEVICE_PTRS = 1u << 1,
    JJ_ASYNC = 1u << 2,
    JJ_SUBTRACT = 1u << 3, /* point add entry points compute p - q (src/lib.rs:922-940, 970-988, 1001-1008) */
    JJ_SCALAR_MONT = 1u << 4,
    JJ_OUT_AFFINE = 1u << 5, /* scalar-mul writes normalised AffinePoint (64 B) instead of Extended */
    JJ_OUT_BYTES = 1u << 6,  /* scalar-mul writes the 32-byte encoding (src/lib.rs:455-464)      */
    JJ_PRE_ZIP216 = 1u << 7, /* jj_batch_from_bytes: from_bytes_pre_zip216_compatibility (src/lib.rs:485-490) */
    JJ_CHECK_SUBGROUP = 1u << 8, /* jj_scalar_mul_encoded: ok[i] also requires is_torsion_free, i.e. the decode is
                                    SubgroupPoint::from_bytes (src/lib.rs:1427-1429) */
    JJ_TORSION_LADDER = 1u << 9, /* jj_is_torsion_free: decide by the reference's own [r]P == O (src/lib.rs:709-711)
                                    instead of the pairing test -- same flags, ~7x the work; the cross-check */
    JJ_CONST_TIME = 1u << 10     /* jj_scalar_mul / jj_scalar_mul_encoded / jj_scalar_mul_sharded*: no branch and no memory
                                    address depends on the SCALAR (window table scanned, sign by selects, every addition
                                    executed -- the batch analogue of the reference's "always add P or identity",
                                    src/lib.rs:356-379).  Same results, slower kernel; points are treated as public. */
};

/* ---- context ------------------------------------------------------------------------------ */
int32_t jj_init(int device, jj_ctx** out);
int32_t jj_destroy(jj_ctx* ctx);
int32_t jj_sync(jj_ctx* ctx);
const char* jj_last_error(const jj_ctx* ctx);
const char* jj_version(void);
int32_t jj_device_info(jj_ctx* ctx, int32_t* sm_count, int32_t* sm_clock_khz, uint64_t* hbm_bytes);
/* Tuning knob; 0 = library defaults.  13 / 24 / 5: variable-base kernel with 16 / 24 / 8 warps
 * per SM; fixed-base table: default 12-bit windows (4.1 MB in global memory, 22 additions per scalar-mul), 116: 16-bit windows
 * (50 MB, 17 additions: for bases that live long enough to pay the 18 ms table build), 107 / 100: 7- / 4-bit windows in shared
 * memory (216 KB TMA-staged / 47 KB); 200 / 201: converted outputs (JJ_OUT_AFFINE / JJ_OUT_BYTES) of
 * device-resident variable-base batches always / never use the kernel's fused normalise epilogue (default: only for the
 * fused all-gather, where it shrinks the peer stores to 32 bytes; on one GPU the separate pass is 1 % faster). */
int32_t jj_set_scalar_mul_variant(jj_ctx* ctx, int32_t variant);
/* Number of this library's kernel launches issued on ctx so far (bench.py's gpu_launches). */
uint64_t jj_launch_count(const jj_ctx* ctx);

/* Device/pinned memory and a device timer, so hosts without a CUDA binding can keep
 * batches resident and time them on the context's stream with CUDA events. */
int32_t jj_malloc(jj_ctx* ctx, size_t bytes, void** dptr);
int32_t jj_free(jj_ctx* ctx, void* dptr);
int32_t jj_host_alloc(jj_ctx* ctx, size_t bytes, void** hptr); /* pinned */
int32_t jj_host_free(jj_ctx* ctx, void* hptr);
int32_t jj_memcpy_h2d(jj_ctx* ctx, void* dptr, const void* hptr, size_t bytes);
int32_t jj_memcpy_d2h(jj_ctx* ctx, void* hptr, const void* dptr, size_t bytes);
int32_t jj_timer_start(jj_ctx* ctx);
int32_t jj_timer_stop(jj_ctx* ctx, float* elapsed_ms);
int32_t jj_flush_l2(jj_ctx* ctx); /* overwrites a scratch buffer larger than L2 */
/* CUDA graphs for launch-bound sequences (many small field/point batches): between jj_graph_begin and
 * jj_graph_end every call made with JJ_DEVICE_PTRS | JJ_ASYNC is captured on the context's stream instead of
 * executed; jj_graph_launch replays the whole sequence with one launch (ordered on the stream, jj_sync waits).
 * Rules (violations return JJ_ERR_INVALID_ARG, nothing is corrupted):
 *  - scratch buffers must already exist and be large enough: run the sequence once eagerly before capturing it --
 *    a captured call that would have to allocate fails, and so does any other call inside a capture that is not
 *    JJ_DEVICE_PTRS | JJ_ASYNC (host-pointer calls synchronise), jj_scalar_mul_fixed (its table cache is checked on
 *    the host) and jj_scalar_mul_sharded;
 *  - a captured graph holds the scratch pointers of its kernels, so while any graph of this context is alive
 *    (until jj_graph_destroy) a call that would have to GROW a scratch buffer of the main stream fails instead of
 *    reallocating it: run the largest batch first, or destroy the graphs. */
int32_t jj_graph_begin(jj_ctx* ctx);
int32_t jj_graph_end(jj_ctx* ctx, void** graph_exec);
int32_t jj_graph_launch(jj_ctx* ctx, void* graph_exec);
int32_t jj_graph_destroy(jj_ctx* ctx, void* graph_exec);
/* Measures the chip's IMAD.WIDE.U32 issue rate (instructions x 32 lanes per second) with register-only kernels -- the best
 * of three probes (accumulate chains with register operands, with an immediate multiplier, and dependent chains of whole
 * Fq products): the integer-pipe roofline denominator. */
int32_t jj_measure_imad_peak(jj_ctx* ctx, double* imad_per_sec);

/* ---- field batches: out[i] = a[i] (op) b[i]; field = Fq (jj_fq_*) or Fr (jj_fr_*) --------- */
/* Fr::mul src/fr.rs:592-616 (+ montgomery_reduce :544-588); Fq = [ext] bls12_381::Scalar, same shape */
int32_t jj_fq_mul(jj_ctx* ctx, const void* a, const void* b, void* out, size_t n, uint32_t flags);
int32_t jj_fr_mul(jj_ctx* ctx, const void* a, const void* b, void* out, size_t n, uint32_t flags);
/* Fr::square src/fr.rs:353-381 */
int32_t jj_fq_square(jj_ctx* ctx, const void* a, void* out, size_t n, uint32_t flags);
int32_t jj_fr_square(jj_ctx* ctx, const void* a, void* out, size_t n, uint32_t flags);
/* Fr::add src/fr.rs:638-647, Fr::sub :620-634, Fr::neg :651-665, Fr::double :261-263 */
int32_t jj_fq_add(jj_ctx* ctx, const void* a, const void* b, void* out, size_t n, uint32_t flags);
int32_t jj_fr_add(jj_ctx* ctx, const void* a, const void* b, void* out, size_t n, uint32_t flags);
int32_t jj_fq_sub(jj_ctx* ctx, const void* a, const void* b, void* out, size_t n, uint32_t flags);
int32_t jj_fr_sub(jj_ctx* ctx, const void* a, const void* b, void* out, size_t n, uint32_t flags);
int32_t jj_fq_neg(jj_ctx* ctx, const void* a, void* out, size_t n, uint32_t flags);
int32_t jj_fr_neg(jj_ctx* ctx, const void* a, void* out, size_t n, uint32_t flags);
int32_t jj_fq_double(jj_ctx* ctx, const void* a, void* out, size_t n, uint32_t flags);
int32_t jj_fr_double(jj_ctx* ctx, const void* a, void* out, size_t n, uint32_t flags);
/* Fr::invert src/fr.rs:438-540; ok[i] = 0 and out[i] = 0 for a[i] = 0 (CtOption::none) */
int32_t jj_fq_invert(jj_ctx* ctx, const void* a, void* out, uint8_t* ok, size_t n, uint32_t flags);
int32_t jj_fr_invert(jj_ctx* ctx, const void* a, void* out, uint8_t* ok, size_t n, uint32_t flags);
/* Fr::sqrt src/fr.rs:384-399 (a^((r+1)/4)); Fq::sqrt = [ext] bls12_381::Scalar::sqrt (Tonelli-Shanks, call sites
 * src/lib.rs:515, 603).  ok[i] = 0 and out[i] = 0 for a non-residue.  Which of the two roots is returned is
 * not part of the contract (callers fix the sign from the parity, src/lib.rs:518-520): out[i]^2 == a[i]. */
int32_t jj_fq_sqrt(jj_ctx* ctx, const void* a, void* out, uint8_t* ok, size_t n, uint32_t flags);
int32_t jj_fr_sqrt(jj_ctx* ctx, const void* a, void* out, uint8_t* ok, size_t n, uint32_t flags);
/* Fr::to_bytes src/fr.rs:296-308: Montgomery limbs -> 32 canonical LE bytes */
int32_t jj_fq_to_bytes(jj_ctx* ctx, const void* a, void* out32, size_t n, uint32_t flags);
int32_t jj_fr_to_bytes(jj_ctx* ctx, const void* a, void* out32, size_t n, uint32_t flags);
/* Fr::from_bytes src/fr.rs:268-292: ok[i] = 0 when the 32 bytes are >= m */
int32_t jj_fq_from_bytes(jj_ctx* ctx, const void* in32, void* out, uint8_t* ok, size_t n, uint32_t flags);
int32_t jj_fr_from_bytes(jj_ctx* ctx, const void* in32, void* out, uint8_t* ok, size_t n, uint32_t flags);
/* Fr::from_bytes_wide src/fr.rs:312-343: 64 LE bytes -> d0*R2 + d1*R3 */
int32_t jj_fq_from_bytes_wide(jj_ctx* ctx, const void* in64, void* out, size_t n, uint32_t flags);
int32_t jj_fr_from_bytes_wide(jj_ctx* ctx, const void* in64, void* out, size_t n, uint32_t flags);
/* Synthetic inputs: element i = from_bytes_wide(SplitMix64(seed) outputs [8(first+i), 8(first+i)+8)). */
int32_t jj_fq_stream(jj_ctx* ctx, uint64_t seed, size_t first, void* out, size_t n, uint32_t flags);
int32_t jj_fr_stream(jj_ctx* ctx, uint64_t seed, size_t first, void* out, size_t n, uint32_t flags);

/* ---- point batches (ExtendedPoint in, ExtendedPoint out; formulas verbatim => 160 B bit-exact) */
/* ExtendedPoint::double src/lib.rs:739-828 */
int32_t jj_point_double(jj_ctx* ctx, const void* p_ext, void* out_ext, size_t n, uint32_t flags);
/* &ExtendedPoint + &ExtendedPoint src/lib.rs:992-1008 (to_niels :728-735, then the 8M add :883-940) */
int32_t jj_point_add(jj_ctx* ctx, const void* p_ext, const void* q_ext, void* out_ext, size_t n, uint32_t flags);
/* &ExtendedPoint + &ExtendedNielsPoint src/lib.rs:883-940 */
int32_t jj_point_add_niels(jj_ctx* ctx, const void* p_ext, const void* q_niels, void* out_ext, size_t n, uint32_t flags);
/* &ExtendedPoint + &AffineNielsPoint src/lib.rs:944-988 */
int32_t jj_point_add_affine_niels(jj_ctx* ctx, const void* p_ext, const void* q_aniels, void* out_ext, size_t n, uint32_t flags);
/* ExtendedPoint::to_niels src/lib.rs:728-735; AffinePoint::to_niels :652-658 */
int32_t jj_point_to_niels(jj_ctx* ctx, const void* p_ext, void* out_niels, size_t n, uint32_t flags);
int32_t jj_affine_to_niels(jj_ctx* ctx, const void* p_affine, void* out_aniels, size_t n, uint32_t flags);

/* out[i] = [scalars[i]] points[i]: `&ExtendedPoint * &Fr` src/lib.rs:873-879 ->
 * ExtendedPoint::multiply :830-833 -> ExtendedNielsPoint::multiply :356-379.
 * Output: ExtendedPoint (projectively equal to the reference's, affine-identical), or with
 * JJ_OUT_AFFINE / JJ_OUT_BYTES the normalised point / its encoding (bit-exact).
 * Variable-time in the scalars by default (signed radix-16 windows, zero digits skipped); with JJ_CONST_TIME no branch and no
 * memory address depends on the scalars (about 4 % slower): the mode that keeps the reference's policy (src/lib.rs:12-17). */
int32_t jj_scalar_mul(jj_ctx* ctx, const void* points_ext, const void* scalars32, void* out, size_t n, uint32_t flags);
/* Wire format in: points32[i] is the 32-byte encoding of an AffinePoint (src/lib.rs:455-464).  Decodes on the
 * device exactly as jj_batch_from_bytes (ZIP-216 rule unless JJ_PRE_ZIP216), multiplies by scalars32[i] and writes
 * the result in the format of jj_scalar_mul (Extended, JJ_OUT_AFFINE or JJ_OUT_BYTES).  ok[i] = 0 marks a rejected
 * encoding; its output unit is then unspecified.  This is `AffinePoint::from_bytes(..) * scalar` for a whole batch
 * with nothing but public types of the reference crate on the caller's side.
 * Host buffers: pass page-locked memory (jj_host_alloc, cudaHostAlloc, cudaHostRegister; 32-byte aligned) and the
 * decode kernel reads points32 in place, scalars32 is uploaded while the chunk is decoded, and JJ_OUT_BYTES results are
 * stored in place -- no staging copy in front of the first kernel or behind the last (38.4 instead of 39.3 ms per 2^20).
 * Pageable buffers are staged; the results are the same. */
int32_t jj_scalar_mul_encoded(jj_ctx* ctx, const void* points32, const void* scalars32, void* out, uint8_t* ok, size_t n,
                              uint32_t flags);

/* out[i] = [scalars[i]] base: `&AffinePoint * &Fr` src/lib.rs:1109-1115 -> AffineNielsPoint::multiply :271-295,
 * one shared base; the per-window AffineNiels table (no doublings: one mixed addition per window) is built on the device
 * once per base and cached in ctx -- the first call with a new base pays for it (5 ms for the default 12-bit windows). */
int32_t jj_scalar_mul_fixed(jj_ctx* ctx, const void* base_affine, const void* scalars32, void* out, size_t n, uint32_t flags);
/* Sum<ExtendedPoint> src/lib.rs:183-193 (`iter.fold(identity, |acc, p| acc + p)`; SubgroupPoint :1161-1171), batched: the
 * input holds `groups` consecutive groups of `group_size` points, out[j] = sum of group j (groups = 1: one sum of the
 * whole batch; group_size = 0: identities).  Together with jj_scalar_mul this is sum_i [k_i] P_i on the device.  The
 * association order is a tree, not the reference's left fold: the result is the same point in other projective
 * coordinates -- JJ_OUT_AFFINE / JJ_OUT_BYTES outputs are bit-exact.  Host or device pointers; not capturable. */
int32_t jj_point_sum(jj_ctx* ctx, const void* points_ext, void* out, size_t groups, size_t group_size, uint32_t flags);
/* Neg for ExtendedPoint src/lib.rs:195-210: (-U, V, Z, -T1, T2), all 160 B bit-exact */
int32_t jj_point_neg(jj_ctx* ctx, const void* p_ext, void* out_ext, size_t n, uint32_t flags);
/* ConstantTimeEq / PartialEq for ExtendedPoint src/lib.rs:153-163, 177-181: flags_out[i] = (u z' == u' z) & (v z' == v' z) */
int32_t jj_point_eq(jj_ctx* ctx, const void* p_ext, const void* q_ext, uint8_t* flags_out, size_t n, uint32_t flags);
/* From<AffinePoint> for ExtendedPoint src/lib.rs:214-226: (u, v) -> (u, v, 1, u, v) */
int32_t jj_affine_to_extended(jj_ctx* ctx, const void* p_affine, void* out_ext, size_t n, uint32_t flags);
/* ExtendedPoint::mul_by_cofactor src/lib.rs:722-724 (= double().double().double(), all 160 B bit-exact) */
int32_t jj_mul_by_cofactor(jj_ctx* ctx, const void* p_ext, void* out_ext, size_t n, uint32_t flags);
/* ExtendedPoint::batch_normalize src/lib.rs:840-858: ExtendedPoint -> AffinePoint (z = 0 gives (0, 0)
 * like ff::BatchInverter's skipped zeros).  With JJ_OUT_BYTES the same pass also encodes: out is n x 32 B,
 * GroupEncoding::to_bytes for ExtendedPoint (`AffinePoint::from(self).to_bytes()`, src/lib.rs:1419-1421). */
int32_t jj_batch_normalize(jj_ctx* ctx, const void* in_ext, void* out_affine, size_t n, uint32_t flags);
/* The free function batch_normalize src/lib.rs:1084-1107: normalises the ExtendedPoints themselves,
 * (u, v, z, t1, t2) -> (u/z, v/z, 1, u/z, v/z) ((0, 0, 1, 0, 0) for z = 0); out_ext may be in_ext (in place, as the
 * reference's `&mut [ExtendedPoint]`); the affine points are the first 64 bytes of every output unit. */
int32_t jj_batch_normalize_extended(jj_ctx* ctx, const void* in_ext, void* out_ext, size_t n, uint32_t flags);
/* AffinePoint::to_bytes src/lib.rs:455-464 */
int32_t jj_affine_to_bytes(jj_ctx* ctx, const void* in_affine, void* out32, size_t n, uint32_t flags);
/* AffinePoint::batch_from_bytes src/lib.rs:541-627 (per element: from_bytes_inner :492-534): 32-byte
 * encodings -> AffinePoint; ok[i] = 0 and (0, 0) when v is non-canonical, off the curve, or (ZIP 216) a
 * non-canonical encoding of (0, +-1).  JJ_PRE_ZIP216 accepts the latter like
 * from_bytes_pre_zip216_compatibility (:485-490). */
int32_t jj_batch_from_bytes(jj_ctx* ctx, const void* in32, void* out_affine, uint8_t* ok, size_t n, uint32_t flags);
/* ExtendedPoint::is_torsion_free src/lib.rs:709-711 ([r]P == identity; decided by the order-8 Tate pairing with the
 * 8-torsion point, one 223-bit power instead of a scalar multiplication -- same booleans, csrc/torsion.cuh),
 * is_prime_order :717-719 (torsion free and not the identity), is_identity :691-696, is_small_order :699-705;
 * flags_out[i] in {0, 1} */
int32_t jj_is_torsion_free(jj_ctx* ctx, const void* p_ext, uint8_t* flags_out, size_t n, uint32_t flags);
int32_t jj_is_prime_order(jj_ctx* ctx, const void* p_ext, uint8_t* flags_out, size_t n, uint32_t flags);
int32_t jj_is_identity(jj_ctx* ctx, const void* p_ext, uint8_t* flags_out, size_t n, uint32_t flags);
int32_t jj_is_small_order(jj_ctx* ctx, const void* p_ext, uint8_t* flags_out, size_t n, uint32_t flags);

/* ---- multi-GPU: one context (one process) per GPU, contiguous block partition of the batch --- */
/* rank g of G owns units [g*n/G, (g+1)*n/G).  jj_comm_unique_id fills a 128-byte NCCL id on one
 * rank; the host distributes it (torch.distributed / MPI / files) and every rank calls jj_comm_init. */
int32_t jj_comm_unique_id(void* id128);
int32_t jj_comm_init(jj_ctx* ctx, int32_t nranks, int32_t rank, const void* id128);
int32_t jj_comm_destroy(jj_ctx* ctx);
/* Computes this rank's shard of out = [scalars] points (shard-local device inputs of n_local units)
 * and all-gathers the results: out_all (device, nranks * n_local units) holds every rank's outputs
 * in rank order.  Output unit = ExtendedPoint, or per JJ_OUT_AFFINE / JJ_OUT_BYTES.  Equal n_local on all ranks. */
int32_t jj_scalar_mul_sharded(jj_ctx* ctx, const void* points_ext_local, const void* scalars32_local,
                              void* out_all, size_t n_local, uint32_t flags);
/* The same for a batch of n_total units that need not divide evenly: rank g of G owns the block
 * [g*n_total/G, (g+1)*n_total/G) (blocks differ by at most one unit; ragged blocks are gathered by one ncclBroadcast
 * per rank in a group).  points/scalars hold this rank's block only.  Without JJ_DEVICE_PTRS the two INPUTS are host
 * buffers, staged in round-sized chunks that overlap the kernels; out_all is always the device-resident gathered
 * buffer.  out_local_host (may be NULL): this rank's own block of results is also copied to that host buffer. */
int32_t jj_scalar_mul_sharded_n(jj_ctx* ctx, const void* points_ext_local, const void* scalars32_local,
                                void* out_all, void* out_local_host, size_t n_total, uint32_t flags);

/* Sum<ExtendedPoint> over a batch spread across the ranks: local jj_point_sum, ncclAllGather of the nranks partial sums
 * (160 B each), sum of those in rank order; `out` (one point, format per JJ_OUT_*) is the same on every rank.  Device
 * pointers.  An empty local block contributes the identity. */
int32_t jj_point_sum_sharded(jj_ctx* ctx, const void* points_ext_local, void* out, size_t n_local, uint32_t flags);

/* Fused compute + all-gather over NVLink peer memory.  Each rank exports its gathered-output buffer
 * (jj_ipc_export -> 64-byte cudaIpcMemHandle), the host exchanges the handles, every rank opens its
 * peers' buffers (jj_ipc_open) and registers the nranks pointers in rank order (own pointer at index
 * rank).  jj_scalar_mul_sharded then stores every result -- ExtendedPoint, or with JJ_OUT_AFFINE / JJ_OUT_BYTES the
 * normalised point / its 32-byte encoding produced by the kernel's fused normalise epilogue -- straight into all
 * ranks' buffers from the scalar-mul kernel (P2P stores), bracketed by two 4-byte stream-ordered NCCL rendezvous
 * instead of a separate ncclAllGather.  Ordering contract: the call may overwrite EVERY rank's registered buffer, and
 * it starts doing so only after all ranks have entered it (leading rendezvous).  So a rank must be done reading its own
 * copy of out_all (work ordered before the call on the context's stream, or finished on the host) before IT calls
 * again; it need not know what its peers are doing.  When the call has completed on a rank (stream order / jj_sync)
 * all peers' stores into that rank's buffer have landed (trailing rendezvous). */
int32_t jj_ipc_export(jj_ctx* ctx, const void* dptr, void* handle64);
int32_t jj_ipc_open(jj_ctx* ctx, const void* handle64, void** dptr);
int32_t jj_ipc_close(jj_ctx* ctx, void* dptr);
int32_t jj_comm_set_peer_outputs(jj_ctx* ctx, void* const* peer_out_all, int32_t count);

#ifdef __cplusplus
}
#endif
#endif
