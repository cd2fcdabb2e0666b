//! UNCOMPILED reference text: this image has no Rust toolchain, so nothing here has been through `rustc`.
//! It is the binding a maintainer adds next to `zkcrypto/jubjub`; the same C ABI is exercised
//! in this repository by the Python binding (`jubjub_b200/engine.py`) and the C++ mirror (`include/jubjub_b200.hpp`).
//!
//! `jubjub-b200`: batched Jubjub operations on an NVIDIA B200 behind the reference crate's own types.
//!
//! The reference keeps `Fq`'s limbs private (`bls12_381::Scalar`, `src/lib.rs:62`) and forbids `unsafe`
//! (`src/lib.rs:24`), so an out-of-crate shim cannot hand Montgomery limbs to the device.  It does not need to:
//! everything below moves **wire formats** obtained through public API only —
//!
//! * points travel as their 32-byte encodings (`AffinePoint::to_bytes`, `src/lib.rs:455-464`); the device decodes them
//!   (`jj_scalar_mul_encoded` = `AffinePoint::batch_from_bytes`, `src/lib.rs:541-627`, then the scalar-mul kernel) and
//!   encodes the results (`JJ_OUT_BYTES`); the host turns them back into points with
//!   `AffinePoint::batch_from_bytes`, or keeps them as the wire format they already are;
//! * scalars travel as `Fr::to_bytes` (`src/fr.rs:296-308`), exactly what `Mul<&Fr>` feeds to `multiply`
//!   (`src/lib.rs:877`);
//! * base-field elements travel as canonical bytes with `JJ_CANON` (`Fq::to_bytes` / `Fq::from_bytes`).
//!
//! All batch entry points are variable-time in their data and are named `*_vartime`, following the reference's rule
//! for non-constant-time functions (`src/lib.rs:12-17`).
#![allow(clippy::missing_safety_doc)]

use core::ffi::{c_char, c_int, c_void};
use std::ffi::CStr;

use group::GroupEncoding;
use jubjub::{AffinePoint, ExtendedPoint, Fq, Fr, SubgroupPoint};
use subtle::CtOption;

// ---------------------------------------------------------------------------------------------------------
// Raw ABI (include/jubjub_b200.h).  One line per exported symbol; `tests/test_host_logic.py` checks that the
// shared library exports every symbol the header declares.
// ---------------------------------------------------------------------------------------------------------

#[repr(C)]
pub struct JjCtx {
    _private: [u8; 0],
}

pub const JJ_OK: i32 = 0;
pub const JJ_ERR_INVALID_ARG: i32 = -1;
pub const JJ_ERR_CUDA: i32 = -2;
pub const JJ_ERR_NCCL: i32 = -3;
pub const JJ_ERR_OOM: i32 = -4;
pub const JJ_ERR_NO_DEVICE: i32 = -5;

pub const JJ_MONT: u32 = 0;
pub const JJ_CANON: u32 = 1 << 0;
pub const JJ_DEVICE_PTRS: u32 = 1 << 1;
pub const JJ_ASYNC: u32 = 1 << 2;
pub const JJ_SUBTRACT: u32 = 1 << 3;
pub const JJ_SCALAR_MONT: u32 = 1 << 4;
pub const JJ_OUT_AFFINE: u32 = 1 << 5;
pub const JJ_OUT_BYTES: u32 = 1 << 6;
pub const JJ_PRE_ZIP216: u32 = 1 << 7;
pub const JJ_CHECK_SUBGROUP: u32 = 1 << 8;
pub const JJ_TORSION_LADDER: u32 = 1 << 9;
pub const JJ_CONST_TIME: u32 = 1 << 10;

#[link(name = "jubjub_b200")]
extern "C" {
    pub fn jj_init(device: c_int, out: *mut *mut JjCtx) -> i32;
    pub fn jj_destroy(ctx: *mut JjCtx) -> i32;
    pub fn jj_sync(ctx: *mut JjCtx) -> i32;
    pub fn jj_last_error(ctx: *const JjCtx) -> *const c_char;
    pub fn jj_version() -> *const c_char;
    pub fn jj_device_info(ctx: *mut JjCtx, sm_count: *mut i32, sm_clock_khz: *mut i32, hbm_bytes: *mut u64) -> i32;
    pub fn jj_set_scalar_mul_variant(ctx: *mut JjCtx, variant: i32) -> i32;

    pub fn jj_malloc(ctx: *mut JjCtx, bytes: usize, dptr: *mut *mut c_void) -> i32;
    pub fn jj_free(ctx: *mut JjCtx, dptr: *mut c_void) -> i32;
    pub fn jj_host_alloc(ctx: *mut JjCtx, bytes: usize, hptr: *mut *mut c_void) -> i32;
    pub fn jj_host_free(ctx: *mut JjCtx, hptr: *mut c_void) -> i32;
    pub fn jj_memcpy_h2d(ctx: *mut JjCtx, dptr: *mut c_void, hptr: *const c_void, bytes: usize) -> i32;
    pub fn jj_memcpy_d2h(ctx: *mut JjCtx, hptr: *mut c_void, dptr: *const c_void, bytes: usize) -> i32;
    pub fn jj_timer_start(ctx: *mut JjCtx) -> i32;
    pub fn jj_timer_stop(ctx: *mut JjCtx, elapsed_ms: *mut f32) -> i32;
    pub fn jj_flush_l2(ctx: *mut JjCtx) -> i32;
    pub fn jj_graph_begin(ctx: *mut JjCtx) -> i32;
    pub fn jj_graph_end(ctx: *mut JjCtx, graph_exec: *mut *mut c_void) -> i32;
    pub fn jj_graph_launch(ctx: *mut JjCtx, graph_exec: *mut c_void) -> i32;
    pub fn jj_graph_destroy(ctx: *mut JjCtx, graph_exec: *mut c_void) -> i32;
    pub fn jj_measure_imad_peak(ctx: *mut JjCtx, imad_per_sec: *mut f64) -> i32;

    // field batches: a, b, out are n x 32 bytes (Montgomery limbs, or canonical integers with JJ_CANON)
    pub fn jj_fq_mul(ctx: *mut JjCtx, a: *const c_void, b: *const c_void, out: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_fr_mul(ctx: *mut JjCtx, a: *const c_void, b: *const c_void, out: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_fq_square(ctx: *mut JjCtx, a: *const c_void, out: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_fr_square(ctx: *mut JjCtx, a: *const c_void, out: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_fq_add(ctx: *mut JjCtx, a: *const c_void, b: *const c_void, out: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_fr_add(ctx: *mut JjCtx, a: *const c_void, b: *const c_void, out: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_fq_sub(ctx: *mut JjCtx, a: *const c_void, b: *const c_void, out: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_fr_sub(ctx: *mut JjCtx, a: *const c_void, b: *const c_void, out: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_fq_neg(ctx: *mut JjCtx, a: *const c_void, out: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_fr_neg(ctx: *mut JjCtx, a: *const c_void, out: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_fq_double(ctx: *mut JjCtx, a: *const c_void, out: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_fr_double(ctx: *mut JjCtx, a: *const c_void, out: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_fq_invert(ctx: *mut JjCtx, a: *const c_void, out: *mut c_void, ok: *mut u8, n: usize, flags: u32) -> i32;
    pub fn jj_fr_invert(ctx: *mut JjCtx, a: *const c_void, out: *mut c_void, ok: *mut u8, n: usize, flags: u32) -> i32;
    pub fn jj_fq_sqrt(ctx: *mut JjCtx, a: *const c_void, out: *mut c_void, ok: *mut u8, n: usize, flags: u32) -> i32;
    pub fn jj_fr_sqrt(ctx: *mut JjCtx, a: *const c_void, out: *mut c_void, ok: *mut u8, n: usize, flags: u32) -> i32;
    pub fn jj_fq_to_bytes(ctx: *mut JjCtx, a: *const c_void, out32: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_fr_to_bytes(ctx: *mut JjCtx, a: *const c_void, out32: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_fq_from_bytes(ctx: *mut JjCtx, in32: *const c_void, out: *mut c_void, ok: *mut u8, n: usize, flags: u32) -> i32;
    pub fn jj_fr_from_bytes(ctx: *mut JjCtx, in32: *const c_void, out: *mut c_void, ok: *mut u8, n: usize, flags: u32) -> i32;
    pub fn jj_fq_from_bytes_wide(ctx: *mut JjCtx, in64: *const c_void, out: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_fr_from_bytes_wide(ctx: *mut JjCtx, in64: *const c_void, out: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_fq_stream(ctx: *mut JjCtx, seed: u64, first: usize, out: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_fr_stream(ctx: *mut JjCtx, seed: u64, first: usize, out: *mut c_void, n: usize, flags: u32) -> i32;

    // point batches: ExtendedPoint 160 B, AffinePoint 64 B, ExtendedNiels 128 B, AffineNiels 96 B (Montgomery limbs)
    pub fn jj_point_double(ctx: *mut JjCtx, p_ext: *const c_void, out_ext: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_point_add(ctx: *mut JjCtx, p_ext: *const c_void, q_ext: *const c_void, out_ext: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_point_add_niels(ctx: *mut JjCtx, p_ext: *const c_void, q_niels: *const c_void, out_ext: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_point_add_affine_niels(ctx: *mut JjCtx, p_ext: *const c_void, q_aniels: *const c_void, out_ext: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_point_to_niels(ctx: *mut JjCtx, p_ext: *const c_void, out_niels: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_affine_to_niels(ctx: *mut JjCtx, p_affine: *const c_void, out_aniels: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_scalar_mul(ctx: *mut JjCtx, points_ext: *const c_void, scalars32: *const c_void, out: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_scalar_mul_encoded(ctx: *mut JjCtx, points32: *const c_void, scalars32: *const c_void, out: *mut c_void, ok: *mut u8, n: usize, flags: u32) -> i32;
    pub fn jj_scalar_mul_fixed(ctx: *mut JjCtx, base_affine: *const c_void, scalars32: *const c_void, out: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_point_sum(ctx: *mut JjCtx, points_ext: *const c_void, out: *mut c_void, groups: usize, group_size: usize, flags: u32) -> i32;
    pub fn jj_mul_by_cofactor(ctx: *mut JjCtx, p_ext: *const c_void, out_ext: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_point_neg(ctx: *mut JjCtx, p_ext: *const c_void, out_ext: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_point_eq(ctx: *mut JjCtx, p_ext: *const c_void, q_ext: *const c_void, flags_out: *mut u8, n: usize, flags: u32) -> i32;
    pub fn jj_affine_to_extended(ctx: *mut JjCtx, p_affine: *const c_void, out_ext: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_batch_normalize(ctx: *mut JjCtx, in_ext: *const c_void, out_affine: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_batch_normalize_extended(ctx: *mut JjCtx, in_ext: *const c_void, out_ext: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_affine_to_bytes(ctx: *mut JjCtx, in_affine: *const c_void, out32: *mut c_void, n: usize, flags: u32) -> i32;
    pub fn jj_batch_from_bytes(ctx: *mut JjCtx, in32: *const c_void, out_affine: *mut c_void, ok: *mut u8, n: usize, flags: u32) -> i32;
    pub fn jj_is_torsion_free(ctx: *mut JjCtx, p_ext: *const c_void, flags_out: *mut u8, n: usize, flags: u32) -> i32;
    pub fn jj_is_prime_order(ctx: *mut JjCtx, p_ext: *const c_void, flags_out: *mut u8, n: usize, flags: u32) -> i32;
    pub fn jj_is_identity(ctx: *mut JjCtx, p_ext: *const c_void, flags_out: *mut u8, n: usize, flags: u32) -> i32;
    pub fn jj_is_small_order(ctx: *mut JjCtx, p_ext: *const c_void, flags_out: *mut u8, n: usize, flags: u32) -> i32;

    // one process per GPU: shard the batch, gather every rank's results (ncclAllGather or fused P2P stores)
    pub fn jj_comm_unique_id(id128: *mut c_void) -> i32;
    pub fn jj_comm_init(ctx: *mut JjCtx, nranks: i32, rank: i32, id128: *const c_void) -> i32;
    pub fn jj_comm_destroy(ctx: *mut JjCtx) -> i32;
    pub fn jj_scalar_mul_sharded(ctx: *mut JjCtx, points_ext_local: *const c_void, scalars32_local: *const c_void, out_all: *mut c_void, n_local: usize, flags: u32) -> i32;
    pub fn jj_scalar_mul_sharded_n(ctx: *mut JjCtx, points_ext_local: *const c_void, scalars32_local: *const c_void, out_all: *mut c_void, out_local_host: *mut c_void, n_total: usize, flags: u32) -> i32;
    pub fn jj_point_sum_sharded(ctx: *mut JjCtx, points_ext_local: *const c_void, out: *mut c_void, n_local: usize, flags: u32) -> i32;
    pub fn jj_ipc_export(ctx: *mut JjCtx, dptr: *const c_void, handle64: *mut c_void) -> i32;
    pub fn jj_ipc_open(ctx: *mut JjCtx, handle64: *const c_void, dptr: *mut *mut c_void) -> i32;
    pub fn jj_ipc_close(ctx: *mut JjCtx, dptr: *mut c_void) -> i32;
    pub fn jj_comm_set_peer_outputs(ctx: *mut JjCtx, peer_out_all: *const *mut c_void, count: i32) -> i32;
}

// ---------------------------------------------------------------------------------------------------------
// Safe surface
// ---------------------------------------------------------------------------------------------------------

/// Error of a batch call: the ABI's status code and the context's message (`jj_last_error`).
#[derive(Debug, Clone)]
pub struct Error {
    pub code: i32,
    pub message: String,
}

/// One GPU, one stream, one set of staging buffers.  Not `Sync`: use one engine per thread or per GPU, as the
/// header says (a context is not thread-safe).
pub struct Engine {
    ctx: *mut JjCtx,
}

// The context owns device resources only; moving it to another thread is fine, sharing it is not.
unsafe impl Send for Engine {}

impl Drop for Engine {
    fn drop(&mut self) {
        unsafe {
            jj_destroy(self.ctx);
        }
    }
}

/// Page-locked host bytes (`Engine::pinned`), freed with the borrow of their engine still alive.
pub struct PinnedBytes<'e> {
    eng: &'e Engine,
    ptr: *mut u8,
    len: usize,
}

impl<'e> PinnedBytes<'e> {
    pub fn len(&self) -> usize {
        self.len
    }
    pub fn is_empty(&self) -> bool {
        self.len == 0
    }
    pub fn as_ptr(&self) -> *const u8 {
        self.ptr
    }
    pub fn as_mut_ptr(&mut self) -> *mut u8 {
        self.ptr
    }
    pub fn as_slice(&self) -> &[u8] {
        unsafe { core::slice::from_raw_parts(self.ptr, self.len) }
    }
    pub fn as_mut_slice(&mut self) -> &mut [u8] {
        unsafe { core::slice::from_raw_parts_mut(self.ptr, self.len) }
    }
}

impl<'e> Drop for PinnedBytes<'e> {
    fn drop(&mut self) {
        unsafe {
            jj_host_free(self.eng.ctx, self.ptr as *mut c_void);
        }
    }
}

impl Engine {
    /// `Err(JJ_ERR_NO_DEVICE)` when there is no sm_100 GPU: the library has no CPU path — use the reference's own
    /// scalar API (`&point * &scalar`) in that case.
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut ctx = core::ptr::null_mut();
        let rc = unsafe { jj_init(device, &mut ctx) };
        if rc != JJ_OK {
            return Err(Error { code: rc, message: format!("jj_init failed with {rc}") });
        }
        Ok(Engine { ctx })
    }

    fn check(&self, rc: i32) -> Result<(), Error> {
        if rc == JJ_OK {
            return Ok(());
        }
        let message = unsafe { CStr::from_ptr(jj_last_error(self.ctx)) }.to_string_lossy().into_owned();
        Err(Error { code: rc, message })
    }

    /// `points[i] * scalars[i]` for every `i` (`&AffinePoint * &Fr`, `src/lib.rs:1109-1115`), as 32-byte encodings.
    ///
    /// Panics on a length mismatch like the reference's batch helpers (`assert_eq!`, `src/lib.rs:841`).
    pub fn batch_mul_vartime(&self, points: &[AffinePoint], scalars: &[Fr]) -> Result<Vec<[u8; 32]>, Error> {
        assert_eq!(points.len(), scalars.len());
        let enc: Vec<[u8; 32]> = points.iter().map(|p| p.to_bytes()).collect(); // src/lib.rs:455-464
        let (out, ok) = self.batch_mul_encoded_vartime(&enc, scalars)?;
        debug_assert!(ok.iter().all(|&b| b == 1), "encodings of valid points always decode");
        Ok(out)
    }

    /// Wire format in, wire format out: `encodings[i]` is decoded on the device with the ZIP-216 rule
    /// (`AffinePoint::from_bytes`, `src/lib.rs:470-483`), multiplied by `scalars[i]`, and encoded again.
    /// `ok[i] == 0` marks an encoding the reference would reject (`CtOption::none`); its output is unspecified.
    pub fn batch_mul_encoded_vartime(&self, encodings: &[[u8; 32]], scalars: &[Fr]) -> Result<(Vec<[u8; 32]>, Vec<u8>), Error> {
        self.batch_mul_encoded_flags(encodings, scalars, 0)
    }

    /// The same with the kernel's constant-time-in-the-scalar mode (`JJ_CONST_TIME`): no branch and no memory address
    /// depends on `scalars` (window table scanned, sign by selects, every addition executed -- the batch analogue of the
    /// reference's "always add P or identity", `src/lib.rs:356-379`).  About 4 % slower; the points stay public data.
    /// This is the entry point that keeps the reference's policy (`src/lib.rs:12-17`), hence no `_vartime` suffix.
    pub fn batch_mul_encoded(&self, encodings: &[[u8; 32]], scalars: &[Fr]) -> Result<(Vec<[u8; 32]>, Vec<u8>), Error> {
        self.batch_mul_encoded_flags(encodings, scalars, JJ_CONST_TIME)
    }

    /// `SubgroupPoint::from_bytes(..)` semantics for the decode (`src/lib.rs:1427-1429`): `ok[i]` also requires the point to
    /// be torsion free (decided on the device by a pairing, not by `[r]P`), then the multiplication.
    pub fn batch_mul_encoded_subgroup_vartime(&self, encodings: &[[u8; 32]], scalars: &[Fr]) -> Result<(Vec<[u8; 32]>, Vec<u8>), Error> {
        self.batch_mul_encoded_flags(encodings, scalars, JJ_CHECK_SUBGROUP)
    }

    fn batch_mul_encoded_flags(&self, encodings: &[[u8; 32]], scalars: &[Fr], flags: u32) -> Result<(Vec<[u8; 32]>, Vec<u8>), Error> {
        assert_eq!(encodings.len(), scalars.len());
        let n = encodings.len();
        let k: Vec<[u8; 32]> = scalars.iter().map(|s| s.to_bytes()).collect(); // src/fr.rs:296-308
        let mut out = vec![[0u8; 32]; n];
        let mut ok = vec![0u8; n];
        let rc = unsafe {
            jj_scalar_mul_encoded(
                self.ctx,
                encodings.as_ptr() as *const c_void,
                k.as_ptr() as *const c_void,
                out.as_mut_ptr() as *mut c_void,
                ok.as_mut_ptr(),
                n,
                JJ_OUT_BYTES | flags,
            )
        };
        self.check(rc)?;
        Ok((out, ok))
    }

    /// Page-locked buffers (`PinnedBytes`): the decode kernel reads `encodings` in place over PCIe, `scalars` (32-byte
    /// `Fr::to_bytes` forms) are uploaded while the chunk is decoded, and the results are stored into `out` in place --
    /// no staging copy on either side (38.4 instead of 39.3 ms per 2^20 on one B200).  All four buffers hold `n` units.
    pub fn batch_mul_encoded_in_place(&self, encodings: &PinnedBytes, scalars: &PinnedBytes, out: &mut PinnedBytes,
                                      ok: &mut PinnedBytes, n: usize, flags: u32) -> Result<(), Error> {
        assert!(encodings.len() >= 32 * n && scalars.len() >= 32 * n && out.len() >= 32 * n && ok.len() >= n);
        let rc = unsafe {
            jj_scalar_mul_encoded(
                self.ctx,
                encodings.as_ptr() as *const c_void,
                scalars.as_ptr() as *const c_void,
                out.as_mut_ptr() as *mut c_void,
                ok.as_mut_ptr(),
                n,
                JJ_OUT_BYTES | flags,
            )
        };
        self.check(rc)
    }

    /// `len` bytes of page-locked host memory owned by this engine's context (`jj_host_alloc`).
    pub fn pinned(&self, len: usize) -> Result<PinnedBytes<'_>, Error> {
        let mut p: *mut c_void = core::ptr::null_mut();
        self.check(unsafe { jj_host_alloc(self.ctx, len, &mut p) })?;
        Ok(PinnedBytes { eng: self, ptr: p as *mut u8, len })
    }

    /// The same, returned as points: the 32-byte results go through the reference's own batched decoder
    /// (`AffinePoint::batch_from_bytes`, `src/lib.rs:541-627`), so the values are the reference's by construction.
    pub fn batch_mul_points_vartime(&self, points: &[AffinePoint], scalars: &[Fr]) -> Result<Vec<ExtendedPoint>, Error> {
        let enc = self.batch_mul_vartime(points, scalars)?;
        Ok(AffinePoint::batch_from_bytes(enc.into_iter())
            .into_iter()
            .map(|p| ExtendedPoint::from(Option::<AffinePoint>::from(p).expect("device output is a valid encoding")))
            .collect())
    }

    /// `SubgroupPoint` version (`src/lib.rs:1231-1237`): results of multiplying prime-order points stay in the
    /// subgroup, so `from_bytes_unchecked` is sound here and skips the `[r]P` check of `from_bytes` (`:1427-1429`).
    pub fn batch_mul_subgroup_vartime(&self, points: &[SubgroupPoint], scalars: &[Fr]) -> Result<Vec<SubgroupPoint>, Error> {
        assert_eq!(points.len(), scalars.len());
        let enc: Vec<[u8; 32]> = points.iter().map(|p| p.to_bytes()).collect();
        let (out, _ok) = self.batch_mul_encoded_vartime(&enc, scalars)?;
        Ok(out
            .iter()
            .map(|b| Option::<SubgroupPoint>::from(SubgroupPoint::from_bytes_unchecked(b)).expect("valid encoding"))
            .collect())
    }

    /// `[scalars[i]] base` with one shared base (`&AffinePoint * &Fr`, `AffineNielsPoint::multiply`,
    /// `src/lib.rs:271-295`): the device builds a 12-bit window table of the base (4.1 MB, L2-resident)
    /// once per base and every scalar is then 22 table additions.
    ///
    /// The base crosses the ABI as Montgomery limbs; out of crate they are obtained without touching private
    /// fields by decoding the base's encoding on the device (`jj_batch_from_bytes` into a device buffer).
    pub fn batch_mul_fixed_vartime(&self, base: &AffinePoint, scalars: &[Fr]) -> Result<Vec<[u8; 32]>, Error> {
        let n = scalars.len();
        let k: Vec<[u8; 32]> = scalars.iter().map(|s| s.to_bytes()).collect();
        let enc = base.to_bytes();
        let mut base_limbs = [0u8; 64];
        let mut ok = [0u8; 1];
        let rc = unsafe {
            jj_batch_from_bytes(self.ctx, enc.as_ptr() as *const c_void, base_limbs.as_mut_ptr() as *mut c_void, ok.as_mut_ptr(), 1, 0)
        };
        self.check(rc)?;
        let mut out = vec![[0u8; 32]; n];
        let rc = unsafe {
            jj_scalar_mul_fixed(
                self.ctx,
                base_limbs.as_ptr() as *const c_void,
                k.as_ptr() as *const c_void,
                out.as_mut_ptr() as *mut c_void,
                n,
                JJ_OUT_BYTES,
            )
        };
        self.check(rc)?;
        Ok(out)
    }

    /// Decodes a batch of encodings on the device (`AffinePoint::batch_from_bytes`, `src/lib.rs:541-627`) and
    /// reports which are valid *and* torsion free — the work of `SubgroupPoint::from_bytes` (`:1427-1429`) for a
    /// whole batch.  Valid entries can then be rebuilt with `from_bytes_unchecked`.
    pub fn batch_check_subgroup_encodings_vartime(&self, encodings: &[[u8; 32]]) -> Result<Vec<bool>, Error> {
        let n = encodings.len();
        // [1] * P through the encoded entry point leaves the decoded points as ExtendedPoint limbs in `ext`.
        let one = vec![Fr::one().to_bytes(); n];
        let mut ext = vec![[0u8; 160]; n];
        let mut ok = vec![0u8; n];
        let rc = unsafe {
            jj_scalar_mul_encoded(self.ctx, encodings.as_ptr() as *const c_void, one.as_ptr() as *const c_void,
                                  ext.as_mut_ptr() as *mut c_void, ok.as_mut_ptr(), n, 0)
        };
        self.check(rc)?;
        let mut tf = vec![0u8; n];
        let rc = unsafe { jj_is_torsion_free(self.ctx, ext.as_ptr() as *const c_void, tf.as_mut_ptr(), n, 0) };
        self.check(rc)?;
        Ok(ok.iter().zip(tf.iter()).map(|(&a, &b)| a == 1 && b == 1).collect())
    }

    /// `a[i] * b[i]` in the base field, exchanged as canonical bytes (`JJ_CANON`; `Fq::to_bytes` /
    /// `Fq::from_bytes`).  Worth it only for large batches that are already in byte form: the conversions cost
    /// more than the products.
    pub fn fq_batch_mul(&self, a: &[Fq], b: &[Fq]) -> Result<Vec<Fq>, Error> {
        assert_eq!(a.len(), b.len());
        let n = a.len();
        let ab: Vec<[u8; 32]> = a.iter().map(|x| x.to_bytes()).collect();
        let bb: Vec<[u8; 32]> = b.iter().map(|x| x.to_bytes()).collect();
        let mut out = vec![[0u8; 32]; n];
        let rc = unsafe {
            jj_fq_mul(self.ctx, ab.as_ptr() as *const c_void, bb.as_ptr() as *const c_void, out.as_mut_ptr() as *mut c_void, n, JJ_CANON)
        };
        self.check(rc)?;
        Ok(out.iter().map(|x| Option::<Fq>::from(Fq::from_bytes(x)).expect("canonical output")).collect())
    }

    /// `1 / a[i]` with `CtOption::none` for zero (`Fr::invert`, `src/fr.rs:438-540`), one batched inversion.
    pub fn fr_batch_invert(&self, a: &[Fr]) -> Result<Vec<CtOption<Fr>>, Error> {
        let n = a.len();
        let ab: Vec<[u8; 32]> = a.iter().map(|x| x.to_bytes()).collect();
        let mut out = vec![[0u8; 32]; n];
        let mut ok = vec![0u8; n];
        let rc = unsafe {
            jj_fr_invert(self.ctx, ab.as_ptr() as *const c_void, out.as_mut_ptr() as *mut c_void, ok.as_mut_ptr(), n, JJ_CANON)
        };
        self.check(rc)?;
        Ok(out
            .iter()
            .zip(ok.iter())
            .map(|(x, &f)| CtOption::new(Option::<Fr>::from(Fr::from_bytes(x)).unwrap_or_else(Fr::zero), f.into()))
            .collect())
    }
}

/// Inside a fork of the reference crate `Fr(pub(crate) [u64; 4])` (`src/fr.rs:23`) is visible and is exactly the
/// `JJ_SCALAR_MONT` scalar.  `Fq` is NOT: it is `bls12_381::Scalar` (`src/lib.rs:62`), whose limbs are private even to a
/// fork of jubjub, and whose struct layout Rust does not specify (no `#[repr(C)]` / `#[repr(transparent)]` upstream).
/// `ExtendedPoint` is five of them in declaration order (`src/lib.rs:139-145`); in practice that is 5 x `[u64; 4]`
/// Montgomery limbs, byte-identical to the ABI's 160-byte unit, but the pointer cast below RELIES ON UNSPECIFIED LAYOUT.
/// A fork that wants it to be sound must also vendor `bls12_381` with `#[repr(transparent)]` on `Scalar` (and lift the
/// crate's `#![deny(unsafe_code)]`, `src/lib.rs:24`, for this one module), or stay on the wire-format path above.
/// With that caveat, the in-crate entry point is one call:
///
/// ```ignore
/// impl ExtendedPoint {
///     pub fn batch_mul_vartime(engine: &Engine, points: &[ExtendedPoint], scalars: &[Fr]) -> Vec<ExtendedPoint> {
///         assert_eq!(points.len(), scalars.len());
///         let mut out = vec![ExtendedPoint::identity(); points.len()];
///         let rc = unsafe { jj_scalar_mul(engine.ctx, points.as_ptr().cast(), scalars.as_ptr().cast(),
///                                         out.as_mut_ptr().cast(), points.len(), JJ_SCALAR_MONT) };
///         assert_eq!(rc, JJ_OK);
///         out            // projectively equal to the reference's ladder; identical after normalisation
///     }
/// }
/// ```
pub mod in_crate_notes {}
