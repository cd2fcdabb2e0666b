/*
 * jj_oracle.h -- CPU restatement of the zkcrypto/jubjub hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle and the CPU baseline
 * ("reference algorithm, C restatement") for jubjub_b200.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product library (libjubjub_b200.so) never links or calls it.
 *
 * The reference is Rust (no toolchain in this image) and its base field Fq is
 * the un-vendored crate bls12_381 0.8.0 (`Scalar`, Cargo.lock:50-53), so the
 * reference itself cannot be compiled here.  Every function below cites the
 * reference file:line whose algorithm it follows; Fq follows src/fr.rs with
 * q's constants.  Pinned against the reference's own
 * known-answer tests by tests/test_oracle_kat.py (every KAT of src/fr.rs:787-1244
 * and src/lib.rs:1456-1935).  PARITY UNPINNED for one part, stated plainly: the
 * reference holds no known-answer test for Fq on its own (tests/fq_blackbox.rs is
 * property-only) and bls12_381 cannot be run here, so Fq limb-level results are
 * pinned only indirectly -- through the point KATs, which exercise Fq
 * mul/square/add/sub/invert/sqrt/to_bytes/from_bytes -- plus the big-integer model
 * (oracle/model.py) and the uniqueness of fully reduced Montgomery limbs.
 *
 * Layout: field element = 4 x u64 little-endian limbs.  Unless a function says
 * "canonical" or "bytes", limbs are in the reference's internal Montgomery form
 * (R = 2^256), i.e. exactly what `Fr(pub(crate) [u64; 4])` holds (src/fr.rs:23).
 */
#ifndef JJ_ORACLE_H
#define JJ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[4]; } jo_fe;

/* src/lib.rs:81-84, 139-145, 255-259, 327-332 (declaration order kept). */
typedef struct { jo_fe u, v; } jo_affine;
typedef struct { jo_fe u, v, z, t1, t2; } jo_extended;
typedef struct { jo_fe v_plus_u, v_minus_u, t2d; } jo_affine_niels;
typedef struct { jo_fe v_plus_u, v_minus_u, z, t2d; } jo_extended_niels;

/* which = 0: Fq (base field, bls12_381::Scalar), which = 1: Fr (src/fr.rs). */
enum { JO_FQ = 0, JO_FR = 1 };

/* ---- field: single element --------------------------------------------- */
void jo_fe_modulus(int which, jo_fe *out);                        /* raw limbs of m            */
void jo_fe_one(int which, jo_fe *out);                            /* R mod m   src/fr.rs:217   */
void jo_fe_mul(int which, const jo_fe *a, const jo_fe *b, jo_fe *out);   /* src/fr.rs:592-616 */
void jo_fe_square(int which, const jo_fe *a, jo_fe *out);                /* src/fr.rs:353-381 */
void jo_fe_add(int which, const jo_fe *a, const jo_fe *b, jo_fe *out);   /* src/fr.rs:638-647 */
void jo_fe_sub(int which, const jo_fe *a, const jo_fe *b, jo_fe *out);   /* src/fr.rs:620-634 */
void jo_fe_neg(int which, const jo_fe *a, jo_fe *out);                   /* src/fr.rs:651-665 */
void jo_fe_double(int which, const jo_fe *a, jo_fe *out);                /* src/fr.rs:261-263 */
int  jo_fe_invert(int which, const jo_fe *a, jo_fe *out);   /* 1 = some, 0 = none; src/fr.rs:438-540 */
void jo_fe_pow_vartime(int which, const jo_fe *a, const uint64_t e[4], jo_fe *out); /* src/fr.rs:422-434 */
int  jo_fe_sqrt(int which, const jo_fe *a, jo_fe *out);     /* 1 = some; src/fr.rs:384-399 / [ext] */
void jo_fe_from_raw(int which, const uint64_t v[4], jo_fe *out);         /* src/fr.rs:347-349 */
void jo_fe_to_bytes(int which, const jo_fe *a, uint8_t out[32]);         /* src/fr.rs:296-308 */
int  jo_fe_from_bytes(int which, const uint8_t in[32], jo_fe *out);      /* src/fr.rs:268-292 */
void jo_fe_from_bytes_wide(int which, const uint8_t in[64], jo_fe *out); /* src/fr.rs:312-343 */

/* ---- field: batches (op: 0 mul, 1 square, 2 add, 3 sub, 4 neg, 5 double) - */
void jo_fe_batch(int which, int op, const jo_fe *a, const jo_fe *b, jo_fe *out, size_t n);
void jo_fe_batch_invert(int which, const jo_fe *a, jo_fe *out, uint8_t *ok, size_t n);
void jo_fe_batch_to_bytes(int which, const jo_fe *a, uint8_t *out, size_t n);
void jo_fe_batch_from_bytes(int which, const uint8_t *in, jo_fe *out, uint8_t *ok, size_t n);
/* SplitMix64 stream -> from_bytes_wide; element i uses outputs [8i, 8i+8). */
void jo_fe_stream(int which, uint64_t seed, size_t first, size_t n, jo_fe *out);

/* ---- points -------------------------------------------------------------- */
void jo_ext_identity(jo_extended *out);                                  /* src/lib.rs:680-688   */
void jo_generator(jo_affine *out);                                       /* src/lib.rs:1380-1396 */
void jo_affine_to_extended(const jo_affine *a, jo_extended *out);        /* src/lib.rs:214-226   */
void jo_ext_to_affine(const jo_extended *p, jo_affine *out);             /* src/lib.rs:227-243   */
void jo_ext_to_niels(const jo_extended *p, jo_extended_niels *out);      /* src/lib.rs:728-735   */
void jo_affine_to_niels(const jo_affine *p, jo_affine_niels *out);       /* src/lib.rs:652-658   */
void jo_ext_neg(const jo_extended *p, jo_extended *out);                 /* src/lib.rs:196-210   */
void jo_ext_double(const jo_extended *p, jo_extended *out);              /* src/lib.rs:739-828   */
void jo_ext_add_niels(const jo_extended *p, const jo_extended_niels *n, jo_extended *out);     /* :883-920 */
void jo_ext_sub_niels(const jo_extended *p, const jo_extended_niels *n, jo_extended *out);     /* :922-940 */
void jo_ext_add_affine_niels(const jo_extended *p, const jo_affine_niels *n, jo_extended *out);/* :944-968 */
void jo_ext_sub_affine_niels(const jo_extended *p, const jo_affine_niels *n, jo_extended *out);/* :970-988 */
void jo_ext_add(const jo_extended *p, const jo_extended *q, jo_extended *out);                 /* :992-999 */
void jo_ext_sub(const jo_extended *p, const jo_extended *q, jo_extended *out);                 /* :1001-1008 */
/* [k]P, bitwise MSB-first ladder over bits 251..0 of 32 LE bytes (src/lib.rs:356-379, 830-833). */
void jo_ext_mul_bits(const jo_extended *p, const uint8_t by[32], jo_extended *out);
/* AffineNielsPoint::multiply (src/lib.rs:271-295). */
void jo_affine_niels_mul_bits(const jo_affine_niels *n, const uint8_t by[32], jo_extended *out);
void jo_ext_mul_by_cofactor(const jo_extended *p, jo_extended *out);     /* src/lib.rs:722-724   */
int  jo_ext_is_identity(const jo_extended *p);                           /* src/lib.rs:691-696   */
int  jo_ext_is_small_order(const jo_extended *p);                        /* src/lib.rs:699-705   */
int  jo_ext_is_torsion_free(const jo_extended *p);                       /* src/lib.rs:709-711   */
int  jo_ext_eq(const jo_extended *p, const jo_extended *q);              /* src/lib.rs:153-163   */
int  jo_affine_is_on_curve(const jo_affine *p);                          /* src/lib.rs:670-675   */
void jo_affine_to_bytes(const jo_affine *p, uint8_t out[32]);            /* src/lib.rs:455-464   */
int  jo_affine_from_bytes(const uint8_t in[32], int zip216, jo_affine *out); /* src/lib.rs:492-534 */

/* ---- point batches -------------------------------------------------------- */
/* batch_normalize: one inversion per call, zeros skipped like ff::BatchInverter (src/lib.rs:840-858). */
void jo_batch_normalize(const jo_extended *p, jo_affine *out, size_t n);
void jo_batch_to_bytes(const jo_affine *p, uint8_t *out, size_t n);
void jo_batch_from_bytes(const uint8_t *in, jo_affine *out, uint8_t *ok, size_t n); /* src/lib.rs:541-627 */
void jo_batch_double(const jo_extended *p, jo_extended *out, size_t n);
void jo_batch_add(const jo_extended *p, const jo_extended *q, jo_extended *out, size_t n);
void jo_batch_add_niels(const jo_extended *p, const jo_extended_niels *q, jo_extended *out, size_t n);
void jo_batch_add_affine_niels(const jo_extended *p, const jo_affine_niels *q, jo_extended *out, size_t n);
void jo_batch_is_torsion_free(const jo_extended *p, uint8_t *out, size_t n, int nthreads);
/* out[i] = [scalars[i]] points[i], reference ladder, contiguous shards over nthreads pthreads. */
void jo_batch_scalar_mul(const jo_extended *points, const uint8_t *scalars32,
                         jo_extended *out, size_t n, int nthreads);
/* out[i] = [scalars[i]] base via AffineNielsPoint::multiply (src/lib.rs:1109-1115). */
void jo_batch_scalar_mul_fixed(const jo_affine *base, const uint8_t *scalars32,
                               jo_extended *out, size_t n, int nthreads);
/* Timed loops for bench.py's cpu_baseline (returns seconds of the best of `reps`). */
double jo_time_fe_mul(int which, size_t n, int reps);
double jo_time_scalar_mul(const jo_extended *points, const uint8_t *scalars32,
                          jo_extended *out, size_t n, int nthreads, int reps);

#ifdef __cplusplus
}
#endif
#endif
