/*
 * jj_oracle.c -- CPU restatement of the zkcrypto/jubjub hot path (see jj_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: parity oracle + CPU baseline.  Never linked into,
 * loaded by, or called from the product library.
 *
 * Algorithms follow the reference line by line in *structure* (same limb width,
 * same schoolbook product, same four-round Montgomery reduction with the second
 * carry word, same mask-based conditional subtract, same bitwise 252-step
 * ladder), re-expressed in C with `unsigned __int128`.  Citations are relative
 * to /root/reference.
 */
#include "jj_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef unsigned __int128 u128;

/* ---- limb primitives: src/util.rs:3-20 ------------------------------------ */
static inline uint64_t adc(uint64_t a, uint64_t b, uint64_t *carry) {
    u128 t = (u128)a + b + *carry;
    *carry = (uint64_t)(t >> 64);
    return (uint64_t)t;
}
/* borrow is 0 or all-ones (src/util.rs:10-13 uses borrow >> 63). */
static inline uint64_t sbb(uint64_t a, uint64_t b, uint64_t *borrow) {
    u128 t = (u128)a - ((u128)b + (*borrow >> 63));
    *borrow = (uint64_t)(t >> 64);
    return (uint64_t)t;
}
static inline uint64_t mac(uint64_t a, uint64_t b, uint64_t c, uint64_t *carry) {
    u128 t = (u128)a + (u128)b * c + *carry;
    *carry = (uint64_t)(t >> 64);
    return (uint64_t)t;
}

/* ---- field parameters ------------------------------------------------------ */
typedef struct {
    uint64_t m[4];   /* modulus                         */
    uint64_t inv;    /* -(m^-1) mod 2^64                */
    jo_fe r, r2, r3; /* 2^256, 2^512, 2^768 mod m       */
} field_t;

/* Fq = bls12_381::Scalar [ext]; constants per SURVEY.md 8a (q-1 at src/lib.rs:1629-1634). */
static const field_t FQ = {
    {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL},
    0xfffffffeffffffffULL,
    {{0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL}},
    {{0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL}},
    {{0xc62c1807439b73afULL, 0x1b3e0d188cf06990ULL, 0x73d13c71c7b5f418ULL, 0x6e2a5bb9c8db33e9ULL}},
};
/* Fr: src/fr.rs:77-82 (MODULUS), :214 (INV), :217-238 (R, R2, R3). */
static const field_t FR = {
    {0xd0970e5ed6f72cb7ULL, 0xa6682093ccc81082ULL, 0x06673b0101343b00ULL, 0x0e7db4ea6533afa9ULL},
    0x1ba3a358ef788ef9ULL,
    {{0x25f80bb3b99607d9ULL, 0xf315d62f66b6e750ULL, 0x932514eeeb8814f4ULL, 0x09a6fc6f479155c6ULL}},
    {{0x67719aa495e57731ULL, 0x51b0cef09ce3fc26ULL, 0x69dab7fac026e9a5ULL, 0x04f6547b8d127688ULL}},
    {{0xe0d6c6563d830544ULL, 0x323e3883598d0f85ULL, 0xf0fea3004c2e2ba8ULL, 0x05874f84946737ecULL}},
};

#define AI static inline __attribute__((always_inline))

/* src/fr.rs:620-634: subtract, then add the modulus back under the borrow mask. */
AI void f_sub(const field_t *F, const jo_fe *a, const jo_fe *b, jo_fe *out) {
    uint64_t d[4], borrow = 0, carry = 0;
    for (int i = 0; i < 4; i++) d[i] = sbb(a->l[i], b->l[i], &borrow);
    for (int i = 0; i < 4; i++) out->l[i] = adc(d[i], F->m[i] & borrow, &carry);
}
/* src/fr.rs:638-647: add without looking at the top carry, then sub(MODULUS). */
AI void f_add(const field_t *F, const jo_fe *a, const jo_fe *b, jo_fe *out) {
    jo_fe s, m;
    uint64_t carry = 0;
    for (int i = 0; i < 4; i++) s.l[i] = adc(a->l[i], b->l[i], &carry);
    memcpy(m.l, F->m, sizeof m.l);
    f_sub(F, &s, &m, out);
}
/* src/fr.rs:651-665 */
AI void f_neg(const field_t *F, const jo_fe *a, jo_fe *out) {
    uint64_t borrow = 0, d[4];
    for (int i = 0; i < 4; i++) d[i] = sbb(F->m[i], a->l[i], &borrow);
    uint64_t mask = (uint64_t)((a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0) - 1;
    for (int i = 0; i < 4; i++) out->l[i] = d[i] & mask;
}
/* src/fr.rs:544-588: HAC 14.32, four rounds, carry2 carried between rounds. */
AI void f_mont_reduce(const field_t *F, uint64_t t[8], jo_fe *out) {
    uint64_t carry2 = 0;
    for (int i = 0; i < 4; i++) {
        uint64_t k = t[i] * F->inv, carry = 0;
        (void)mac(t[i], k, F->m[0], &carry);
        for (int j = 1; j < 4; j++) t[i + j] = mac(t[i + j], k, F->m[j], &carry);
        u128 s = (u128)t[i + 4] + carry2 + carry; /* adc(r, carry2, carry) */
        t[i + 4] = (uint64_t)s;
        carry2 = (uint64_t)(s >> 64);
    }
    jo_fe hi = {{t[4], t[5], t[6], t[7]}}, m;
    memcpy(m.l, F->m, sizeof m.l);
    f_sub(F, &hi, &m, out);
}
/* src/fr.rs:592-616: 4x4 schoolbook, then reduce. */
AI void f_mul(const field_t *F, const jo_fe *a, const jo_fe *b, jo_fe *out) {
    uint64_t t[8] = {0};
    for (int i = 0; i < 4; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < 4; j++) t[i + j] = mac(t[i + j], a->l[i], b->l[j], &carry);
        t[i + 4] = carry;
    }
    f_mont_reduce(F, t, out);
}
/* src/fr.rs:353-381: off-diagonal products once, doubled by shifting, plus diagonal. */
AI void f_square(const field_t *F, const jo_fe *a, jo_fe *out) {
    uint64_t t[8] = {0}, carry;
    for (int i = 0; i < 3; i++) {
        carry = 0;
        for (int j = i + 1; j < 4; j++) t[i + j] = mac(t[i + j], a->l[i], a->l[j], &carry);
        t[i + 4] = carry;
    }
    t[7] = t[6] >> 63;
    for (int i = 6; i >= 2; i--) t[i] = (t[i] << 1) | (t[i - 1] >> 63);
    t[1] <<= 1;
    carry = 0;
    for (int i = 0; i < 4; i++) {
        t[2 * i] = mac(t[2 * i], a->l[i], a->l[i], &carry);
        t[2 * i + 1] = adc(0, t[2 * i + 1], &carry);
    }
    f_mont_reduce(F, t, out);
}
AI int f_is_zero(const jo_fe *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
AI int f_eq(const jo_fe *a, const jo_fe *b) {
    return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}
/* src/fr.rs:422-434 */
static void f_pow_vartime(const field_t *F, const jo_fe *a, const uint64_t e[4], jo_fe *out) {
    jo_fe res = F->r;
    for (int w = 3; w >= 0; w--)
        for (int i = 63; i >= 0; i--) {
            f_square(F, &res, &res);
            if ((e[w] >> i) & 1) f_mul(F, &res, a, &res);
        }
    *out = res;
}
/* src/fr.rs:296-308: one reduction of (a, 0). */
static void f_to_bytes(const field_t *F, const jo_fe *a, uint8_t out[32]) {
    uint64_t t[8] = {a->l[0], a->l[1], a->l[2], a->l[3], 0, 0, 0, 0};
    jo_fe c;
    f_mont_reduce(F, t, &c);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(c.l[i] >> (8 * j));
}
static void load_le(const uint8_t *in, uint64_t *l, int nlimbs) {
    for (int i = 0; i < nlimbs; i++) {
        uint64_t v = 0;
        for (int j = 7; j >= 0; j--) v = (v << 8) | in[8 * i + j];
        l[i] = v;
    }
}
/* src/fr.rs:268-292: canonical check by trial subtraction, then * R2. */
static int f_from_bytes(const field_t *F, const uint8_t in[32], jo_fe *out) {
    jo_fe tmp;
    load_le(in, tmp.l, 4);
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) (void)sbb(tmp.l[i], F->m[i], &borrow);
    f_mul(F, &tmp, &F->r2, out);
    return (int)(borrow & 1);
}
/* src/fr.rs:312-343: d0*R2 + d1*R3. */
static void f_from_bytes_wide(const field_t *F, const uint8_t in[64], jo_fe *out) {
    uint64_t l[8];
    load_le(in, l, 8);
    jo_fe d0 = {{l[0], l[1], l[2], l[3]}}, d1 = {{l[4], l[5], l[6], l[7]}};
    f_mul(F, &d0, &F->r2, &d0);
    f_mul(F, &d1, &F->r3, &d1);
    f_add(F, &d0, &d1, out);
}

/* Fr::invert addition chain, src/fr.rs:438-540, as data.
 * Phase 1 builds t[] from products; phase 2 is "square n times, multiply by t[k]". */
static int fr_invert(const jo_fe *self, jo_fe *out) {
    const field_t *F = &FR;
    jo_fe t[20]; /* t[0..19]; index 20 ("self") handled separately */
    jo_fe t0, t1, t3;
#define M(dst, x, y) f_mul(F, &(x), &(y), &(dst))
    f_square(F, self, &t1);
    f_square(F, &t1, &t0);
    M(t3, t0, t1);
    M(t[6], t3, *self);
    M(t[7], t[6], t1);
    M(t[12], t[7], t3);
    M(t[13], t[12], t0);
    M(t[16], t[12], t3);
    M(t[2], t[13], t3);
    M(t[15], t[16], t3);
    M(t[19], t[2], t0);
    M(t[9], t[15], t3);
    M(t[18], t[9], t3);
    M(t[14], t[18], t1);
    M(t[4], t[18], t0);
    M(t[8], t[18], t3);
    M(t[17], t[14], t3);
    M(t[11], t[8], t3);
    M(t1, t[17], t3);
    M(t[5], t[11], t3);
    M(t3, t[5], t0);
    f_square(F, &t[5], &t0);
    t[1] = t1;
    t[3] = t3;
#undef M
    /* (squarings, operand): operand 20 = self. First entry continues from t0 = t5^2. */
    static const uint8_t chain[][2] = {
        {5, 3},  {6, 8},   {7, 19}, {6, 13}, {8, 14}, {6, 18},  {7, 17}, {5, 16}, {3, 20},
        {11, 11}, {8, 5},  {5, 15}, {8, 20}, {12, 13}, {7, 9},  {5, 15}, {14, 14}, {5, 13},
        {2, 20}, {6, 20},  {9, 7},  {6, 12}, {8, 11}, {3, 20},  {12, 9}, {11, 8}, {8, 7},
        {4, 6},  {10, 5},  {7, 3},  {6, 4},  {7, 3},  {5, 2},   {6, 2},  {7, 1},
    };
    for (size_t s = 0; s < sizeof chain / sizeof chain[0]; s++) {
        for (int k = 0; k < chain[s][0]; k++) f_square(F, &t0, &t0);
        f_mul(F, &t0, chain[s][1] == 20 ? self : &t[chain[s][1]], &t0);
    }
    *out = t0;
    return !f_is_zero(self);
}
/* [ext] bls12_381::Scalar::invert = a^(q-2) (published: Fermat; addition chain differs
 * but the fully reduced result is unique).  Call sites src/lib.rs:236, 514. */
static int fq_invert(const jo_fe *a, jo_fe *out) {
    static const uint64_t e[4] = {0xfffffffeffffffffULL, 0x53bda402fffe5bfeULL,
                                  0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
    f_pow_vartime(&FQ, a, e, out);
    return !f_is_zero(a);
}
/* Fr::sqrt, src/fr.rs:384-399: a^((r+1)/4). */
static int fr_sqrt(const jo_fe *a, jo_fe *out) {
    static const uint64_t e[4] = {0xb425c397b5bdcb2eULL, 0x299a0824f3320420ULL,
                                  0x4199cec0404d0ec0ULL, 0x039f6d3a994cebeaULL};
    jo_fe s, s2;
    f_pow_vartime(&FR, a, e, &s);
    f_square(&FR, &s, &s2);
    *out = s;
    return f_eq(&s2, a);
}
/* [ext] bls12_381::Scalar::sqrt: Tonelli-Shanks, q-1 = 2^32 * t, 2-Sylow generator 7^t.
 * Which of the two roots is returned is irrelevant to callers on this path: the sign is
 * fixed from the parity afterwards (src/lib.rs:518-520). */
static int fq_sqrt(const jo_fe *a, jo_fe *out) {
    const field_t *F = &FQ;
    static const uint64_t t[4] = {0xfffe5bfeffffffffULL, 0x09a1d80553bda402ULL,
                                  0x299d7d483339d808ULL, 0x0000000073eda753ULL};
    static const uint64_t t_plus1_half[4] = {0x7fff2dff80000000ULL, 0x04d0ec02a9ded201ULL,
                                             0x94cebea4199cec04ULL, 0x0000000039f6d3a9ULL};
    if (f_is_zero(a)) { memset(out, 0, sizeof *out); return 1; }
    jo_fe seven, z, x, b;
    uint64_t seven_raw[4] = {7, 0, 0, 0};
    jo_fe sr = {{seven_raw[0], 0, 0, 0}};
    f_mul(F, &sr, &F->r2, &seven);
    f_pow_vartime(F, &seven, t, &z);
    f_pow_vartime(F, a, t_plus1_half, &x);
    f_pow_vartime(F, a, t, &b);
    int m = 32;
    while (!f_eq(&b, &F->r)) {
        int i = 0;
        jo_fe bb = b;
        while (!f_eq(&bb, &F->r)) {
            f_square(F, &bb, &bb);
            if (++i >= m) return 0; /* non-residue */
        }
        jo_fe w = z;
        for (int k = 0; k < m - i - 1; k++) f_square(F, &w, &w);
        f_mul(F, &x, &w, &x);
        f_square(F, &w, &z);
        f_mul(F, &b, &z, &b);
        m = i;
    }
    *out = x;
    return 1;
}

static const field_t *pick(int which) { return which == JO_FR ? &FR : &FQ; }

/* ---- exported single-element field API ------------------------------------ */
void jo_fe_modulus(int which, jo_fe *out) { memcpy(out->l, pick(which)->m, 32); }
void jo_fe_one(int which, jo_fe *out) { *out = pick(which)->r; }
void jo_fe_mul(int which, const jo_fe *a, const jo_fe *b, jo_fe *out) {
    if (which == JO_FR) f_mul(&FR, a, b, out); else f_mul(&FQ, a, b, out);
}
void jo_fe_square(int which, const jo_fe *a, jo_fe *out) {
    if (which == JO_FR) f_square(&FR, a, out); else f_square(&FQ, a, out);
}
void jo_fe_add(int which, const jo_fe *a, const jo_fe *b, jo_fe *out) {
    if (which == JO_FR) f_add(&FR, a, b, out); else f_add(&FQ, a, b, out);
}
void jo_fe_sub(int which, const jo_fe *a, const jo_fe *b, jo_fe *out) {
    if (which == JO_FR) f_sub(&FR, a, b, out); else f_sub(&FQ, a, b, out);
}
void jo_fe_neg(int which, const jo_fe *a, jo_fe *out) {
    if (which == JO_FR) f_neg(&FR, a, out); else f_neg(&FQ, a, out);
}
void jo_fe_double(int which, const jo_fe *a, jo_fe *out) { jo_fe_add(which, a, a, out); }
int jo_fe_invert(int which, const jo_fe *a, jo_fe *out) {
    return which == JO_FR ? fr_invert(a, out) : fq_invert(a, out);
}
void jo_fe_pow_vartime(int which, const jo_fe *a, const uint64_t e[4], jo_fe *out) {
    f_pow_vartime(pick(which), a, e, out);
}
int jo_fe_sqrt(int which, const jo_fe *a, jo_fe *out) {
    return which == JO_FR ? fr_sqrt(a, out) : fq_sqrt(a, out);
}
void jo_fe_from_raw(int which, const uint64_t v[4], jo_fe *out) {
    jo_fe t = {{v[0], v[1], v[2], v[3]}};
    jo_fe_mul(which, &t, &pick(which)->r2, out);
}
void jo_fe_to_bytes(int which, const jo_fe *a, uint8_t out[32]) { f_to_bytes(pick(which), a, out); }
int jo_fe_from_bytes(int which, const uint8_t in[32], jo_fe *out) { return f_from_bytes(pick(which), in, out); }
void jo_fe_from_bytes_wide(int which, const uint8_t in[64], jo_fe *out) { f_from_bytes_wide(pick(which), in, out); }

/* ---- field batches ---------------------------------------------------------- */
void jo_fe_batch(int which, int op, const jo_fe *a, const jo_fe *b, jo_fe *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        switch (op) {
        case 0: jo_fe_mul(which, &a[i], &b[i], &out[i]); break;
        case 1: jo_fe_square(which, &a[i], &out[i]); break;
        case 2: jo_fe_add(which, &a[i], &b[i], &out[i]); break;
        case 3: jo_fe_sub(which, &a[i], &b[i], &out[i]); break;
        case 4: jo_fe_neg(which, &a[i], &out[i]); break;
        default: jo_fe_double(which, &a[i], &out[i]); break;
        }
    }
}
void jo_fe_batch_invert(int which, const jo_fe *a, jo_fe *out, uint8_t *ok, size_t n) {
    for (size_t i = 0; i < n; i++) ok[i] = (uint8_t)jo_fe_invert(which, &a[i], &out[i]);
}
void jo_fe_batch_to_bytes(int which, const jo_fe *a, uint8_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) jo_fe_to_bytes(which, &a[i], out + 32 * i);
}
void jo_fe_batch_from_bytes(int which, const uint8_t *in, jo_fe *out, uint8_t *ok, size_t n) {
    for (size_t i = 0; i < n; i++) ok[i] = (uint8_t)jo_fe_from_bytes(which, in + 32 * i, &out[i]);
}
static inline uint64_t splitmix_at(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
void jo_fe_stream(int which, uint64_t seed, size_t first, size_t n, jo_fe *out) {
    for (size_t i = 0; i < n; i++) {
        uint8_t w[64];
        for (int k = 0; k < 8; k++) {
            uint64_t v = splitmix_at(seed, (uint64_t)(first + i) * 8 + k);
            for (int j = 0; j < 8; j++) w[8 * k + j] = (uint8_t)(v >> (8 * j));
        }
        jo_fe_from_bytes_wide(which, w, &out[i]);
    }
}

/* ---- curve constants --------------------------------------------------------- */
/* d and 2d, raw limbs src/lib.rs:399-404, 407-412; generator :1380-1396. */
static jo_fe EDWARDS_D, EDWARDS_D2, GEN_U, GEN_V;
static pthread_once_t consts_once = PTHREAD_ONCE_INIT;
static void init_consts(void) {
    static const uint64_t d[4] = {0x01065fd6d6343eb1ULL, 0x292d7f6d37579d26ULL,
                                  0xf5fd9207e6bd7fd4ULL, 0x2a9318e74bfa2b48ULL};
    static const uint64_t d2[4] = {0x020cbfadac687d62ULL, 0x525afeda6eaf3a4cULL,
                                   0xebfb240fcd7affa8ULL, 0x552631ce97f45691ULL};
    static const uint64_t gu[4] = {0xe4b3d35df1a7adfeULL, 0xcaf55d1b29bf81afULL,
                                   0x8b0f03ddd60a8187ULL, 0x62edcbb8bf3787c8ULL};
    static const uint64_t gv[4] = {0xb, 0, 0, 0};
    jo_fe_from_raw(JO_FQ, d, &EDWARDS_D);
    jo_fe_from_raw(JO_FQ, d2, &EDWARDS_D2);
    jo_fe_from_raw(JO_FQ, gu, &GEN_U);
    jo_fe_from_raw(JO_FQ, gv, &GEN_V);
}
#define CONSTS() pthread_once(&consts_once, init_consts)

#define QMUL(a, b, o) f_mul(&FQ, (a), (b), (o))
#define QSQR(a, o) f_square(&FQ, (a), (o))
#define QADD(a, b, o) f_add(&FQ, (a), (b), (o))
#define QSUB(a, b, o) f_sub(&FQ, (a), (b), (o))

/* ---- points -------------------------------------------------------------------- */
typedef struct { jo_fe u, v, z, t; } completed_t; /* src/lib.rs:1036-1041 */

/* src/lib.rs:1052-1060 */
AI void into_extended(const completed_t *c, jo_extended *out) {
    jo_extended r;
    QMUL(&c->u, &c->t, &r.u);
    QMUL(&c->v, &c->z, &r.v);
    QMUL(&c->z, &c->t, &r.z);
    r.t1 = c->u;
    r.t2 = c->v;
    *out = r;
}
void jo_ext_identity(jo_extended *out) {
    memset(out, 0, sizeof *out);
    out->v = FQ.r;
    out->z = FQ.r;
}
void jo_generator(jo_affine *out) { CONSTS(); out->u = GEN_U; out->v = GEN_V; }
void jo_affine_to_extended(const jo_affine *a, jo_extended *out) {
    out->u = a->u; out->v = a->v; out->z = FQ.r; out->t1 = a->u; out->t2 = a->v;
}
void jo_ext_to_affine(const jo_extended *p, jo_affine *out) {
    jo_fe zinv;
    fq_invert(&p->z, &zinv);
    QMUL(&p->u, &zinv, &out->u);
    QMUL(&p->v, &zinv, &out->v);
}
void jo_ext_to_niels(const jo_extended *p, jo_extended_niels *out) {
    CONSTS();
    jo_extended_niels n;
    QADD(&p->v, &p->u, &n.v_plus_u);
    QSUB(&p->v, &p->u, &n.v_minus_u);
    n.z = p->z;
    QMUL(&p->t1, &p->t2, &n.t2d);
    QMUL(&n.t2d, &EDWARDS_D2, &n.t2d);
    *out = n;
}
void jo_affine_to_niels(const jo_affine *p, jo_affine_niels *out) {
    CONSTS();
    jo_affine_niels n;
    QADD(&p->v, &p->u, &n.v_plus_u);
    QSUB(&p->v, &p->u, &n.v_minus_u);
    QMUL(&p->u, &p->v, &n.t2d);
    QMUL(&n.t2d, &EDWARDS_D2, &n.t2d);
    *out = n;
}
void jo_ext_neg(const jo_extended *p, jo_extended *out) {
    jo_extended r = *p;
    f_neg(&FQ, &p->u, &r.u);
    f_neg(&FQ, &p->t1, &r.t1);
    *out = r;
}
/* src/lib.rs:812-827 */
void jo_ext_double(const jo_extended *p, jo_extended *out) {
    jo_fe uu, vv, zz2, uv2, vpu, vmu;
    completed_t c;
    QSQR(&p->u, &uu);
    QSQR(&p->v, &vv);
    QSQR(&p->z, &zz2);
    QADD(&zz2, &zz2, &zz2);
    QADD(&p->u, &p->v, &uv2);
    QSQR(&uv2, &uv2);
    QADD(&vv, &uu, &vpu);
    QSUB(&vv, &uu, &vmu);
    QSUB(&uv2, &vpu, &c.u);
    c.v = vpu;
    c.z = vmu;
    QSUB(&zz2, &vmu, &c.t);
    into_extended(&c, out);
}
/* Shared body of the four mixed additions; `sub` swaps the Niels halves and the
 * d+c / d-c outputs (src/lib.rs:905-918 vs :927-938; :953-966 vs :975-986). */
AI void add_core(const jo_extended *p, const jo_fe *n_vpu, const jo_fe *n_vmu, const jo_fe *n_z,
                 const jo_fe *n_t2d, int sub, jo_extended *out) {
    jo_fe a, b, c, d, t;
    completed_t r;
    QSUB(&p->v, &p->u, &t);
    QMUL(&t, sub ? n_vpu : n_vmu, &a);
    QADD(&p->v, &p->u, &t);
    QMUL(&t, sub ? n_vmu : n_vpu, &b);
    QMUL(&p->t1, &p->t2, &c);
    QMUL(&c, n_t2d, &c);
    if (n_z) { QMUL(&p->z, n_z, &d); QADD(&d, &d, &d); }
    else QADD(&p->z, &p->z, &d);
    QSUB(&b, &a, &r.u);
    QADD(&b, &a, &r.v);
    if (sub) { QSUB(&d, &c, &r.z); QADD(&d, &c, &r.t); }
    else     { QADD(&d, &c, &r.z); QSUB(&d, &c, &r.t); }
    into_extended(&r, out);
}
void jo_ext_add_niels(const jo_extended *p, const jo_extended_niels *n, jo_extended *out) {
    add_core(p, &n->v_plus_u, &n->v_minus_u, &n->z, &n->t2d, 0, out);
}
void jo_ext_sub_niels(const jo_extended *p, const jo_extended_niels *n, jo_extended *out) {
    add_core(p, &n->v_plus_u, &n->v_minus_u, &n->z, &n->t2d, 1, out);
}
void jo_ext_add_affine_niels(const jo_extended *p, const jo_affine_niels *n, jo_extended *out) {
    add_core(p, &n->v_plus_u, &n->v_minus_u, NULL, &n->t2d, 0, out);
}
void jo_ext_sub_affine_niels(const jo_extended *p, const jo_affine_niels *n, jo_extended *out) {
    add_core(p, &n->v_plus_u, &n->v_minus_u, NULL, &n->t2d, 1, out);
}
void jo_ext_add(const jo_extended *p, const jo_extended *q, jo_extended *out) {
    jo_extended_niels n;
    jo_ext_to_niels(q, &n);
    jo_ext_add_niels(p, &n, out);
}
void jo_ext_sub(const jo_extended *p, const jo_extended *q, jo_extended *out) {
    jo_extended_niels n;
    jo_ext_to_niels(q, &n);
    jo_ext_sub_niels(p, &n, out);
}
/* src/lib.rs:356-379.  The reference selects between identity and the point in
 * constant time (:375); the select is value-equivalent to the branch-free mask below. */
static void niels_mul_bits(const jo_extended_niels *n, const uint8_t by[32], jo_extended *out) {
    jo_extended acc;
    jo_extended_niels id = {FQ.r, FQ.r, FQ.r, {{0, 0, 0, 0}}}, sel;
    jo_ext_identity(&acc);
    for (int bit = 251; bit >= 0; bit--) {
        uint64_t mask = (uint64_t)0 - ((by[bit >> 3] >> (bit & 7)) & 1);
        const uint64_t *x = (const uint64_t *)&id, *y = (const uint64_t *)n;
        uint64_t *s = (uint64_t *)&sel;
        for (int k = 0; k < 16; k++) s[k] = x[k] ^ (mask & (x[k] ^ y[k]));
        jo_ext_double(&acc, &acc);
        jo_ext_add_niels(&acc, &sel, &acc);
    }
    *out = acc;
}
void jo_ext_mul_bits(const jo_extended *p, const uint8_t by[32], jo_extended *out) {
    jo_extended_niels n;
    jo_ext_to_niels(p, &n); /* src/lib.rs:830-833 */
    niels_mul_bits(&n, by, out);
}
/* src/lib.rs:271-295 */
void jo_affine_niels_mul_bits(const jo_affine_niels *n, const uint8_t by[32], jo_extended *out) {
    jo_extended acc;
    jo_affine_niels id = {FQ.r, FQ.r, {{0, 0, 0, 0}}}, sel;
    jo_ext_identity(&acc);
    for (int bit = 251; bit >= 0; bit--) {
        uint64_t mask = (uint64_t)0 - ((by[bit >> 3] >> (bit & 7)) & 1);
        const uint64_t *x = (const uint64_t *)&id, *y = (const uint64_t *)n;
        uint64_t *s = (uint64_t *)&sel;
        for (int k = 0; k < 12; k++) s[k] = x[k] ^ (mask & (x[k] ^ y[k]));
        jo_ext_double(&acc, &acc);
        jo_ext_add_affine_niels(&acc, &sel, &acc);
    }
    *out = acc;
}
void jo_ext_mul_by_cofactor(const jo_extended *p, jo_extended *out) {
    jo_ext_double(p, out);
    jo_ext_double(out, out);
    jo_ext_double(out, out);
}
int jo_ext_is_identity(const jo_extended *p) { return f_is_zero(&p->u) && f_eq(&p->v, &p->z); }
int jo_ext_is_small_order(const jo_extended *p) {
    jo_extended t;
    jo_ext_double(p, &t);
    jo_ext_double(&t, &t);
    return f_is_zero(&t.u);
}
static const uint8_t FR_MODULUS_BYTES[32] = { /* src/lib.rs:73-76 */
    183, 44, 247, 214, 94, 14, 151, 208, 130, 16, 200, 204, 147, 32, 104, 166,
    0, 59, 52, 1, 1, 59, 103, 6, 169, 175, 51, 101, 234, 180, 125, 14};
int jo_ext_is_torsion_free(const jo_extended *p) {
    jo_extended t;
    jo_ext_mul_bits(p, FR_MODULUS_BYTES, &t);
    return jo_ext_is_identity(&t);
}
int jo_ext_eq(const jo_extended *p, const jo_extended *q) {
    jo_fe a, b, c, d;
    QMUL(&p->u, &q->z, &a);
    QMUL(&q->u, &p->z, &b);
    QMUL(&p->v, &q->z, &c);
    QMUL(&q->v, &p->z, &d);
    return f_eq(&a, &b) && f_eq(&c, &d);
}
int jo_affine_is_on_curve(const jo_affine *p) {
    CONSTS();
    jo_fe u2, v2, lhs, rhs;
    QSQR(&p->u, &u2);
    QSQR(&p->v, &v2);
    QSUB(&v2, &u2, &lhs);
    QMUL(&EDWARDS_D, &u2, &rhs);
    QMUL(&rhs, &v2, &rhs);
    QADD(&FQ.r, &rhs, &rhs);
    return f_eq(&lhs, &rhs);
}
void jo_affine_to_bytes(const jo_affine *p, uint8_t out[32]) {
    uint8_t ub[32];
    f_to_bytes(&FQ, &p->v, out);
    f_to_bytes(&FQ, &p->u, ub);
    out[31] |= (uint8_t)(ub[0] << 7);
}
/* Tail of from_bytes_inner / batch_from_bytes once 1/(1 + d v^2) is known
 * (src/lib.rs:514-533, :603-624). */
static int finish_decode(const jo_fe *v, const jo_fe *num, const jo_fe *inv_den, int sign,
                         int zip216, jo_affine *out) {
    jo_fe u2, u, un;
    uint8_t ub[32];
    QMUL(num, inv_den, &u2);
    if (!fq_sqrt(&u2, &u)) return 0;
    f_to_bytes(&FQ, &u, ub);
    int flip = (ub[0] ^ sign) & 1;
    f_neg(&FQ, &u, &un);
    out->u = flip ? un : u;
    out->v = *v;
    return !(zip216 && f_is_zero(&u) && flip);
}
int jo_affine_from_bytes(const uint8_t in[32], int zip216, jo_affine *out) {
    CONSTS();
    uint8_t b[32];
    memcpy(b, in, 32);
    int sign = b[31] >> 7;
    b[31] &= 0x7f;
    jo_fe v, v2, num, den, inv;
    if (!f_from_bytes(&FQ, b, &v)) return 0;
    QSQR(&v, &v2);
    QSUB(&v2, &FQ.r, &num);
    QMUL(&EDWARDS_D, &v2, &den);
    QADD(&FQ.r, &den, &den);
    if (!fq_invert(&den, &inv)) memset(&inv, 0, sizeof inv); /* unwrap_or(zero), :514 */
    return finish_decode(&v, &num, &inv, sign, zip216, out);
}

/* ---- point batches ---------------------------------------------------------------- */
/* [ext] ff 0.13.1 BatchInverter::invert_with_internal_scratch as called at
 * src/lib.rs:849 / :1086: running products with zeros skipped, one inversion. */
static void batch_invert_skip_zero(jo_fe *elems, jo_fe *scratch, size_t n) {
    jo_fe acc = FQ.r, tmp;
    for (size_t i = 0; i < n; i++) {
        scratch[i] = acc;
        if (!f_is_zero(&elems[i])) QMUL(&acc, &elems[i], &acc);
    }
    fq_invert(&acc, &acc);
    for (size_t i = n; i-- > 0;) {
        if (f_is_zero(&elems[i])) continue;
        QMUL(&scratch[i], &acc, &tmp);
        QMUL(&acc, &elems[i], &acc);
        elems[i] = tmp;
    }
}
void jo_batch_normalize(const jo_extended *p, jo_affine *out, size_t n) {
    jo_fe *z = malloc(n * sizeof *z + 1), *scratch = malloc(n * sizeof *scratch + 1);
    for (size_t i = 0; i < n; i++) z[i] = p[i].z;
    batch_invert_skip_zero(z, scratch, n);
    for (size_t i = 0; i < n; i++) {
        QMUL(&p[i].u, &z[i], &out[i].u);
        QMUL(&p[i].v, &z[i], &out[i].v);
    }
    free(z);
    free(scratch);
}
void jo_batch_to_bytes(const jo_affine *p, uint8_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) jo_affine_to_bytes(&p[i], out + 32 * i);
}
void jo_batch_from_bytes(const uint8_t *in, jo_affine *out, uint8_t *ok, size_t n) {
    CONSTS();
    jo_fe *v = malloc(n * sizeof *v + 1), *num = malloc(n * sizeof *num + 1);
    jo_fe *den = malloc(n * sizeof *den + 1), *scratch = malloc(n * sizeof *scratch + 1);
    uint8_t *sign = malloc(n + 1), *some = malloc(n + 1);
    for (size_t i = 0; i < n; i++) {
        uint8_t b[32];
        jo_fe v2;
        memcpy(b, in + 32 * i, 32);
        sign[i] = b[31] >> 7;
        b[31] &= 0x7f;
        some[i] = (uint8_t)f_from_bytes(&FQ, b, &v[i]);
        QSQR(&v[i], &v2);
        QSUB(&v2, &FQ.r, &num[i]);
        QMUL(&EDWARDS_D, &v2, &den[i]);
        QADD(&FQ.r, &den[i], &den[i]);
        if (!some[i]) memset(&den[i], 0, sizeof den[i]); /* unwrap_or(zero), :598 */
    }
    batch_invert_skip_zero(den, scratch, n);
    for (size_t i = 0; i < n; i++) {
        memset(&out[i], 0, sizeof out[i]);
        ok[i] = some[i] ? (uint8_t)finish_decode(&v[i], &num[i], &den[i], sign[i], 1, &out[i]) : 0;
    }
    free(v); free(num); free(den); free(scratch); free(sign); free(some);
}
void jo_batch_double(const jo_extended *p, jo_extended *out, size_t n) {
    for (size_t i = 0; i < n; i++) jo_ext_double(&p[i], &out[i]);
}
void jo_batch_add(const jo_extended *p, const jo_extended *q, jo_extended *out, size_t n) {
    for (size_t i = 0; i < n; i++) jo_ext_add(&p[i], &q[i], &out[i]);
}
void jo_batch_add_niels(const jo_extended *p, const jo_extended_niels *q, jo_extended *out, size_t n) {
    for (size_t i = 0; i < n; i++) jo_ext_add_niels(&p[i], &q[i], &out[i]);
}
void jo_batch_add_affine_niels(const jo_extended *p, const jo_affine_niels *q, jo_extended *out, size_t n) {
    for (size_t i = 0; i < n; i++) jo_ext_add_affine_niels(&p[i], &q[i], &out[i]);
}

typedef struct {
    int kind; /* 0 var-base, 1 fixed-base, 2 torsion-free */
    const jo_extended *points;
    const jo_affine_niels *base;
    const uint8_t *scalars;
    jo_extended *out;
    uint8_t *flags;
    size_t lo, hi;
} shard_t;
static void *shard_main(void *arg) {
    shard_t *s = arg;
    for (size_t i = s->lo; i < s->hi; i++) {
        if (s->kind == 0) jo_ext_mul_bits(&s->points[i], s->scalars + 32 * i, &s->out[i]);
        else if (s->kind == 1) jo_affine_niels_mul_bits(s->base, s->scalars + 32 * i, &s->out[i]);
        else s->flags[i] = (uint8_t)jo_ext_is_torsion_free(&s->points[i]);
    }
    return NULL;
}
static void run_shards(shard_t proto, size_t n, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    pthread_t *th = malloc(sizeof *th * nthreads);
    shard_t *sh = malloc(sizeof *sh * nthreads);
    for (int t = 0; t < nthreads; t++) {
        sh[t] = proto;
        sh[t].lo = n * t / nthreads;
        sh[t].hi = n * (t + 1) / nthreads;
        if (t + 1 < nthreads) pthread_create(&th[t], NULL, shard_main, &sh[t]);
    }
    shard_main(&sh[nthreads - 1]);
    for (int t = 0; t + 1 < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    free(sh);
}
void jo_batch_scalar_mul(const jo_extended *points, const uint8_t *scalars32, jo_extended *out,
                         size_t n, int nthreads) {
    CONSTS();
    shard_t s = {0, points, NULL, scalars32, out, NULL, 0, 0};
    run_shards(s, n, nthreads);
}
void jo_batch_scalar_mul_fixed(const jo_affine *base, const uint8_t *scalars32, jo_extended *out,
                               size_t n, int nthreads) {
    CONSTS();
    jo_affine_niels bn;
    jo_affine_to_niels(base, &bn);
    shard_t s = {1, NULL, &bn, scalars32, out, NULL, 0, 0};
    run_shards(s, n, nthreads);
}
void jo_batch_is_torsion_free(const jo_extended *p, uint8_t *out, size_t n, int nthreads) {
    CONSTS();
    shard_t s = {2, p, NULL, NULL, NULL, out, 0, 0};
    run_shards(s, n, nthreads);
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
/* Dependent chain of n Montgomery muls (x <- x * y), like benches/fq_bench.rs:25-33. */
double jo_time_fe_mul(int which, size_t n, int reps) {
    double best = 1e30;
    jo_fe x, y;
    jo_fe_stream(which, 0x4A55424A55420001ULL, 0, 1, &x);
    jo_fe_stream(which, 0x4A55424A55420002ULL, 0, 1, &y);
    for (int r = 0; r < reps; r++) {
        double t0 = now_s();
        for (size_t i = 0; i < n; i++) jo_fe_mul(which, &x, &y, &x);
        double dt = now_s() - t0;
        if (dt < best) best = dt;
    }
    volatile uint64_t sink = x.l[0];
    (void)sink;
    return best;
}
double jo_time_scalar_mul(const jo_extended *points, const uint8_t *scalars32, jo_extended *out,
                          size_t n, int nthreads, int reps) {
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        double t0 = now_s();
        jo_batch_scalar_mul(points, scalars32, out, n, nthreads);
        double dt = now_s() - t0;
        if (dt < best) best = dt;
    }
    return best;
}
