// jubjub_b200.hpp -- C++ host mirror of the reference's batch-facing surface over the C ABI.
//
// The reference is compiled code (Rust) whose toolchain is absent from this image, so the host
// layer above include/jubjub_b200.h is provided in C++ (and Python, jubjub_b200/engine.py).
// Names follow the reference: jubjub::{Fq, Fr, AffinePoint, ExtendedPoint, ExtendedNielsPoint,
// AffineNielsPoint} (src/lib.rs:62-63, 81-84, 139-145, 255-259, 327-332) are plain-old-data
// structs with the reference's field order; the new batch entry points (`batch_mul`,
// `batch_mul_fixed`, `batch_add`, `batch_double`, `batch_normalize`, `Fq::batch_mul` ...) call
// the kernels.  Length mismatches throw (the reference panics, src/lib.rs:841); per-element
// failures come back as flag vectors (the reference's CtOption).
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "jubjub_b200.h"

namespace jubjub {

struct Fq { std::array<uint64_t, 4> limbs; };  // Montgomery form, as bls12_381::Scalar holds it
struct Fr { std::array<uint64_t, 4> limbs; };  // Montgomery form, src/fr.rs:23
struct AffinePoint { Fq u, v; };
struct ExtendedPoint { Fq u, v, z, t1, t2; };
struct AffineNielsPoint { Fq v_plus_u, v_minus_u, t2d; };
struct ExtendedNielsPoint { Fq v_plus_u, v_minus_u, z, t2d; };
static_assert(sizeof(ExtendedPoint) == 160 && sizeof(AffinePoint) == 64 && sizeof(ExtendedNielsPoint) == 128 &&
                  sizeof(AffineNielsPoint) == 96 && sizeof(Fr) == 32,
              "layouts must match include/jubjub_b200.h");

class Error : public std::runtime_error {
   public:
    Error(int32_t code, const std::string& what) : std::runtime_error(what), code(code) {}
    int32_t code;
};

class Engine {
   public:
    explicit Engine(int device = 0) {
        int32_t rc = jj_init(device, &ctx_);
        if (rc != JJ_OK) throw Error(rc, "jj_init failed: a B200 (sm_100) device is required; there is no CPU fallback");
    }
    ~Engine() { if (ctx_) jj_destroy(ctx_); }
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;
    jj_ctx* raw() { return ctx_; }

    // ---- Fq / Fr batches (operator* / + / - / square / double / neg / invert of the reference, element-wise)
    std::vector<Fq> batch_mul(const std::vector<Fq>& a, const std::vector<Fq>& b) { return bin(jj_fq_mul, a, b); }
    std::vector<Fq> batch_add(const std::vector<Fq>& a, const std::vector<Fq>& b) { return bin(jj_fq_add, a, b); }
    std::vector<Fq> batch_sub(const std::vector<Fq>& a, const std::vector<Fq>& b) { return bin(jj_fq_sub, a, b); }
    std::vector<Fq> batch_square(const std::vector<Fq>& a) { return un(jj_fq_square, a); }
    std::vector<Fr> batch_mul(const std::vector<Fr>& a, const std::vector<Fr>& b) { return bin(jj_fr_mul, a, b); }
    std::vector<Fr> batch_add(const std::vector<Fr>& a, const std::vector<Fr>& b) { return bin(jj_fr_add, a, b); }
    std::vector<Fr> batch_sub(const std::vector<Fr>& a, const std::vector<Fr>& b) { return bin(jj_fr_sub, a, b); }
    std::vector<Fr> batch_square(const std::vector<Fr>& a) { return un(jj_fr_square, a); }
    // invert: is_some[i] == 0 where a[i] == 0 (CtOption::none, src/fr.rs:539)
    std::vector<Fq> batch_invert(const std::vector<Fq>& a, std::vector<uint8_t>& is_some) {
        std::vector<Fq> out(a.size());
        is_some.assign(a.size(), 0);
        check(jj_fq_invert(ctx_, a.data(), out.data(), is_some.data(), a.size(), 0));
        return out;
    }

    // ---- points
    // The scalar multiplications are variable-time in the scalar (the reference's `*` is constant-time by policy,
    // src/lib.rs:12-17) and are named *_vartime after the reference's rule (src/lib.rs:14-15): public scalars only.
    // `&ExtendedPoint * &Fr` element-wise (src/lib.rs:873-879)
    std::vector<ExtendedPoint> batch_mul_vartime(const std::vector<ExtendedPoint>& p, const std::vector<Fr>& k) {
        same(p.size(), k.size());
        std::vector<ExtendedPoint> out(p.size());
        check(jj_scalar_mul(ctx_, p.data(), k.data(), out.data(), p.size(), JJ_SCALAR_MONT));
        return out;
    }
    // the same in the engine's constant-time-in-the-scalar mode (JJ_CONST_TIME): no branch or address depends on k
    std::vector<ExtendedPoint> batch_mul(const std::vector<ExtendedPoint>& p, const std::vector<Fr>& k) {
        same(p.size(), k.size());
        std::vector<ExtendedPoint> out(p.size());
        check(jj_scalar_mul(ctx_, p.data(), k.data(), out.data(), p.size(), JJ_SCALAR_MONT | JJ_CONST_TIME));
        return out;
    }
    // `&AffinePoint * &Fr` for one shared base (src/lib.rs:1109-1115)
    std::vector<ExtendedPoint> batch_mul_fixed_vartime(const AffinePoint& base, const std::vector<Fr>& k) {
        std::vector<ExtendedPoint> out(k.size());
        check(jj_scalar_mul_fixed(ctx_, &base, k.data(), out.data(), k.size(), JJ_SCALAR_MONT));
        return out;
    }
    // `&ExtendedPoint + &ExtendedPoint` / `-` element-wise (src/lib.rs:992-1008)
    std::vector<ExtendedPoint> batch_add(const std::vector<ExtendedPoint>& p, const std::vector<ExtendedPoint>& q,
                                         bool subtract = false) {
        same(p.size(), q.size());
        std::vector<ExtendedPoint> out(p.size());
        check(jj_point_add(ctx_, p.data(), q.data(), out.data(), p.size(), subtract ? JJ_SUBTRACT : 0));
        return out;
    }
    std::vector<ExtendedPoint> batch_add(const std::vector<ExtendedPoint>& p, const std::vector<ExtendedNielsPoint>& q,
                                         bool subtract = false) {
        same(p.size(), q.size());
        std::vector<ExtendedPoint> out(p.size());
        check(jj_point_add_niels(ctx_, p.data(), q.data(), out.data(), p.size(), subtract ? JJ_SUBTRACT : 0));
        return out;
    }
    // Sum<ExtendedPoint> (src/lib.rs:183-193) of consecutive groups of `group_size` points (0 = the whole batch)
    std::vector<ExtendedPoint> batch_sum(const std::vector<ExtendedPoint>& p, size_t group_size = 0) {
        const size_t g = group_size ? group_size : p.size();
        if (g && p.size() % g) throw Error(JJ_ERR_INVALID_ARG, "batch does not split into groups of that size");
        const size_t groups = g ? p.size() / g : 1;
        std::vector<ExtendedPoint> out(groups);
        check(jj_point_sum(ctx_, p.data(), out.data(), groups, g, 0));
        return out;
    }
    std::vector<ExtendedPoint> batch_double(const std::vector<ExtendedPoint>& p) {
        std::vector<ExtendedPoint> out(p.size());
        check(jj_point_double(ctx_, p.data(), out.data(), p.size(), 0));
        return out;
    }
    // batch_normalize (src/lib.rs:1084-1107)
    std::vector<AffinePoint> batch_normalize(const std::vector<ExtendedPoint>& p) {
        std::vector<AffinePoint> out(p.size());
        check(jj_batch_normalize(ctx_, p.data(), out.data(), p.size(), 0));
        return out;
    }
    // GroupEncoding::to_bytes for ExtendedPoint (src/lib.rs:1419-1421): normalise + encode in one pass
    std::vector<std::array<uint8_t, 32>> batch_to_bytes(const std::vector<ExtendedPoint>& p) {
        std::vector<std::array<uint8_t, 32>> out(p.size());
        check(jj_batch_normalize(ctx_, p.data(), out.data(), p.size(), JJ_OUT_BYTES));
        return out;
    }
    // AffinePoint::to_bytes (src/lib.rs:455-464)
    std::vector<std::array<uint8_t, 32>> batch_to_bytes(const std::vector<AffinePoint>& p) {
        std::vector<std::array<uint8_t, 32>> out(p.size());
        check(jj_affine_to_bytes(ctx_, p.data(), out.data(), p.size(), 0));
        return out;
    }
    // AffinePoint::batch_from_bytes (src/lib.rs:541-627): is_some[i] == 0 for a rejected encoding
    std::vector<AffinePoint> batch_from_bytes(const std::vector<std::array<uint8_t, 32>>& enc, std::vector<uint8_t>& is_some) {
        std::vector<AffinePoint> out(enc.size());
        is_some.assign(enc.size(), 0);
        check(jj_batch_from_bytes(ctx_, enc.data(), out.data(), is_some.data(), enc.size(), 0));
        return out;
    }
    // wire format in and out: AffinePoint::from_bytes(enc[i]) * k[i], encoded (decode + scalar-mul + encode on the device)
    std::vector<std::array<uint8_t, 32>> batch_mul_encoded_vartime(const std::vector<std::array<uint8_t, 32>>& enc,
                                                           const std::vector<Fr>& k, std::vector<uint8_t>& is_some) {
        same(enc.size(), k.size());
        std::vector<std::array<uint8_t, 32>> out(enc.size());
        is_some.assign(enc.size(), 0);
        check(jj_scalar_mul_encoded(ctx_, enc.data(), k.data(), out.data(), is_some.data(), enc.size(),
                                    JJ_SCALAR_MONT | JJ_OUT_BYTES));
        return out;
    }
    // is_torsion_free / is_prime_order (src/lib.rs:709-719), mul_by_cofactor (:722-724)
    std::vector<uint8_t> batch_is_torsion_free(const std::vector<ExtendedPoint>& p) {
        std::vector<uint8_t> out(p.size());
        check(jj_is_torsion_free(ctx_, p.data(), out.data(), p.size(), 0));
        return out;
    }
    std::vector<uint8_t> batch_is_prime_order(const std::vector<ExtendedPoint>& p) {
        std::vector<uint8_t> out(p.size());
        check(jj_is_prime_order(ctx_, p.data(), out.data(), p.size(), 0));
        return out;
    }
    // Neg / PartialEq for ExtendedPoint (src/lib.rs:195-210, 153-181), From<AffinePoint> (:214-226), element by element
    std::vector<ExtendedPoint> batch_neg(const std::vector<ExtendedPoint>& p) {
        std::vector<ExtendedPoint> out(p.size());
        check(jj_point_neg(ctx_, p.data(), out.data(), p.size(), 0));
        return out;
    }
    std::vector<uint8_t> batch_eq(const std::vector<ExtendedPoint>& p, const std::vector<ExtendedPoint>& q) {
        same(p.size(), q.size());
        std::vector<uint8_t> out(p.size());
        check(jj_point_eq(ctx_, p.data(), q.data(), out.data(), p.size(), 0));
        return out;
    }
    std::vector<ExtendedPoint> batch_from_affine(const std::vector<AffinePoint>& a) {
        std::vector<ExtendedPoint> out(a.size());
        check(jj_affine_to_extended(ctx_, a.data(), out.data(), a.size(), 0));
        return out;
    }
    std::vector<ExtendedPoint> batch_mul_by_cofactor(const std::vector<ExtendedPoint>& p) {
        std::vector<ExtendedPoint> out(p.size());
        check(jj_mul_by_cofactor(ctx_, p.data(), out.data(), p.size(), 0));
        return out;
    }
    // the free function batch_normalize (src/lib.rs:1084-1107): normalises `p` in place (z = 1, t1 = u, t2 = v)
    // and returns the affine points
    std::vector<AffinePoint> batch_normalize_in_place(std::vector<ExtendedPoint>& p) {
        check(jj_batch_normalize_extended(ctx_, p.data(), p.data(), p.size(), 0));
        std::vector<AffinePoint> out(p.size());
        for (size_t i = 0; i < p.size(); i++) out[i] = AffinePoint{p[i].u, p[i].v};
        return out;
    }

   private:
    template <class Fn, class T>
    std::vector<T> bin(Fn fn, const std::vector<T>& a, const std::vector<T>& b) {
        same(a.size(), b.size());
        std::vector<T> out(a.size());
        check(fn(ctx_, a.data(), b.data(), out.data(), a.size(), 0));
        return out;
    }
    template <class Fn, class T>
    std::vector<T> un(Fn fn, const std::vector<T>& a) {
        std::vector<T> out(a.size());
        check(fn(ctx_, a.data(), out.data(), a.size(), 0));
        return out;
    }
    static void same(size_t a, size_t b) {
        if (a != b) throw Error(JJ_ERR_INVALID_ARG, "length mismatch (the reference panics: assert_eq!, src/lib.rs:841)");
    }
    void check(int32_t rc) {
        if (rc != JJ_OK) throw Error(rc, jj_last_error(ctx_));
    }
    jj_ctx* ctx_ = nullptr;
};

}  // namespace jubjub
