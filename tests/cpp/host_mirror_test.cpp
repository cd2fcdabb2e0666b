// Exercises include/jubjub_b200.hpp on the GPU box (built by __graft_entry__.build(), run by
// tests/test_gpu_cpp_mirror.py).  Reads points/scalars/expected encodings from a binary file
// written by the Python test (expected values come from the oracle) and checks the C++ path.
#include <cstdio>
#include <cstring>
#include <fstream>

#include "../../include/jubjub_b200.hpp"

using namespace jubjub;

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::ifstream f(argv[1], std::ios::binary);
    uint64_t n = 0;
    f.read((char*)&n, 8);
    std::vector<ExtendedPoint> p(n);
    std::vector<Fr> k(n);
    std::vector<std::array<uint8_t, 32>> want(n);
    f.read((char*)p.data(), n * sizeof(ExtendedPoint));
    f.read((char*)k.data(), n * sizeof(Fr));
    f.read((char*)want.data(), n * 32);
    if (!f) return 3;
    try {
        Engine eng(0);
        auto prod = eng.batch_mul_vartime(p, k);                       // (p * k) element-wise
        auto enc = eng.batch_to_bytes(eng.batch_normalize(prod));
        for (uint64_t i = 0; i < n; i++)
            if (std::memcmp(enc[i].data(), want[i].data(), 32) != 0) {
                std::printf("mismatch at %llu\n", (unsigned long long)i);
                return 1;
            }
        // constant-time mode: the same points
        auto enc_ct = eng.batch_to_bytes(eng.batch_normalize(eng.batch_mul(p, k)));
        for (uint64_t i = 0; i < n; i++)
            if (std::memcmp(enc_ct[i].data(), want[i].data(), 32) != 0) return 12;
        // p + p == p.double() (projectively): compare through normalisation
        auto a = eng.batch_normalize(eng.batch_add(p, p));
        auto d = eng.batch_normalize(eng.batch_double(p));
        if (std::memcmp(a.data(), d.data(), n * sizeof(AffinePoint)) != 0) return 4;
        // wire format in and out: encode p, then decode + multiply + encode on the device
        std::vector<uint8_t> some;
        auto enc2 = eng.batch_mul_encoded_vartime(eng.batch_to_bytes(eng.batch_normalize(p)), k, some);
        for (uint64_t i = 0; i < n; i++)
            if (!some[i] || std::memcmp(enc2[i].data(), want[i].data(), 32) != 0) return 7;
        // in-place batch_normalize (src/lib.rs:1084-1107), mul_by_cofactor = three doublings, subgroup flags
        auto pn = p;
        auto an = eng.batch_normalize_in_place(pn);
        auto a0 = eng.batch_normalize(p);
        if (std::memcmp(an.data(), a0.data(), n * sizeof(AffinePoint)) != 0) return 8;
        for (uint64_t i = 0; i < n; i++)
            if (std::memcmp(&pn[i].t1, &pn[i].u, 32) != 0 || std::memcmp(&pn[i].t2, &pn[i].v, 32) != 0) return 9;
        auto c8 = eng.batch_mul_by_cofactor(p);
        auto d3 = eng.batch_double(eng.batch_double(eng.batch_double(p)));
        if (std::memcmp(c8.data(), d3.data(), n * sizeof(ExtendedPoint)) != 0) return 10;
        auto tf = eng.batch_is_torsion_free(c8), po = eng.batch_is_prime_order(c8);
        for (uint64_t i = 0; i < n; i++)
            if (!tf[i] || !po[i]) return 11;  // [8]P lies in the prime-order subgroup (and is not O for these inputs)
        // Sum: p0 + p1 + ... in groups of n/2 and as one sum; (sum of halves) == whole sum
        if (n >= 4 && n % 2 == 0) {
            auto halves = eng.batch_sum(p, n / 2);
            auto whole = eng.batch_normalize(eng.batch_sum(p));
            auto again = eng.batch_normalize(eng.batch_sum(halves));
            if (std::memcmp(whole.data(), again.data(), sizeof(AffinePoint)) != 0) return 13;
        }
        // ExtendedPoint::to_bytes (src/lib.rs:1419-1421) == AffinePoint::from(p).to_bytes()
        auto e1 = eng.batch_to_bytes(p);
        auto e2 = eng.batch_to_bytes(a0);
        for (uint64_t i = 0; i < n; i++)
            if (std::memcmp(e1[i].data(), e2[i].data(), 32) != 0) return 16;
        // Neg / PartialEq / From<AffinePoint> (src/lib.rs:153-226): p == from_affine(normalize(p)), -(-p) == p bit for bit,
        // p + (-p) == identity == its own negation
        auto back = eng.batch_from_affine(a0);
        auto eq1 = eng.batch_eq(p, back);
        auto nn = eng.batch_neg(eng.batch_neg(p));
        if (std::memcmp(nn.data(), p.data(), n * sizeof(ExtendedPoint)) != 0) return 14;
        auto zero = eng.batch_add(p, eng.batch_neg(p));
        auto eq2 = eng.batch_eq(zero, eng.batch_neg(zero)), eq3 = eng.batch_eq(p, eng.batch_double(p));
        for (uint64_t i = 0; i < n; i++)
            if (!eq1[i] || !eq2[i] || eq3[i]) return 15;
        bool threw = false;
        try {
            k.pop_back();
            eng.batch_mul_vartime(p, k);
        } catch (const Error& e) {
            threw = e.code == JJ_ERR_INVALID_ARG;
        }
        if (!threw) return 5;
    } catch (const Error& e) {
        std::printf("error %d: %s\n", e.code, e.what());
        return 6;
    }
    std::printf("cpp mirror ok: %llu scalar-muls\n", (unsigned long long)n);
    return 0;
}
