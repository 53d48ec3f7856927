// host_mirror_check.cpp -- compiles the C++ mirror against the C ABI and checks the error mapping on a
// box without a GPU (validation errors precede the device check; compute refuses with CudaError).
#include <cstdio>

#include "../../toy-heaan-ckks_b200/host/rns_poly.hpp"

static int expect(const char *what, int want, uint64_t n, std::vector<uint64_t> moduli) {
    try {
        ckks::RnsBasis::create(n, moduli);
    } catch (const ckks::RnsNttError &e) {
        if (e.code == want) return 0;
        std::printf("%s: got %d (%s), wanted %d\n", what, e.code, e.what(), want);
        return 1;
    }
    if (want == CKKS_OK) return 0;
    std::printf("%s: no error, wanted %d\n", what, want);
    return 1;
}
int main() {
    int bad = 0;
    bad += expect("empty", CKKS_EMPTY_BASIS, 8, {});
    bad += expect("non-friendly", CKKS_NON_NTT_FRIENDLY_MODULUS, 8, {19});
    bad += expect("degree", CKKS_INVALID_DEGREE, 12, {17});
    auto p = ckks::generate_primes(31, 4, 16);
    bad += !(p.size() == 4 && p[0] == 2147483489ull && p[3] == 2147482273ull);
    if (ckks_device_count() == 0) bad += expect("no gpu", CKKS_CUDA_ERROR, 8, {17, 97, 113});
    else {
        auto b = ckks::RnsBasis::create(8, {17, 97, 113});
        auto x = ckks::RnsPoly::from_coeffs({1, 1, 0, 0, 0, 0, 0, 0}, 1, b);
        auto y = x;
        x *= y;  // (1+x)^2 = 1 + 2x + x^2, poly.rs:789-802
        auto ch = x.channels();
        bad += !(ch[0] == 1 && ch[1] == 2 && ch[2] == 1 && ch[3] == 0);
    }
    std::printf(bad ? "FAIL\n" : "ok\n");
    return bad;
}
