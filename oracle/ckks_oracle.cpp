// ckks_oracle.cpp -- CPU ORACLE (test infrastructure, NOT the product).  See ckks_oracle.h.
//
// Restates the reference algorithm and schedule; every block cites the reference file:line.
// Deliberately slow in the same places the reference is slow (u128 %, table lookups with stride,
// re-transforming operands on every multiply) because it doubles as the honest CPU baseline.
#include "ckks_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstring>
#include <thread>
#include <vector>

typedef unsigned __int128 u128;
typedef __int128 i128;
typedef uint64_t u64;
typedef int64_t i64;

// ---------------------------------------------------------------------------------------------
// modular helpers -- poly.rs:629-653, primes.rs:24-45
// ---------------------------------------------------------------------------------------------
static inline u64 mul_mod(u64 a, u64 b, u64 q) { return (u64)(((u128)a * (u128)b) % (u128)q); }
static inline u64 add_mod(u64 a, u64 b, u64 q) {
    u64 s = a + b;  // q < 2^63 in every reference configuration, so no wrap
    return s >= q ? s - q : s;
}
static inline u64 sub_mod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
static u64 mod_pow(u64 base, u64 e, u64 q) {
    if (q == 1) return 0;
    u64 acc = 1 % q;
    base %= q;
    while (e) {
        if (e & 1) acc = mul_mod(acc, base, q);
        base = mul_mod(base, base, q);
        e >>= 1;
    }
    return acc;
}

// basis.rs:198-210 -- extended Euclid over i128, result normalised into [0, m)
static void egcd(i128 a, i128 b, i128 &g, i128 &x, i128 &y) {
    if (a == 0) {
        g = b; x = 0; y = 1;
        return;
    }
    i128 g1, x1, y1;
    egcd(b % a, a, g1, x1, y1);
    g = g1;
    x = y1 - (b / a) * x1;
    y = x1;
}
static u64 mod_inverse(u64 v, u64 m) {
    i128 g, x, y;
    egcd((i128)v, (i128)m, g, x, y);
    i128 mm = (i128)m;
    return (u64)(((x % mm) + mm) % mm);
}

// ---------------------------------------------------------------------------------------------
// primes -- src/math/primes.rs
// ---------------------------------------------------------------------------------------------
static const u64 MR_BASES[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};  // primes.rs:21

extern "C" int orc_is_prime(u64 n) {  // primes.rs:67-93
    if (n < 2) return 0;
    if (n == 2 || n == 3) return 1;
    if ((n & 1) == 0) return 0;
    u64 d = n - 1;
    unsigned r = 0;
    while ((d & 1) == 0) { d >>= 1; ++r; }
    for (u64 a : MR_BASES) {
        if (a >= n) continue;
        u64 x = mod_pow(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool witness = true;
        for (unsigned i = 1; i < r; ++i) {
            x = mul_mod(x, x, n);
            if (x == n - 1) { witness = false; break; }
        }
        if (witness) return 0;
    }
    return 1;
}

extern "C" int orc_is_ntt_friendly_prime(u64 p, u64 n) {  // primes.rs:125-131
    u64 m = 2 * n;
    return orc_is_prime(p) && (p % m == 1);
}

extern "C" u64 orc_get_first_prime_up(uint32_t logq, u64 n) {  // primes.rs:171-187 (+ snap_up :134-148)
    u64 step = 2 * n;
    u64 value = ((u64)1 << logq) + 1;
    u64 rem = value % step;
    u64 cand = (rem == 1) ? value : value + (step + 1 - rem) % step;
    for (;;) {
        if (orc_is_prime(cand)) return cand;
        cand += step;
    }
}

extern "C" u64 orc_get_first_prime_down(u64 bound, u64 n) {  // primes.rs:198-219 (+ snap_down :151-161)
    if (bound <= 2) return 0;
    u64 step = 2 * n;
    u64 value = bound - 1;
    u64 rem = value % step;
    u64 delta = (rem + step - 1) % step;
    if (delta > value) return 0;
    u64 cand = value - delta;
    for (;;) {
        if (cand <= 2) return 0;
        if (orc_is_prime(cand)) return cand;
        if (cand < step) return 0;
        cand -= step;
    }
}

extern "C" int orc_generate_primes(int bit_size, int count, u64 degree, u64 *out) {  // utils.rs:47-80
    if (bit_size < 4 || bit_size > 63 || count <= 0 || degree == 0) return ORC_PANIC;
    u64 upper = ((u64)1 << bit_size) - 1;
    u64 lower = (u64)1 << (bit_size - 1);
    u64 cursor = orc_get_first_prime_down(upper + 1, degree);
    if (cursor == 0) return ORC_PANIC;
    int found = 0;
    while (found < count) {
        if (cursor < lower) break;
        out[found++] = cursor;
        cursor = orc_get_first_prime_down(cursor, degree);
        if (cursor == 0) break;
    }
    return found == count ? ORC_OK : ORC_PANIC;
}

// ---------------------------------------------------------------------------------------------
// basis -- src/rings/backends/rns_ntt/basis.rs
// ---------------------------------------------------------------------------------------------
struct NttTable {  // basis.rs:6-17
    std::vector<u64> forward_roots, inverse_roots, twist, untwist;
    u64 n_inv, modulus, psi;
};
struct orc_basis {  // basis.rs:91-94
    u64 n;
    std::vector<u64> moduli;
    std::vector<NttTable> tables;
};

static u64 find_primitive_root(u64 q, u64 order) {  // basis.rs:217-237 (+ distinct_prime_factors :239-255)
    u64 exponent = (q - 1) / order;
    std::vector<u64> factors;
    u64 v = order;
    for (u64 d = 2; d * d <= v; ++d) {
        if (v % d == 0) {
            factors.push_back(d);
            while (v % d == 0) v /= d;
        }
    }
    if (v > 1) factors.push_back(v);
    for (u64 cand = 2; cand < q; ++cand) {
        u64 root = mod_pow(cand, exponent, q);
        if (root == 1) continue;
        bool ok = true;
        for (u64 f : factors)
            if (mod_pow(root, order / f, q) == 1) { ok = false; break; }
        if (ok) return root;
    }
    return 0;
}

static int build_table(u64 n, u64 q, NttTable &t) {  // basis.rs:21-84
    if (n == 0 || (n & (n - 1)) != 0) return ORC_INVALID_DEGREE;
    if (!orc_is_ntt_friendly_prime(q, n)) return ORC_NON_NTT_FRIENDLY_MODULUS;
    u64 psi = find_primitive_root(q, 2 * n);
    u64 omega = mod_pow(psi, 2, q);
    u64 omega_inv = mod_inverse(omega, q);
    u64 psi_inv = mod_inverse(psi, q);
    t.modulus = q;
    t.psi = psi;
    t.forward_roots.assign(n, 1);
    t.inverse_roots.assign(n, 1);
    t.twist.assign(n, 1);
    t.untwist.assign(n, 1);
    // The reference fills each entry with an independent mod_pow; the running product below
    // yields the same values (exact arithmetic) without the O(N log q) table build.
    for (u64 i = 1; i < n; ++i) {
        t.forward_roots[i] = mul_mod(t.forward_roots[i - 1], omega, q);
        t.inverse_roots[i] = mul_mod(t.inverse_roots[i - 1], omega_inv, q);
        t.twist[i] = mul_mod(t.twist[i - 1], psi, q);
        t.untwist[i] = mul_mod(t.untwist[i - 1], psi_inv, q);
    }
    t.n_inv = mod_inverse(n % q, q);
    return ORC_OK;
}

extern "C" int orc_basis_new(u64 n, const u64 *moduli, size_t l, orc_basis **out) {  // basis.rs:97-106
    *out = nullptr;
    if (l == 0) return ORC_EMPTY_BASIS;
    orc_basis *b = new orc_basis();
    b->n = n;
    b->moduli.assign(moduli, moduli + l);
    b->tables.resize(l);
    for (size_t i = 0; i < l; ++i) {
        int rc = build_table(n, moduli[i], b->tables[i]);
        if (rc != ORC_OK) { delete b; return rc; }
    }
    *out = b;
    return ORC_OK;
}
extern "C" void orc_basis_free(orc_basis *b) { delete b; }

extern "C" int orc_basis_drop_last(const orc_basis *b, size_t drop, orc_basis **out) {  // basis.rs:121-134
    *out = nullptr;
    size_t cc = b->moduli.size();
    if (drop >= cc) return ORC_INVALID_MOD_DROP;
    size_t keep = cc - drop;
    orc_basis *r = new orc_basis();
    r->n = b->n;
    r->moduli.assign(b->moduli.begin(), b->moduli.begin() + keep);
    r->tables.assign(b->tables.begin(), b->tables.begin() + keep);  // deep copy, like the reference
    *out = r;
    return ORC_OK;
}
extern "C" u64 orc_basis_degree(const orc_basis *b) { return b->n; }
extern "C" size_t orc_basis_channel_count(const orc_basis *b) { return b->moduli.size(); }
extern "C" void orc_basis_moduli(const orc_basis *b, u64 *out) {
    std::copy(b->moduli.begin(), b->moduli.end(), out);
}
extern "C" uint32_t orc_basis_total_bits(const orc_basis *b) {  // basis.rs:140-145
    uint32_t s = 0;
    for (u64 q : b->moduli) s += 63 - (uint32_t)__builtin_clzll(q);
    return s;
}
extern "C" u64 orc_basis_psi(const orc_basis *b, size_t ch) { return b->tables[ch].psi; }
extern "C" void orc_basis_table(const orc_basis *b, size_t ch, int which, u64 *out) {
    const NttTable &t = b->tables[ch];
    const std::vector<u64> *v = nullptr;
    switch (which) {
        case 0: v = &t.forward_roots; break;
        case 1: v = &t.inverse_roots; break;
        case 2: v = &t.twist; break;
        case 3: v = &t.untwist; break;
        default: out[0] = t.n_inv; return;
    }
    std::copy(v->begin(), v->end(), out);
}

extern "C" i64 orc_reconstruct_centered_coeff(const orc_basis *b, const u64 *res) {  // basis.rs:158-180
    u128 q = 1;
    for (u64 m : b->moduli) q *= (u128)m;  // wraps exactly where the reference's u128 product would
    u128 acc = 0;
    for (size_t i = 0; i < b->moduli.size(); ++i) {
        u64 m = b->moduli[i];
        u128 qi = q / (u128)m;
        u64 qi_inv = mod_inverse((u64)(qi % (u128)m), m);
        u128 s = ((u128)res[i] * (u128)qi_inv) % (u128)m;
        u128 term = s * qi % q;
        acc = (acc + term) % q;
    }
    if (acc > q / 2) return (i64)((i128)acc - (i128)q);
    return (i64)acc;
}

// ---------------------------------------------------------------------------------------------
// CPU-baseline threading: the reference is single-threaded; the baseline drivers may spread the
// independent per-limb transforms of one polynomial over host threads (same arithmetic, same
// results).  limb_threads == 1 (the default, and what every parity test uses) is purely serial.
// ---------------------------------------------------------------------------------------------
static thread_local int limb_threads = 1;
template <class F>
static void for_limbs(size_t l, F fn) {
    int t = limb_threads;
    if (t <= 1 || l <= 1) {
        for (size_t c = 0; c < l; ++c) fn(c);
        return;
    }
    if ((size_t)t > l) t = (int)l;
    std::vector<std::thread> pool;
    for (int w = 0; w < t; ++w)
        pool.emplace_back([&, w]() {
            for (size_t c = (size_t)w; c < l; c += (size_t)t) fn(c);
        });
    for (auto &th : pool) th.join();
}

// ---------------------------------------------------------------------------------------------
// polynomial kernels -- poly.rs:574-625
// ---------------------------------------------------------------------------------------------
static void bit_reverse_permute(u64 *v, u64 n) {  // poly.rs:617-625, basis.rs:257-259
    unsigned bits = (unsigned)__builtin_ctzll(n);
    if (bits == 0) return;
    for (u64 i = 0; i < n; ++i) {
        u64 j = 0;
        for (unsigned k = 0; k < bits; ++k) j |= ((i >> k) & 1) << (bits - 1 - k);
        if (i < j) std::swap(v[i], v[j]);
    }
}
static void cooley_tukey_ntt(u64 *v, const u64 *roots, u64 q, u64 n) {  // poly.rs:593-615
    for (u64 len = 2; len <= n; len *= 2) {
        u64 half = len / 2, step = n / len;
        for (u64 start = 0; start < n; start += len) {
            for (u64 off = 0; off < half; ++off) {
                u64 left = start + off, right = left + half;
                u64 t = mul_mod(v[right], roots[off * step], q);
                u64 u = v[left];
                v[left] = add_mod(u, t, q);
                v[right] = sub_mod(u, t, q);
            }
        }
    }
}
static void forward_ntt(u64 *v, const NttTable &t, u64 n) {  // poly.rs:574-580
    bit_reverse_permute(v, n);
    cooley_tukey_ntt(v, t.forward_roots.data(), t.modulus, n);
}
static void inverse_ntt(u64 *v, const NttTable &t, u64 n) {  // poly.rs:582-591
    bit_reverse_permute(v, n);
    cooley_tukey_ntt(v, t.inverse_roots.data(), t.modulus, n);
    for (u64 i = 0; i < n; ++i) v[i] = mul_mod(v[i], t.n_inv, t.modulus);
}

extern "C" void orc_from_coeffs(const orc_basis *b, const i64 *coeffs, u64 *out) {  // poly.rs:49-66
    u64 n = b->n;
    for (size_t ch = 0; ch < b->moduli.size(); ++ch) {
        i128 q = (i128)b->moduli[ch];
        for (u64 i = 0; i < n; ++i) {
            i128 r = (i128)coeffs[i] % q;
            if (r < 0) r += q;  // rem_euclid
            out[ch * n + i] = (u64)r;
        }
    }
}
extern "C" int orc_from_channels_check(const orc_basis *b, const u64 *ch, size_t nch) {  // poly.rs:72-99
    if (nch != b->moduli.size()) return ORC_CHANNEL_COUNT_MISMATCH;
    u64 n = b->n;
    for (size_t c = 0; c < nch; ++c) {
        u64 q = b->moduli[c];
        for (u64 i = 0; i < n; ++i)
            if (ch[c * n + i] >= q) return ORC_NON_REDUCED_COEFFICIENT;
    }
    return ORC_OK;
}
extern "C" void orc_to_ntt_domain(const orc_basis *b, u64 *ch) {  // poly.rs:136-148
    u64 n = b->n;
    for_limbs(b->moduli.size(), [&](size_t c) {
        const NttTable &t = b->tables[c];
        u64 *v = ch + c * n;
        for (u64 j = 0; j < n; ++j) v[j] = mul_mod(v[j], t.twist[j], t.modulus);
        forward_ntt(v, t, n);
    });
}
extern "C" void orc_to_coeff_domain(const orc_basis *b, u64 *ch) {  // poly.rs:154-166
    u64 n = b->n;
    for_limbs(b->moduli.size(), [&](size_t c) {
        const NttTable &t = b->tables[c];
        u64 *v = ch + c * n;
        inverse_ntt(v, t, n);
        for (u64 j = 0; j < n; ++j) v[j] = mul_mod(v[j], t.untwist[j], t.modulus);
    });
}
extern "C" void orc_add_assign(const orc_basis *b, u64 *a, const u64 *rhs) {  // poly.rs:254-275
    u64 n = b->n;
    for (size_t c = 0; c < b->moduli.size(); ++c) {
        u64 q = b->moduli[c];
        for (u64 i = 0; i < n; ++i) a[c * n + i] = add_mod(a[c * n + i], rhs[c * n + i], q);
    }
}
extern "C" void orc_neg(const orc_basis *b, u64 *a) {  // poly.rs:370-385
    u64 n = b->n;
    for (size_t c = 0; c < b->moduli.size(); ++c) {
        u64 q = b->moduli[c];
        for (u64 i = 0; i < n; ++i)
            if (a[c * n + i] != 0) a[c * n + i] = q - a[c * n + i];
    }
}
extern "C" void orc_mul_assign(const orc_basis *b, u64 *a, const u64 *rhs, int in_ntt) {  // poly.rs:277-331
    u64 n = b->n;
    size_t l = b->moduli.size();
    if (in_ntt) {  // :297-306
        for (size_t c = 0; c < l; ++c) {
            u64 q = b->moduli[c];
            for (u64 i = 0; i < n; ++i) a[c * n + i] = mul_mod(a[c * n + i], rhs[c * n + i], q);
        }
        return;
    }
    orc_to_ntt_domain(b, a);                            // :310
    std::vector<u64> rhs_ntt(rhs, rhs + l * n);         // :312 (rhs cloned and re-transformed every time)
    for_limbs(l, [&](size_t c) {                        // :313-319
        const NttTable &t = b->tables[c];
        u64 *v = rhs_ntt.data() + c * n;
        for (u64 j = 0; j < n; ++j) v[j] = mul_mod(v[j], t.twist[j], t.modulus);
        forward_ntt(v, t, n);
    });
    for_limbs(l, [&](size_t c) {                        // :321-326
        u64 q = b->moduli[c];
        for (u64 i = 0; i < n; ++i) a[c * n + i] = mul_mod(a[c * n + i], rhs_ntt[c * n + i], q);
    });
    orc_to_coeff_domain(b, a);                          // :328
}
extern "C" void orc_mul_assign_naive(const orc_basis *b, u64 *a, const u64 *rhs) {  // poly.rs:339-367
    u64 n = b->n;
    std::vector<u64> res(n);
    for (size_t c = 0; c < b->moduli.size(); ++c) {
        u64 q = b->moduli[c];
        const u64 *x = a + c * n, *y = rhs + c * n;
        std::fill(res.begin(), res.end(), 0);
        for (u64 i = 0; i < n; ++i)
            for (u64 j = 0; j < n; ++j) {
                u64 p = mul_mod(x[i], y[j], q);
                if (i + j < n) res[i + j] = add_mod(res[i + j], p, q);
                else res[i + j - n] = sub_mod(res[i + j - n], p, q);  // X^N = -1
            }
        std::copy(res.begin(), res.end(), a + c * n);
    }
}
extern "C" int orc_rescale(const orc_basis *b, const u64 *ch, int in_ntt, u64 *out) {  // poly.rs:187-228
    size_t l = b->moduli.size();
    u64 n = b->n;
    if (l < 2) return ORC_INVALID_MOD_DROP;
    std::vector<u64> tmp;
    const u64 *src = ch;
    if (in_ntt) {  // :199-208 clone + to_coeff_domain
        tmp.assign(ch, ch + l * n);
        orc_to_coeff_domain(b, tmp.data());
        src = tmp.data();
    }
    size_t last = l - 1;
    u64 q_last = b->moduli[last];
    for (size_t i = 0; i < last; ++i) {  // :214-225
        u64 qi = b->moduli[i];
        u64 q_last_inv = mod_inverse(q_last % qi, qi);
        for (u64 j = 0; j < n; ++j) {
            u64 ci = src[i * n + j];
            u64 cl = src[last * n + j] % qi;
            out[i * n + j] = mul_mod(sub_mod(ci, cl, qi), q_last_inv, qi);
        }
    }
    return ORC_OK;
}
extern "C" int orc_automorphism(const orc_basis *b, const u64 *ch, int in_ntt, u64 exponent, u64 *out) {
    // poly.rs:492-541
    size_t l = b->moduli.size();
    u64 n = b->n;
    u64 two_n = 2 * n;
    u64 e = exponent % two_n;
    if (e == 0) {  // :508-511 quirk: returns self.clone() -- the domain flag is preserved
        std::memcpy(out, ch, l * n * sizeof(u64));
        return in_ntt;
    }
    std::vector<u64> tmp;
    const u64 *src = ch;
    if (in_ntt) {  // :494-503
        tmp.assign(ch, ch + l * n);
        orc_to_coeff_domain(b, tmp.data());
        src = tmp.data();
    }
    std::memset(out, 0, l * n * sizeof(u64));
    for (size_t c = 0; c < l; ++c) {  // :515-538 scatter in increasing i (last writer wins for even e)
        u64 q = b->moduli[c];
        for (u64 i = 0; i < n; ++i) {
            u64 coeff = src[c * n + i];
            u64 j_full = (i * e) % two_n;
            u64 j = j_full % n;
            if (coeff == 0) continue;
            out[c * n + j] = (j_full >= n) ? q - coeff : coeff;
        }
    }
    return 0;
}
extern "C" int orc_rotate_slots(const orc_basis *b, const u64 *ch, int in_ntt, int32_t k, u64 *out) {
    // poly.rs:546-569
    u64 two_n = 2 * b->n;
    u64 rot = k >= 0 ? (u64)k : (u64)(-(i64)k);
    u64 e = mod_pow(5, rot, two_n);
    if (k >= 0) return orc_automorphism(b, ch, in_ntt, e, out);
    std::vector<u64> mid(b->moduli.size() * b->n);
    int d = orc_automorphism(b, ch, in_ntt, e, mid.data());
    return orc_automorphism(b, mid.data(), d, two_n - 1, out);
}
extern "C" void orc_to_coeffs(const orc_basis *b, const u64 *ch, int in_ntt, i64 *out) {  // poly.rs:404-427
    size_t l = b->moduli.size();
    u64 n = b->n;
    std::vector<u64> tmp;
    const u64 *src = ch;
    if (in_ntt) {
        tmp.assign(ch, ch + l * n);
        orc_to_coeff_domain(b, tmp.data());
        src = tmp.data();
    }
    std::vector<u64> res(l);
    for (u64 i = 0; i < n; ++i) {
        for (size_t c = 0; c < l; ++c) res[c] = src[c * n + i];
        out[i] = orc_reconstruct_centered_coeff(b, res.data());
    }
}

// ---------------------------------------------------------------------------------------------
// engine -- src/crypto/engine.rs
// ---------------------------------------------------------------------------------------------
extern "C" void orc_encrypt(const orc_basis *b, const u64 *pk_b, const u64 *pk_a, const u64 *u,
                            const u64 *e0, const u64 *e1, const u64 *m, u64 *c0, u64 *c1) {  // :84-112
    size_t w = b->moduli.size() * b->n;
    std::memcpy(c0, pk_b, w * 8);
    orc_mul_assign(b, c0, u, 0);
    orc_add_assign(b, c0, e0);
    orc_add_assign(b, c0, m);
    std::memcpy(c1, pk_a, w * 8);
    orc_mul_assign(b, c1, u, 0);
    orc_add_assign(b, c1, e1);
}
extern "C" void orc_decrypt(const orc_basis *b, const u64 *c0, const u64 *c1, const u64 *s, u64 *out) {  // :114-128
    size_t w = b->moduli.size() * b->n;
    std::memcpy(out, c1, w * 8);
    orc_mul_assign(b, out, s, 0);
    orc_add_assign(b, out, c0);
}
extern "C" void orc_add_ciphertexts(const orc_basis *b, const u64 *a0, const u64 *a1, const u64 *b0,
                                    const u64 *b1, u64 *c0, u64 *c1) {  // :131-151
    size_t w = b->moduli.size() * b->n;
    std::memcpy(c0, a0, w * 8);
    orc_add_assign(b, c0, b0);
    std::memcpy(c1, a1, w * 8);
    orc_add_assign(b, c1, b1);
}

// The gadget loop shared by mul_ciphertexts_gadget (:505-528) and rotate_ciphertext (:429-452):
// for each digit i, alpha_i = limb i of `src` broadcast to every limb j with `% q_j`, then
// acc0 += alpha_i * key_b[i], acc1 += alpha_i * key_a[i] -- each `*` a full coefficient-domain
// multiply (2 forward + 1 inverse NTT per limb), exactly as the reference does it.
static void gadget_accumulate(const orc_basis *b, const u64 *src, const u64 *key_a, const u64 *key_b,
                              u64 *acc0, u64 *acc1) {
    size_t l = b->moduli.size();
    u64 n = b->n;
    size_t w = l * n;
    std::fill(acc0, acc0 + w, 0);
    std::fill(acc1, acc1 + w, 0);
    std::vector<u64> alpha(w), tb(w), ta(w);
    for (size_t i = 0; i < l; ++i) {
        for_limbs(l, [&](size_t j) {
            u64 qj = b->moduli[j];
            for (u64 k = 0; k < n; ++k) alpha[j * n + k] = src[i * n + k] % qj;
        });
        (void)orc_from_channels_check(b, alpha.data(), l);  // from_channels O(L*N) scan (:517)
        tb = alpha;
        orc_mul_assign(b, tb.data(), key_b + i * w, 0);
        orc_add_assign(b, acc0, tb.data());
        ta = alpha;
        orc_mul_assign(b, ta.data(), key_a + i * w, 0);
        orc_add_assign(b, acc1, ta.data());
    }
}

extern "C" void orc_mul_ciphertexts_gadget(const orc_basis *b, const u64 *a0, const u64 *a1, const u64 *b0,
                                           const u64 *b1, const u64 *rlk_a, const u64 *rlk_b, u64 *c0,
                                           u64 *c1) {  // :473-539
    size_t w = b->moduli.size() * b->n;
    std::vector<u64> d0(a0, a0 + w), d1a(a0, a0 + w), d1b(a1, a1 + w), d2(a1, a1 + w);
    orc_mul_assign(b, d0.data(), b0, 0);   // c0*c0'
    orc_mul_assign(b, d1a.data(), b1, 0);  // c0*c1'
    orc_mul_assign(b, d1b.data(), b0, 0);  // c1*c0'
    orc_add_assign(b, d1a.data(), d1b.data());
    orc_mul_assign(b, d2.data(), b1, 0);   // c1*c1' (already coefficient domain)
    std::vector<u64> r0(w), r1(w);
    gadget_accumulate(b, d2.data(), rlk_a, rlk_b, r0.data(), r1.data());
    orc_add_assign(b, d0.data(), r0.data());
    orc_add_assign(b, d1a.data(), r1.data());
    std::memcpy(c0, d0.data(), w * 8);
    std::memcpy(c1, d1a.data(), w * 8);
}

extern "C" int orc_rescale_ciphertext(const orc_basis *b, const u64 *c0, const u64 *c1, u64 *o0, u64 *o1,
                                      uint32_t *bits_dropped) {  // :263-282
    u64 q_last = b->moduli.back();
    if (bits_dropped) *bits_dropped = 64 - (uint32_t)__builtin_clzll(q_last);
    orc_basis *nb = nullptr;  // the reference builds the dropped basis (deep table copy) first
    int rc = orc_basis_drop_last(b, 1, &nb);
    if (rc != ORC_OK) return rc;
    rc = orc_rescale(b, c0, 0, o0);
    if (rc == ORC_OK) rc = orc_rescale(b, c1, 0, o1);
    orc_basis_free(nb);
    return rc;
}

extern "C" void orc_rotate_ciphertext(const orc_basis *b, const u64 *c0, const u64 *c1, const u64 *rotk_a,
                                      const u64 *rotk_b, int32_t rotation, u64 *o0, u64 *o1) {  // :412-463
    size_t w = b->moduli.size() * b->n;
    std::vector<u64> c0r(w), c1r(w), k0(w), k1(w);
    orc_rotate_slots(b, c0, 0, rotation, c0r.data());
    orc_rotate_slots(b, c1, 0, rotation, c1r.data());
    gadget_accumulate(b, c1r.data(), rotk_a, rotk_b, k0.data(), k1.data());
    orc_add_assign(b, c0r.data(), k0.data());
    std::memcpy(o0, c0r.data(), w * 8);
    std::memcpy(o1, k1.data(), w * 8);
}

extern "C" void orc_gen_public_key(const orc_basis *b, const u64 *s, const u64 *a, const u64 *e, u64 *out_b) {
    // src/keys/public_key.rs:111-131
    size_t w = b->moduli.size() * b->n;
    std::memcpy(out_b, a, w * 8);
    orc_mul_assign(b, out_b, s, 0);
    orc_neg(b, out_b);
    orc_add_assign(b, out_b, e);
}

static void gen_gadget_key(const orc_basis *b, const u64 *s, const u64 *target, const u64 *a, const u64 *e,
                           u64 *out_b) {  // engine.rs:304-332 / :364-392
    size_t l = b->moduli.size();
    u64 n = b->n;
    size_t w = l * n;
    std::vector<u64> plain(w);
    for (size_t i = 0; i < l; ++i) {
        std::fill(plain.begin(), plain.end(), 0);
        std::memcpy(plain.data() + i * n, target + i * n, n * 8);  // e_i * target: limb i only
        u64 *bi = out_b + i * w;
        std::memcpy(bi, a + i * w, w * 8);
        orc_mul_assign(b, bi, s, 0);
        orc_neg(b, bi);
        orc_add_assign(b, bi, e + i * w);
        orc_add_assign(b, bi, plain.data());
    }
}
extern "C" void orc_gen_gadget_relin_key(const orc_basis *b, const u64 *s, const u64 *a, const u64 *e,
                                         u64 *out_b) {  // :288-335
    size_t w = b->moduli.size() * b->n;
    std::vector<u64> s2(s, s + w);
    orc_mul_assign(b, s2.data(), s, 0);  // s^2, coefficient domain
    gen_gadget_key(b, s, s2.data(), a, e, out_b);
}
extern "C" void orc_gen_gadget_rotation_key(const orc_basis *b, const u64 *s, int32_t rotation, const u64 *a,
                                            const u64 *e, u64 *out_b) {  // :348-399
    size_t w = b->moduli.size() * b->n;
    std::vector<u64> sk(w);
    orc_rotate_slots(b, s, 0, rotation, sk.data());
    gen_gadget_key(b, s, sk.data(), a, e, out_b);
}

// ---------------------------------------------------------------------------------------------
// encoder -- src/encoding/special_fft.rs, ckks_encoder.rs (f64; tolerance checks only)
// ---------------------------------------------------------------------------------------------
typedef std::complex<double> cplx;
static u64 pow_mod_small(u64 base, u64 e, u64 m) {  // special_fft.rs:8-19 (no u128: 2N is small)
    u64 acc = 1;
    base %= m;
    while (e) {
        if (e & 1) acc = (acc * base) % m;
        base = (base * base) % m;
        e >>= 1;
    }
    return acc;
}
// Complex::powu(u32) of the `num-complex` crate: exponentiation by squaring, in this order.
static cplx powu(cplx base, uint32_t exp) {
    if (exp == 0) return cplx(1.0, 0.0);
    while ((exp & 1) == 0) {
        base = base * base;
        exp >>= 1;
    }
    if (exp == 1) return base;
    cplx acc = base;
    while (exp > 1) {
        exp >>= 1;
        base = base * base;
        if (exp & 1) acc = acc * base;
    }
    return acc;
}
static void slot_roots(u64 n, std::vector<cplx> &roots, std::vector<cplx> &roots_inv) {  // special_fft.rs:88-137
    cplx psi = std::polar(1.0, M_PI / (double)n);
    std::vector<u64> ex;
    u64 m = 2 * n;
    for (u64 h = 0; h < n / 2; ++h) ex.push_back(pow_mod_small(5, h, m));
    for (u64 h = n / 2; h-- > 0;) ex.push_back((m - pow_mod_small(5, h, m)) % m);
    roots.clear();
    roots_inv.clear();
    for (u64 e : ex) {
        cplx r = powu(psi, (uint32_t)e);
        roots.push_back(r);
        roots_inv.push_back(std::conj(r));
    }
}
static inline cplx cmul(cplx a, cplx b) {  // plain (re,im) product as num-complex does, no NaN fix-ups
    return cplx(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real());
}
extern "C" void orc_encode(u64 n, uint32_t scale_bits, const double *values, size_t nvals, i64 *out) {
    // ckks_encoder.rs:65-122: scale, build_conjugate_slots (special_fft.rs:158-178), special_idft (:194-220), round
    double delta = std::ldexp(1.0, (int)scale_bits);
    std::vector<cplx> slots(n, cplx(0, 0));
    for (u64 idx = 0; idx < n / 2; ++idx) {
        cplx v = idx < nvals ? cplx(values[2 * idx] * delta, values[2 * idx + 1] * delta) : cplx(0, 0);
        slots[idx] = v;
        slots[n - 1 - idx] = std::conj(v);
    }
    std::vector<cplx> roots, roots_inv;
    slot_roots(n, roots, roots_inv);
    std::vector<cplx> coeffs(n, cplx(0, 0));
    for (u64 s = 0; s < n; ++s) {
        cplx value = slots[n - 1 - s];  // permuted = reversed input
        cplx power(1.0, 0.0);
        for (u64 c = 0; c < n; ++c) {
            coeffs[c] += cmul(value, power);
            power = cmul(power, roots[s]);
        }
    }
    double inv_n = 1.0 / (double)n;
    for (u64 c = 0; c < n; ++c) out[c] = (i64)std::round(coeffs[c].real() * inv_n);
}
extern "C" void orc_decode(u64 n, uint32_t scale_bits, const i64 *coeffs, size_t nslots, double *out) {
    // ckks_encoder.rs:134-156, special_dft special_fft.rs:224-242
    double delta = std::ldexp(1.0, (int)scale_bits);
    std::vector<cplx> roots, roots_inv;
    slot_roots(n, roots, roots_inv);
    std::vector<cplx> slots(n, cplx(0, 0));
    for (u64 s = 0; s < n; ++s) {
        cplx power(1.0, 0.0);
        cplx acc(0, 0);
        for (u64 c = 0; c < n; ++c) {
            acc += cmul(cplx((double)coeffs[c], 0.0), power);
            power = cmul(power, roots_inv[s]);
        }
        slots[s] = acc;
    }
    std::reverse(slots.begin(), slots.end());
    for (size_t i = 0; i < nslots && i < n; ++i) {
        out[2 * i] = slots[i].real() / delta;
        out[2 * i + 1] = slots[i].imag() / delta;
    }
}

// ---------------------------------------------------------------------------------------------
// CPU baseline drivers
// ---------------------------------------------------------------------------------------------
// `count` independent units over `threads` host threads: min(count, threads) outer workers, each
// spreading the per-limb work of its unit over threads / workers inner threads.
template <class F>
static double run_parallel(size_t count, int threads, F fn) {
    if (threads < 1) threads = 1;
    int outer = (size_t)threads < count ? threads : (int)count;
    if (outer < 1) outer = 1;
    int inner = threads / outer;
    if (inner < 1) inner = 1;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < outer; ++t)
        pool.emplace_back([&, t]() {
            limb_threads = inner;
            for (size_t k = (size_t)t; k < count; k += (size_t)outer) fn(k);
        });
    for (auto &th : pool) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
extern "C" double orc_bench_mul_rescale(const orc_basis *b, size_t count, int threads, const u64 *a0,
                                        const u64 *a1, const u64 *b0, const u64 *b1, const u64 *rlk_a,
                                        const u64 *rlk_b, u64 *o0, u64 *o1) {
    size_t l = b->moduli.size();
    size_t w = l * b->n, wo = (l - 1) * b->n;
    return run_parallel(count, threads, [&](size_t k) {
        std::vector<u64> m0(w), m1(w);
        orc_mul_ciphertexts_gadget(b, a0 + k * w, a1 + k * w, b0 + k * w, b1 + k * w, rlk_a, rlk_b, m0.data(),
                                   m1.data());
        orc_rescale_ciphertext(b, m0.data(), m1.data(), o0 + k * wo, o1 + k * wo, nullptr);
    });
}
// mul_ciphertexts_gadget alone (engine.rs:473-539), same threading: lets the parity tests check the
// unrescaled limbs at N = 2^16, L = 24 in seconds instead of minutes.
extern "C" double orc_bench_mul_gadget(const orc_basis *b, size_t count, int threads, const u64 *a0, const u64 *a1,
                                       const u64 *b0, const u64 *b1, const u64 *rlk_a, const u64 *rlk_b, u64 *o0,
                                       u64 *o1) {
    size_t w = b->moduli.size() * b->n;
    return run_parallel(count, threads, [&](size_t k) {
        orc_mul_ciphertexts_gadget(b, a0 + k * w, a1 + k * w, b0 + k * w, b1 + k * w, rlk_a, rlk_b, o0 + k * w,
                                   o1 + k * w);
    });
}
extern "C" double orc_bench_rotate(const orc_basis *b, size_t count, int threads, const u64 *c0, const u64 *c1,
                                   const u64 *rotk_a, const u64 *rotk_b, int32_t rotation, u64 *o0, u64 *o1) {
    size_t w = b->moduli.size() * b->n;
    return run_parallel(count, threads, [&](size_t k) {
        orc_rotate_ciphertext(b, c0 + k * w, c1 + k * w, rotk_a, rotk_b, rotation, o0 + k * w, o1 + k * w);
    });
}
extern "C" double orc_bench_ntt(const orc_basis *b, size_t count, int threads, int dir, u64 *polys) {
    size_t w = b->moduli.size() * b->n;
    return run_parallel(count, threads, [&](size_t k) {
        if (dir == 0) orc_to_ntt_domain(b, polys + k * w);
        else orc_to_coeff_domain(b, polys + k * w);
    });
}
