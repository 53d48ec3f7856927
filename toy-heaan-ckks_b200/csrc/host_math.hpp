// host_math.hpp -- host-side number theory needed to build device tables with the SAME roots the
// reference picks.  Follows src/math/primes.rs (Miller-Rabin :67-93, is_ntt_friendly_prime
// :125-131, get_first_prime_down :198-219), src/math/utils.rs:47-80 (generate_primes) and
// src/rings/backends/rns_ntt/basis.rs (find_primitive_root :217-237, mod_inverse :198-210,
// reconstruct_centered_coeff :158-180).
#pragma once
#include <cstdint>
#include <vector>

namespace hm {
typedef unsigned __int128 u128;
typedef __int128 i128;
typedef unsigned long long u64;

inline u64 mul_mod(u64 a, u64 b, u64 q) { return (u64)(((u128)a * b) % q); }
inline u64 pow_mod(u64 b, u64 e, u64 q) {
    if (q == 1) return 0;
    u64 acc = 1 % q;
    b %= q;
    while (e) {
        if (e & 1) acc = mul_mod(acc, b, q);
        b = mul_mod(b, b, q);
        e >>= 1;
    }
    return acc;
}
inline u64 inv_mod(u64 v, u64 m) {  // extended Euclid, result in [0, m)
    i128 r0 = (i128)m, r1 = (i128)(v % m), t0 = 0, t1 = 1;
    while (r1 != 0) {
        i128 qq = r0 / r1;
        i128 r2 = r0 - qq * r1;
        r0 = r1;
        r1 = r2;
        i128 t2 = t0 - qq * t1;
        t0 = t1;
        t1 = t2;
    }
    if (t0 < 0) t0 += (i128)m;
    return (u64)t0;
}
inline u64 shoup_of(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }

inline bool is_prime(u64 n) {  // deterministic for 64-bit with these 12 bases
    static const u64 bases[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return false;
    if (n < 4) return true;
    if (!(n & 1)) return false;
    u64 d = n - 1;
    int r = 0;
    while (!(d & 1)) {
        d >>= 1;
        ++r;
    }
    for (u64 a : bases) {
        if (a >= n) continue;
        u64 x = pow_mod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int i = 1; i < r; ++i) {
            x = mul_mod(x, x, n);
            if (x == n - 1) {
                comp = false;
                break;
            }
        }
        if (comp) return false;
    }
    return true;
}
inline bool is_ntt_friendly_prime(u64 p, u64 n) { return is_prime(p) && (p % (2 * n) == 1); }

inline u64 first_prime_down(u64 bound, u64 n) {  // largest prime < bound with p == 1 mod 2n; 0 = none
    if (bound <= 2) return 0;
    u64 step = 2 * n, value = bound - 1;
    u64 delta = (value % step + step - 1) % step;
    if (delta > value) return 0;
    u64 cand = value - delta;
    for (;;) {
        if (cand <= 2) return 0;
        if (is_prime(cand)) return cand;
        if (cand < step) return 0;
        cand -= step;
    }
}
inline bool generate_primes(int bits, int count, u64 degree, u64 *out) {
    if (bits < 4 || bits > 63 || count <= 0 || degree == 0) return false;
    u64 upper = ((u64)1 << bits) - 1, lower = (u64)1 << (bits - 1);
    u64 cur = first_prime_down(upper + 1, degree);
    int found = 0;
    while (cur && found < count && cur >= lower) {
        out[found++] = cur;
        cur = first_prime_down(cur, degree);
    }
    return found == count;
}

// Smallest candidate c >= 2 whose power c^((q-1)/order) has exact order `order` (a power of two
// here, so the only prime factor to test is 2); returns that power.
inline u64 find_primitive_root(u64 q, u64 order) {
    u64 ex = (q - 1) / order;
    std::vector<u64> fac;
    u64 v = order;
    for (u64 d = 2; d * d <= v; ++d)
        if (v % d == 0) {
            fac.push_back(d);
            while (v % d == 0) v /= d;
        }
    if (v > 1) fac.push_back(v);
    for (u64 c = 2; c < q; ++c) {
        u64 r = pow_mod(c, ex, q);
        if (r == 1) continue;
        bool ok = true;
        for (u64 f : fac)
            if (pow_mod(r, order / f, q) == 1) {
                ok = false;
                break;
            }
        if (ok) return r;
    }
    return 0;
}

inline unsigned brv(unsigned x, int bits) {
    unsigned r = 0;
    for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

inline int64_t reconstruct_centered(const std::vector<u64> &moduli, const u64 *res) {
    u128 q = 1;
    for (u64 m : moduli) q *= (u128)m;
    u128 acc = 0;
    for (size_t i = 0; i < moduli.size(); ++i) {
        u64 m = moduli[i];
        u128 qi = q / m;
        u64 qi_inv = inv_mod((u64)(qi % m), m);
        u128 s = ((u128)res[i] * qi_inv) % m;
        u128 term = s * qi % q;
        acc = (acc + term) % q;
    }
    if (acc > q / 2) return (int64_t)((i128)acc - (i128)q);
    return (int64_t)acc;
}
// Mixed-radix digits (radices mod[0], mod[1], ...; least significant first) of floor(prod(mod) / 2): the centring
// threshold of a Garner reconstruction (crt_wide.cuh, aux_crt.cuh).
inline std::vector<u64> half_q_digits(const std::vector<u64> &mod) {
    std::vector<u64> big(1, 1);  // little-endian words of Q
    for (u64 q : mod) {
        u128 carry = 0;
        for (u64 &w : big) {
            u128 t = (u128)w * q + carry;
            w = (u64)t;
            carry = t >> 64;
        }
        if (carry) big.push_back((u64)carry);
    }
    u64 c = 0;  // big >>= 1
    for (size_t i = big.size(); i-- > 0;) {
        u64 w = big[i];
        big[i] = (w >> 1) | (c << 63);
        c = w & 1;
    }
    std::vector<u64> digits;
    for (u64 q : mod) {  // big, rem = divmod(big, q)
        u128 rem = 0;
        for (size_t i = big.size(); i-- > 0;) {
            u128 cur = (rem << 64) | big[i];
            big[i] = (u64)(cur / q);
            rem = cur % q;
        }
        digits.push_back((u64)rem);
    }
    return digits;
}
}  // namespace hm
