// aux_crt.cuh -- the per-coefficient arithmetic of the auxiliary-basis gadget product's reconstruction (aux_ks.cuh),
// and the host-side construction of its constants.  No CUDA runtime in this file: tests/emul compiles it with g++ and
// checks it against Python integers and the oracle on the CPU (tests/test_emul.py).
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

#include "host_math.hpp"
#include "modarith.cuh"

constexpr int AUX_MAX_K = 8;
// The auxiliary primes lie in (2^29, 2^29.5): 32 p^2 < 2^64, so a sum of up to AUX_MAX_L = 32 products of two residues fits
// a 64-bit accumulator without an intermediate reduction (aux_mac_kernel), and v < p_m < 2 p_k for any two of them (Garner).
constexpr unsigned long long AUX_P_BOUND = 759250124ull;  // floor(2^29.5)
constexpr int AUX_MAX_L = 32;
// Constants of the auxiliary basis, passed by value (kernel parameters live in the constant bank: the unrolled Garner
// chain reads them with immediate offsets, no loads).
struct AuxCrtConst {
    u32 p[AUX_MAX_K];                      // auxiliary primes
    u32 half[AUX_MAX_K];                   // mixed-radix digits of floor(P / 2)
    tw32_t inv[AUX_MAX_K * AUX_MAX_K];     // inv[m * AUX_MAX_K + k] = p_m^-1 mod p_k, m < k
    tw32_t chat[AUX_MAX_K];                // (P / p_k)^-1 mod p_k
    double pinv[AUX_MAX_K];                // 1 / p_k
};
// Residues r_k of an integer x in [0, P) -> its mixed-radix digits v_k (x = sum_k v_k prod_{m<k} p_m, 0 <= v_k < p_k), in place.
// Garner: v_k = (..((r_k - v_0) p_0^-1 - v_1) p_1^-1 .. - v_{k-1}) p_{k-1}^-1 mod p_k.  The running value stays in [0, 2 p_k):
// x + 2 p_k - v_m is positive and below 4 p_k < 2^32 (v_m < p_m < 2^30 < 2 p_k: the auxiliary primes lie in (2^29, 2^30)),
// and the lazy Shoup product takes any word.
template <int K>
__device__ __forceinline__ void aux_garner(u32 (&v)[K], const AuxCrtConst &cc) {
#pragma unroll
    for (int k = 1; k < K; ++k) {
        const u32 p = cc.p[k], p2 = 2 * p;
#pragma unroll
        for (int mi = 0; mi < k; ++mi) v[k] = shoup_lazy((u32)(v[k] + p2 - v[mi]), cc.inv[mi * AUX_MAX_K + k], p);
        v[k] = csub(v[k], p);
    }
}
// x > floor(P / 2), i.e. x stands for the negative integer x - P: most significant digit first.
template <int K>
__device__ __forceinline__ bool aux_negative(const u32 (&v)[K], const AuxCrtConst &cc) {
    bool neg = false, decided = false;
#pragma unroll
    for (int k = K - 1; k >= 0; --k) {
        if (!decided && v[k] != cc.half[k]) {
            neg = v[k] > cc.half[k];
            decided = true;
        }
    }
    return neg;
}
// The centred integer mod q: sum_k v_k (prod_{m<k} p_m mod q) - [negative] (P mod q).  mix: K Shoup pairs of this q.
template <int K>
__device__ __forceinline__ u64 aux_image(const u32 (&v)[K], bool neg, const tw_t *mix, u64 pmod, const LimbConst &mq) {
    u64 y = barrett_word((u64)v[0], mq);
#pragma unroll
    for (int k = 1; k < K; ++k) y = addmod(y, shoup((u64)v[k], mix[k], mq.q), mq.q);
    if (neg) y = submod(y, pmod, mq.q);
    return y;
}

// s mod p for any 64-bit s and p in (2^29, 2^30): floor(s / p) < 2^35 estimated in double precision (relative error
// below 2^-51: off by at most one either way), exact remainder by one 64-bit multiply-subtract and two corrections.
__device__ __forceinline__ u32 aux_reduce_sum(u64 s, u32 p, double pinv) {
#ifdef __CUDA_ARCH__
    const u64 q = (u64)__double2ull_rz(__ull2double_rz(s) * pinv);
#else
    const u64 q = (u64)((double)s * pinv);
#endif
    i64 r = (i64)(s - q * (u64)p);  // in (-p, 2p)
    if (r < 0) r += p;
    if (r >= (i64)p) r -= p;
    return (u32)r;
}

// The same image without the sequential Garner chain (K <= 5), in the form of the "fast base conversion with exact
// correction": with z_k = r_k (P/p_k)^-1 mod p_k,  sum_k z_k (P/p_k) = x + t P  for the centred x and an integer 0 <= t <= K;
// since |x| < 2^-1.5 P = 0.354 P (aux_host_build takes primes until P > 2^1.5 * 2^need >= 2^1.5 * L N q_max^2), the fractional
// part of sum_k z_k / p_k = t + x / P stays 0.146 away from one half and t = round(sum_k z_k / p_k) is decided by a
// double-precision sum whose error is below 1e-14.  The image is (sum_k z_k ((P/p_k) mod q) - (t P mod q)) mod q: the
// sum is accumulated exactly in 96 bits and reduced once.  mstar: K words (P/p_k) mod q; tp: K + 1 words t P mod q.
template <int K>
__device__ __forceinline__ u64 aux_image_hps(const u32 (&r)[K], const AuxCrtConst &cc, const u64 *mstar, const u64 *tp, const LimbConst &mq) {
    static_assert(K <= 5, "the 64-bit halves of the 96-bit sum hold five terms");
    u64 lo = 0, mid = 0;  // sum z_k * low32(M_k) < 5 * 2^61.5,  sum z_k * high32(M_k) < 5 * 2^60.5
    double s = 0.5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const u32 z = shoup(r[k], cc.chat[k], cc.p[k]);
        s += (double)z * cc.pinv[k];
        const u64 m = mstar[k];
        lo += (u64)z * (u32)m;
        mid += (u64)z * (u32)(m >> 32);
    }
    const int t = (int)s;
    const u64 lo2 = lo + (mid << 32);
    const u64 hi = (mid >> 32) + (lo2 < lo ? 1ull : 0ull);
    return submod(reduce128(hi, lo2, mq), tp[t], mq.q);
}

// Host side: the auxiliary primes of a basis and every constant derived from them.
struct AuxHost {
    int K = 0;
    std::vector<u64> primes;
    AuxCrtConst cc;
    std::vector<tw_t> mix;  // [L][K]: prod_{m<k} p_m mod q_j
    std::vector<u64> pmod;  // [L]: P mod q_j
    std::vector<u64> mstar;  // [L][K]: (P / p_k) mod q_j
    std::vector<u64> tp;     // [L][AUX_MAX_K + 1]: t P mod q_j
};
// |coefficients of sum_i alpha_i (*) key[i][j]| < L * N * q_max^2 < 2^need; the centred range of P = prod p_k must cover
// it: P > 2^(need + 1).
inline bool aux_host_build(u64 n, int logn, const std::vector<u64> &moduli, AuxHost &A) {
    const size_t L = moduli.size();
    u64 qmax = 0;
    for (u64 q : moduli) qmax = q > qmax ? q : qmax;
    int lbits = 0;
    while (((size_t)1 << lbits) < L) ++lbits;
    const int qbits = 64 - __builtin_clzll(qmax);
    const int need = lbits + logn + 2 * qbits;
    // the NTT-friendly primes below AUX_P_BOUND, largest first
    std::vector<u64> primes;
    double have = 0.0;
    for (u64 cur = hm::first_prime_down(AUX_P_BOUND, n); cur > (1ull << 29) && (int)primes.size() < AUX_MAX_K && have < need + 1.5;
         cur = hm::first_prime_down(cur, n)) {
        primes.push_back(cur);
        have += std::log2((double)cur);
    }
    const int K = (int)primes.size();
    if (have < need + 1.5 || K < 2) return false;
    A.K = K;
    A.primes = primes;
    memset(&A.cc, 0, sizeof(A.cc));
    const std::vector<u64> half = hm::half_q_digits(primes);
    for (int k = 0; k < K; ++k) {
        A.cc.p[k] = (u32)primes[k];
        A.cc.half[k] = (u32)half[k];
        for (int m = 0; m < k; ++m) {
            const u64 w = hm::inv_mod(primes[m] % primes[k], primes[k]);
            A.cc.inv[m * AUX_MAX_K + k].w = (u32)w;
            A.cc.inv[m * AUX_MAX_K + k].ws = (u32)((w << 32) / primes[k]);
        }
    }
    for (int k = 0; k < K; ++k) {
        u64 prod = 1 % primes[k];
        for (int m = 0; m < K; ++m)
            if (m != k) prod = hm::mul_mod(prod, primes[m] % primes[k], primes[k]);
        const u64 w = hm::inv_mod(prod, primes[k]);
        A.cc.chat[k].w = (u32)w;
        A.cc.chat[k].ws = (u32)((w << 32) / primes[k]);
        A.cc.pinv[k] = 1.0 / (double)primes[k];
    }
    A.mstar.assign(L * K, 0);
    A.tp.assign(L * (AUX_MAX_K + 1), 0);
    A.mix.assign(L * K, tw_t{0, 0});
    A.pmod.assign(L, 0);
    for (size_t j = 0; j < L; ++j) {
        const u64 q = moduli[j];
        u64 acc = 1 % q;
        for (int k = 0; k < K; ++k) {
            A.mix[j * K + k].w = acc;
            A.mix[j * K + k].ws = hm::shoup_of(acc, q);
            acc = hm::mul_mod(acc, primes[k] % q, q);
        }
        A.pmod[j] = acc;
        for (int k = 0; k < K; ++k) {
            u64 prod = 1 % q;
            for (int m = 0; m < K; ++m)
                if (m != k) prod = hm::mul_mod(prod, primes[m] % q, q);
            A.mstar[j * K + k] = prod;
        }
        for (int t = 0; t <= AUX_MAX_K; ++t) A.tp[j * (AUX_MAX_K + 1) + t] = hm::mul_mod((u64)t % q, acc, q);
    }
    return true;
}
