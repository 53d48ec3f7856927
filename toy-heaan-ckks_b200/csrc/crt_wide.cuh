// crt_wide.cuh -- centred CRT reconstruction for a basis of ANY size ("a multi-word CRT for decode
// beyond Q < 2^128").
//
// The reference's RnsBasis::reconstruct_centered_coeff (basis.rs:158-180) forms Q = prod q_i in a u128 and therefore
// only works while Q < 2^128 (two 61-bit primes): that is why horner_chain arranges to END with two primes
// (examples/horner_chain.rs:21-35) and why nothing above can be decoded there.  Garner's mixed-radix form needs no
// big integers: with x = v_0 + v_1 q_0 + v_2 q_0 q_1 + ... (0 <= v_i < q_i)
//     v_j = ( ... ((r_j - v_0) q_0^-1 - v_1) q_1^-1 ... - v_{j-1}) q_{j-1}^-1  mod q_j
// costs j modmuls per digit with the L(L-1)/2 precomputed inverses inv[i][j] = q_i^-1 mod q_j.  The centring test
// x > floor(Q/2) (basis.rs:175-179) is a lexicographic comparison of the digits with those of floor(Q/2), and
// Q - x is a digit-wise complement.  One thread per coefficient; the residues of a coefficient sit N words apart, so
// the loads of a warp coalesce.
//
// Outputs (per coefficient):
//   i64: the centred value truncated to 64 bits exactly like the reference's `as i64` (basis.rs:176,179): for
//        Q < 2^128 this is the reference's result bit for bit, and it is the true value whenever |x| < 2^63;
//   f64: the centred value rounded to double (what the decoder consumes; exact below 2^53);
//   overflow flag: some |x| >= 2^63 (the i64 output of that coefficient is the truncation, as in the reference).
#pragma once
#include "modarith.cuh"

constexpr int CRT_MAX_L = 64;

struct CrtWideArgs {
    const u64 *src;       // [batch][L][N] coefficient domain
    const LimbConst *lc;  // [L]
    const tw_t *inv;      // [L][L]: inv[i * L + j] = q_i^-1 mod q_j (Shoup pair), i < j
    const u64 *half;      // [L] mixed-radix digits of floor(Q / 2)
    long long *out_i64;   // [batch][N] or null
    double *out_f64;      // [batch][N] or null
    int *overflow;        // or null
    size_t total;         // batch * N
    int L;
    int logn;
};

__global__ void crt_wide_kernel(CrtWideArgs a) {
    const size_t n = (size_t)1 << a.logn;
    const int L = a.L;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < a.total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t b = t >> a.logn, k = t & (n - 1);
        const u64 *r = a.src + b * (size_t)L * n + k;
        u64 v[CRT_MAX_L];
        for (int j = 0; j < L; ++j) {
            const LimbConst m = a.lc[j];
            u64 x = r[(size_t)j * n];
            for (int i = 0; i < j; ++i) {
                u64 vi = v[i] >= m.q ? barrett_word(v[i], m) : v[i];
                x = shoup(submod(x, vi, m.q), ldg_tw(a.inv + (size_t)i * L + j), m.q);
            }
            v[j] = x;
        }
        // x > floor(Q/2)?  most significant digit first
        bool neg = false;
        for (int j = L - 1; j >= 0; --j) {
            const u64 h = a.half[j];
            if (v[j] != h) {
                neg = v[j] > h;
                break;
            }
        }
        if (neg) {  // digits of Q - x: zeros up to the first non-zero digit f, q_f - v_f there, q_i - 1 - v_i above
            int f = 0;
            while (f < L && v[f] == 0) ++f;
            for (int j = 0; j < L; ++j) {
                const u64 q = a.lc[j].q;
                if (j < f) v[j] = 0;
                else if (j == f) v[j] = q - v[j];
                else v[j] = q - 1 - v[j];
            }
        }
        // magnitude by Horner from the top digit: low 64 bits (wrapping), saturation watch, and in double
        u64 lo = 0;
        bool big = false;
        double d = 0.0;
        for (int j = L - 1; j >= 0; --j) {
            const u64 q = a.lc[j].q;
            const u64 hi = __umul64hi(lo, q);
            u64 nl = lo * q;
            const u64 s = nl + v[j];
            big |= (hi != 0) | (s < nl);
            lo = s;
            d = d * (double)q + (double)v[j];
        }
        if (neg ? (big || lo > 0x8000000000000000ull) : (big || lo >= 0x8000000000000000ull)) {
            if (a.overflow) atomicOr(a.overflow, 1);
        }
        if (a.out_i64) a.out_i64[t] = (long long)(neg ? (u64)0 - lo : lo);
        if (a.out_f64) a.out_f64[t] = neg ? -d : d;
    }
}
