// aux_ks.cuh -- kernels of the exact multi-modular gadget key-switch ("auxiliary-basis key-switch").
//
// The reference's gadget product (engine.rs:501-541) is, per target limb j,
//     ks_j = sum_i alpha_i (*) key[i][j]   in Z_{q_j}[X]/(X^N + 1),          alpha_i = limb i of the digit polynomial,
// computed there with one NTT mod q_j per (digit, target) pair: L^2 transforms of 61-bit words.  The same ring element
// is the image mod q_j of the product over the INTEGERS,
//     S_j = sum_i alpha_i (*) key[i][j]   in Z[X]/(X^N + 1),   |coefficients of S_j| < L * N * q_max^2 =: B,
// and S_j can be computed exactly in a few word-sized NTT primes p_0 .. p_{K-1} (all in (2^29, 2^29.5), prod p_k > 2B) chosen by
// this library: NTT_{p_k}(alpha_i mod p_k) does NOT depend on the target limb, so a ciphertext needs L*K forward and
// 2*L*K inverse transforms of 32-bit words instead of L*(L-1) forward transforms of 64-bit words (cfg4, L = 24,
// K = 5: 360 cheap transforms against 552 expensive ones), a multiply-accumulate over the digits in between, and the
// image mod q_j of the centred integer at the end (aux_crt.cuh: base conversion with exact correction, or Garner's
// mixed radix for deep auxiliary bases).  Every step is exact integer arithmetic -- floating point appears twice, each
// time only to pick an integer that integer arithmetic then uses or corrects: the quotient estimate of aux_reduce_sum
// and the multiple of P in aux_image_hps, decided with a margin of 0.146 -- so the result is bit-identical to the
// reference's: pinned on the CPU by tests/test_emul.py
// (the algorithm end to end with this code compiled by g++) and on the GPU by tests/test_gpu_engine.py against the oracle.
//
// Layouts (all u32 words):
//   x      [ct][k][i][N]   NTT_{p_k}(alpha_i mod p_k), device-internal NTT order of the auxiliary tables
//   key    [k][j][i][N]    NTT_{p_k}(key[i][j] mod p_k), same order (ksk->xb / ->xa)
//   r      [ct][j][k][N]   sum_i x * key  (NTT domain, then transformed back in place to the coefficient domain)
#pragma once
#include "aux_crt.cuh"
#include "modarith.cuh"

// key[i][j][n] mod p_k for every auxiliary prime: src u64 [L(i)][L(j)][N] coefficient domain, dst u32 [K][L(j)][L(i)][N].
__global__ void aux_key_reduce_kernel(const u64 *__restrict__ src, u32 *__restrict__ dst, const LimbConst *__restrict__ alc, int L,
                                      int K, int logn) {
    const size_t n = (size_t)1 << logn;
    const size_t per_k = (size_t)L * L * n;
    const size_t total = per_k * K;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t k = t / per_k, r = t % per_k;
        const size_t e = r & (n - 1), ji = r >> logn;
        const size_t j = ji / L, i = ji % L;
        dst[t] = (u32)barrett_word(src[(i * L + j) * n + e], alc[k]);
    }
}

struct AuxMacArgs {
    const u32 *kb, *ka;    // [K][L][L][N]
    u32 *rb, *ra;          // [cs][L][K][N]
    const LimbConst *alc;  // [K] auxiliary primes
    int L, K, logn;
    int jb;                // target limbs per CTA (blockDim.y); the grid's y counts (prime, j-block) pairs
    unsigned cs;
};
struct AuxMacMap {
    alignas(64) unsigned char x[128];  // CUtensorMap of x seen as (cols N, rows L, slabs cs * K), box [L][32]
};
__device__ __forceinline__ void aux_mad(u64 &acc, u32 x, u32 k) {
#ifdef __CUDA_ARCH__
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x), "r"(k));
#else
    acc += (u64)x * k;
#endif
}
// r[ct][j][k][e] = sum_i x[ct][k][i][e] * key[k][j][i][e] mod p_k for both key halves.
// One thread: one NTT position e (threadIdx.x, 32 per CTA), one target limb j (threadIdx.y), one auxiliary prime.
// Its 2 * L key words are loaded ONCE into registers and stay there while the CTA walks over every ciphertext of the
// launch, so the key -- the largest operand, K * L^2 * N words per half -- crosses HBM once per launch and never
// touches shared memory.  The x rows of a ciphertext (L rows of 32 words, shared by the warps of the CTA: one per target limb, at most 12) arrive
// as one TMA box per ciphertext, AUX_MAC_T ciphertexts per stage, two stages in flight; thread 0 issues the copies,
// an mbarrier per stage signals their arrival, one __syncthreads per stage releases the buffer.
// Products are below p^2 < 2^59 (AUX_P_BOUND): the L <= 32 of them fit a 64-bit accumulator, reduced once at the end
// (aux_reduce_sum: the quotient comes from the FP64 pipe, which is idle here; the integer pipe is the one this kernel
// saturates).
// L4 = ceil(L / 4) (compile time: the key registers); digits L .. 4 L4 - 1 have zero key words (their x rows in
// shared memory are never written and may hold anything).
#ifndef CKKS_AUX_MAC_T
#define CKKS_AUX_MAC_T 4
#endif
constexpr int AUX_MAC_T = CKKS_AUX_MAC_T;  // ciphertexts per stage
template <int L4>
__global__ void __launch_bounds__(L4 <= 6 ? 768 : 512, 1) aux_mac_kernel(AuxMacArgs a, const __grid_constant__ AuxMacMap map) {
    constexpr int T = AUX_MAC_T;
    constexpr int TILE = 4 * L4 * 32;  // words per ciphertext tile
    __shared__ __align__(128) u32 xs[2 * T * TILE];
    __shared__ __align__(8) u64 bars[2];
    const int L = a.L, K = a.K;
    const size_t n = (size_t)1 << a.logn;
    const int tx = threadIdx.x, jy = threadIdx.y, jb = a.jb;
    const int tid = jy * 32 + tx;
    const int e0 = blockIdx.x * 32;
    const size_t e = (size_t)e0 + tx;
    const int jblocks = (L + jb - 1) / jb;
    const int k = blockIdx.y / jblocks, j = (blockIdx.y % jblocks) * jb + jy;
    const bool live = j < L;
    const u32 p = (u32)a.alc[k].q;
    const double pinv = 1.0 / (double)p;
    const unsigned cs = a.cs;
    const unsigned niter = (cs + T - 1) / T;
    const unsigned tile_bytes = (unsigned)L * 32 * 4;
    auto issue = [&](unsigned it) {  // thread 0: the boxes of stage `it`
        u64 *bar = bars + (it & 1);
        const unsigned c0 = it * T, nv = cs - c0 < (unsigned)T ? cs - c0 : (unsigned)T;
        mbar_expect_tx(bar, nv * tile_bytes);
        for (unsigned c = 0; c < nv; ++c) tma_load_tile(xs + ((it & 1) * T + c) * TILE, map.x, e0, (int)((c0 + c) * K + k), bar);
    };
    if (tid == 0) {
        mbar_init(bars + 0, 1);
        mbar_init(bars + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        issue(0);
        if (niter > 1) issue(1);
    }
    u32 kb[4 * L4], ka[4 * L4];
    {   // (volatile: one running pointer instead of 8 L4 addresses held in registers at once)
        const u32 *pb = a.kb + (((size_t)k * L + (live ? j : 0)) * L) * n + e;
        const u32 *pa = a.ka + (((size_t)k * L + (live ? j : 0)) * L) * n + e;
#pragma unroll
        for (int i = 0; i < 4 * L4; ++i) {
            kb[i] = ka[i] = 0u;
            if (live && i < L) {
#ifdef __CUDA_ARCH__
                asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(kb[i]) : "l"(pb));
                asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(ka[i]) : "l"(pa));
#endif
            }
            pb += n;
            pa += n;
        }
    }
    __syncthreads();  // the barriers are initialised
    const size_t oct = (size_t)L * K * n;  // words between ciphertexts in r
    u32 *prb = a.rb + ((size_t)(live ? j : 0) * K + k) * n + e;
    u32 *pra = a.ra + ((size_t)(live ? j : 0) * K + k) * n + e;
    for (unsigned it = 0; it < niter; ++it) {
        const int buf = it & 1;
        mbar_wait(bars + buf, (it >> 1) & 1);
        if (live) {
#pragma unroll
            for (int c = 0; c < T; ++c) {
                if (it * T + c < cs) {
                    const u32 *xc = xs + (buf * T + c) * TILE + tx;
                    u64 sb = 0, sa = 0;
#pragma unroll
                    for (int i = 0; i < 4 * L4; ++i) {
                        const u32 x = xc[i * 32];
                        aux_mad(sb, x, kb[i]);
                        aux_mad(sa, x, ka[i]);
                    }
                    *prb = aux_reduce_sum(sb, p, pinv);
                    *pra = aux_reduce_sum(sa, p, pinv);
                    prb += oct;
                    pra += oct;
                }
            }
        }
        __syncthreads();  // everybody is done with this buffer
        if (tid == 0 && it + 2 < niter) issue(it + 2);
    }
}

struct AuxCrtArgs {
    const u32 *rb, *ra;      // [cs][L][K][N] residues of S_j (coefficient domain)
    const u64 *add0, *add1;  // [cs][L][N] coefficient domain addends (d0, d1; automorphism(c0), none) -- either may be null
    u64 *out0, *out1;        // [cs][outL][N]; limb j lands in slot j - j0
    const u64 *last0, *last1;  // RESCALE: [cs][N] the finished last limb of c0 / c1
    const LimbConst *lc;     // [L] ciphertext primes
    const tw_t *mix;         // [L][K]: prod_{m<k} p_m mod q_j
    const u64 *pmod;         // [L]: P mod q_j
    const u64 *mstar;        // [L][K]: (P / p_k) mod q_j
    const u64 *tp;           // [L][AUX_MAX_K + 1]: t P mod q_j
    const tw_t *ql;          // RESCALE: [L] q_last^-1 mod q_j
    int L, K, logn;
    int j0, nj, outL;        // target limbs j0 .. j0 + nj - 1 of every ciphertext
    size_t rows;             // cs * nj (ciphertext, limb) pairs
};
// Garner mixed-radix digits of the residues (0 <= v_k < p_k, value = sum_k v_k prod_{m<k} p_m in [0, P)), sign by
// comparison with floor(P/2), image mod q_j: sum_k v_k (prod_{m<k} p_m mod q_j) - [negative] (P mod q_j); then the
// addend (d0 / d1 in the coefficient domain), and with RESCALE the epilogue of rescale_into (poly.rs:214-225):
// (c_j - c_last mod q_j) * q_last^-1 mod q_j.  grid = (N / (256 EPT), cs * nj [, continued in z]); a thread owns EPT coefficients
// (independent Garner chains to overlap) of one (ciphertext, target limb); everything that depends only on the
// limb is loaded once per thread.
template <int K, int EPT, bool RESCALE>
__global__ void __launch_bounds__(256) aux_crt_kernel(AuxCrtArgs a, const __grid_constant__ AuxCrtConst cc) {
    const size_t n = (size_t)1 << a.logn;
    const int L = a.L;
    const size_t row = (size_t)blockIdx.z * gridDim.y + blockIdx.y;  // (ciphertext, limb) pairs: y, continued in z past 32768
    if (row >= a.rows) return;
    const size_t ct = row / a.nj;
    const int j = a.j0 + (int)(row % a.nj);
    const LimbConst mq = a.lc[j];
    tw_t mix[K];
#pragma unroll
    for (int k = 1; k < K; ++k) mix[k] = ldg_tw(a.mix + (size_t)j * K + k);
    const u64 pmod = a.pmod[j];
#ifndef CKKS_AUX_CRT_HPS
#define CKKS_AUX_CRT_HPS 1
#endif
    constexpr bool HPS = CKKS_AUX_CRT_HPS && K <= 5;  // (deeper auxiliary bases keep the Garner chain)
    u64 ms[K];
#pragma unroll
    for (int k = 0; k < K; ++k) ms[k] = HPS ? a.mstar[(size_t)j * K + k] : 0;
    tw_t ql;
    if (RESCALE) ql = ldg_tw(a.ql + j);
    const size_t e0 = (size_t)blockIdx.x * (256 * EPT) + threadIdx.x;
    const size_t ro = (ct * L + j) * K * n + e0;
    const size_t ao = (ct * L + j) * n + e0;
    const size_t oo = (ct * a.outL + (j - a.j0)) * n + e0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const u32 *r = (h ? a.ra : a.rb) + ro;
        u32 v[EPT][K];
#pragma unroll
        for (int u = 0; u < EPT; ++u)
#pragma unroll
            for (int k = 0; k < K; ++k) v[u][k] = r[(size_t)k * n + u * 256];
        u64 addv[EPT], lastv[EPT];
#pragma unroll
        for (int u = 0; u < EPT; ++u) {
            const u64 *add = h ? a.add1 : a.add0;
            addv[u] = add ? add[ao + u * 256] : 0ull;  // (rotations: nothing is added to the second sum)
            if (RESCALE) lastv[u] = (h ? a.last1 : a.last0)[ct * n + e0 + u * 256];
        }
#pragma unroll
        for (int u = 0; u < EPT; ++u) {  // (independent chains: the unrolled code interleaves them)
            u64 y;
            if (HPS) {
                y = aux_image_hps<(HPS ? K : 1)>(reinterpret_cast<const u32(&)[HPS ? K : 1]>(v[u]), cc, ms, a.tp + (size_t)j * (AUX_MAX_K + 1), mq);
            } else {
                aux_garner<K>(v[u], cc);
                y = aux_image<K>(v[u], aux_negative<K>(v[u], cc), mix, pmod, mq);
            }
            y = addmod(y, addv[u], mq.q);
            if (RESCALE) y = shoup(submod(y, barrett_word(lastv[u], mq), mq.q), ql, mq.q);
            (h ? a.out1 : a.out0)[oo + u * 256] = y;
        }
    }
}

// Throughput of the multiply-accumulate aux_mac_kernel is made of: 32 x 32 -> 64-bit products added into 64-bit
// accumulators (IMAD.WIDE.U32), eight independent chains per thread, every SM busy (ckks_bench_mac32_peak).
// VARY: the multiplier sits in a general register, different per thread, like the key words of aux_mac_kernel (with a
// warp-uniform multiplier the compiler feeds it from a uniform register and the product issues faster).
template <bool VARY>
__global__ void mac32_peak_kernel(u64 *out, int iters, u32 w0) {
    u64 v[8];
    u32 w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        v[k] = (u64)threadIdx.x * 977 + k * 31 + blockIdx.x;
        w[k] = VARY ? w0 - 2 * (threadIdx.x * 8 + k) : w0;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) aux_mad(v[k], (u32)v[k], w[k]);
    }
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s ^= v[k];
    if (s == 0x123456789abcdefull) out[0] = s;
}
