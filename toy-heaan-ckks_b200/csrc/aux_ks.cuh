// aux_ks.cuh -- kernels of the exact multi-modular gadget key-switch ("auxiliary-basis key-switch").
//
// The reference's gadget product (engine.rs:501-541) is, per target limb j,
//     ks_j = sum_i alpha_i (*) key[i][j]   in Z_{q_j}[X]/(X^N + 1),          alpha_i = limb i of the digit polynomial,
// computed there with one NTT mod q_j per (digit, target) pair: L^2 transforms of 61-bit words.  The same ring element
// is the image mod q_j of the product over the INTEGERS,
//     S_j = sum_i alpha_i (*) key[i][j]   in Z[X]/(X^N + 1),   |coefficients of S_j| < L * N * q_max^2 =: B,
// and S_j can be computed exactly in a few word-sized NTT primes p_0 .. p_{K-1} (all < 2^30, prod p_k > 2B) chosen by
// this library: NTT_{p_k}(alpha_i mod p_k) does NOT depend on the target limb, so a ciphertext needs L*K forward and
// 2*L*K inverse transforms of 32-bit words instead of L*(L-1) forward transforms of 64-bit words (cfg4, L = 24,
// K = 5: 360 cheap transforms against 552 expensive ones), a multiply-accumulate over the digits in between, and a
// Garner reconstruction of the centred integer, reduced mod q_j, at the end.  Every step is exact integer arithmetic,
// so the result is bit-identical to the reference's (tests/test_gpu_engine.py pins it against the oracle).
//
// Layouts (all u32 words):
//   x      [ct][k][i][N]   NTT_{p_k}(alpha_i mod p_k), device-internal NTT order of the auxiliary tables
//   key    [k][j][i][N]    NTT_{p_k}(key[i][j] mod p_k), same order (ksk->xb / ->xa)
//   r      [ct][j][k][N]   sum_i x * key  (NTT domain, then transformed back in place to the coefficient domain)
#pragma once
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
    const u32 *x;          // [cs][K][L][N]
    const u32 *kb, *ka;    // [K][L][L][N]
    u32 *rb, *ra;          // [cs][L][K][N]
    const LimbConst *alc;  // [K] auxiliary primes
    int L, K, logn;
    int jb;                // target limbs per CTA (blockDim.y); the grid's y counts (prime, j-block) pairs
    unsigned cs;
};
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
// r[ct][j][k][e] = sum_i x[ct][k][i][e] * key[k][j][i][e] mod p_k for both key halves.
// One thread: one NTT position e (threadIdx.x, 32 per CTA), one target limb j (threadIdx.y), one auxiliary prime.
// Its 2 * L key words are loaded ONCE into registers and stay there while the CTA walks over every ciphertext of the
// launch, so the key -- the largest operand, K * L^2 * N words per half -- crosses HBM once per launch and never
// touches shared memory.  The x rows of the next two ciphertexts (L rows of 32 words each, shared by the L warps of
// the CTA) are staged with cp.async into a double buffer laid out [ct][i / 4][e][4], so one 16-byte shared load
// feeds four digits.  Products are < 2^60 (p < 2^30): 16 of them plus a carried residue fit a 64-bit accumulator, so
// the accumulators are folded once, after the 16th digit.
// L4 = ceil(L / 4) (compile time: the key registers); digits L .. 4 L4 - 1 are zero padding.
constexpr int AUX_MAC_T = 2;  // ciphertexts per stage
__device__ __forceinline__ void aux_mad(u64 &acc, u32 x, u32 k) {
#ifdef __CUDA_ARCH__
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x), "r"(k));
#else
    acc += (u64)x * k;
#endif
}
template <int L4>
__global__ void __launch_bounds__(L4 <= 6 ? 768 : 512, 1) aux_mac_kernel(AuxMacArgs a) {
    constexpr int T = AUX_MAC_T;
    constexpr int STAGE = T * L4 * 32 * 4;  // words per stage
    __shared__ __align__(16) u32 xs[2 * STAGE];
    const int L = a.L, K = a.K;
    const size_t n = (size_t)1 << a.logn;
    const int tx = threadIdx.x, jy = threadIdx.y, jb = a.jb;
    const int nthr = 32 * jb, tid = jy * 32 + tx;
    const size_t e = (size_t)blockIdx.x * 32 + tx;
    const int jblocks = (L + jb - 1) / jb;
    const int k = blockIdx.y / jblocks, j = (blockIdx.y % jblocks) * jb + jy;
    const bool live = j < L;
    const LimbConst m = a.alc[k];
    for (int w = tid; w < 2 * STAGE; w += nthr) xs[w] = 0;  // the padding digits stay zero for good
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
    __syncthreads();
    const unsigned cs = a.cs;
    const unsigned niter = (cs + T - 1) / T;
    const size_t xct = (size_t)K * L * n;                  // words between ciphertexts in x
    const u32 *px = a.x + ((size_t)k * L) * n + e;         // row i of ciphertext ct: px + ct * xct + i * n
    const size_t oct = (size_t)L * K * n;                  // words between ciphertexts in r
    const size_t o0 = ((size_t)(live ? j : 0) * K + k) * n + e;
    // thread (tx, jy) stages word tx of rows jy, jy + jb, .. of ciphertexts T*it .. T*it+T-1 (clamped to the last one)
    auto stage = [&](unsigned it, int buf) {
#pragma unroll
        for (int c = 0; c < T; ++c) {
            unsigned ct = it * T + c;
            ct = ct < cs ? ct : cs - 1;
            for (int i = jy; i < L; i += jb)
                cp_async4(&xs[buf * STAGE + (((c * L4 + (i >> 2)) * 32 + tx) << 2) + (i & 3)], px + (size_t)ct * xct + (size_t)i * n);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    stage(0, 0);
    for (unsigned it = 0; it < niter; ++it) {
        const int buf = it & 1;
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();  // stage `it` is visible, and everybody is done with the other buffer
        if (it + 1 < niter) stage(it + 1, buf ^ 1);
        if (live) {
#pragma unroll
            for (int c = 0; c < T; ++c) {
                u64 sb = 0, sa = 0;
#pragma unroll
                for (int i4 = 0; i4 < L4; ++i4) {
                    const uint4 xv = *reinterpret_cast<const uint4 *>(&xs[buf * STAGE + (((c * L4 + i4) * 32 + tx) << 2)]);
                    aux_mad(sb, xv.x, kb[4 * i4]);
                    aux_mad(sa, xv.x, ka[4 * i4]);
                    aux_mad(sb, xv.y, kb[4 * i4 + 1]);
                    aux_mad(sa, xv.y, ka[4 * i4 + 1]);
                    aux_mad(sb, xv.z, kb[4 * i4 + 2]);
                    aux_mad(sa, xv.z, ka[4 * i4 + 2]);
                    aux_mad(sb, xv.w, kb[4 * i4 + 3]);
                    aux_mad(sa, xv.w, ka[4 * i4 + 3]);
                    if (i4 == 3 && L4 > 4) {
                        sb = barrett_word(sb, m);
                        sa = barrett_word(sa, m);
                    }
                }
                const unsigned ct = it * T + c;
                if (ct < cs) {
                    a.rb[o0 + (size_t)ct * oct] = (u32)barrett_word(sb, m);
                    a.ra[o0 + (size_t)ct * oct] = (u32)barrett_word(sa, m);
                }
            }
        }
    }
}

constexpr int AUX_MAX_K = 8;
struct AuxCrtArgs {
    const u32 *rb, *ra;    // [cs][L][K][N] residues of S_j (coefficient domain)
    const u64 *add0, *add1;  // [cs][L][N] coefficient domain addends (d0, d1) or null
    u64 *out0, *out1;      // [cs][L][N]
    const LimbConst *lc;   // [L] ciphertext primes
    const LimbConst *alc;  // [K] auxiliary primes
    const tw32_t *inv;     // [K][K]: inv[m * K + k] = p_m^-1 mod p_k, m < k
    const tw_t *mix;       // [L][K]: prod_{m<k} p_m mod q_j
    const u64 *pmod;       // [L]: P mod q_j
    const u32 *half;       // [K]: mixed-radix digits of floor(P / 2)
    int L, K, logn;
    size_t total;          // cs * L * N
};
// Garner mixed-radix digits of the residues (0 <= v_k < p_k, value = sum_k v_k prod_{m<k} p_m in [0, P)), sign by
// comparison with floor(P/2), image mod q_j: sum_k v_k (prod_{m<k} p_m mod q_j) - [negative] (P mod q_j); then the
// addend (d0 / d1 in the coefficient domain).  One thread per (ciphertext, target limb, coefficient).
__global__ void aux_crt_kernel(AuxCrtArgs a) {
    const size_t n = (size_t)1 << a.logn;
    const int L = a.L, K = a.K;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < a.total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t e = t & (n - 1), cj = t >> a.logn;  // cj = ct * L + j
        const int j = (int)(cj % L);
        const LimbConst mq = a.lc[j];
        const tw_t *mix = a.mix + (size_t)j * K;
        const size_t ro = cj * K * n + e;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const u32 *r = (h ? a.ra : a.rb) + ro;
            // digits, sign and image mod q_j
            u32 v[AUX_MAX_K];
#pragma unroll
            for (int k = 0; k < AUX_MAX_K; ++k) {
                if (k < K) {
                    const u32 p = (u32)a.alc[k].q;
                    u32 x = r[(size_t)k * n];
#pragma unroll
                    for (int mi = 0; mi < AUX_MAX_K; ++mi) {
                        if (mi < k) {
                            const u32 vm = csub(v[mi], p);  // auxiliary primes lie in (2^29, 2^30): v_m < p_m < 2 p_k
                            const u32 d = x >= vm ? x - vm : x + p - vm;
                            x = shoup(d, ldg_tw(a.inv + mi * K + k), p);
                        }
                    }
                    v[k] = x;
                }
            }
            bool neg = false, decided = false;
#pragma unroll
            for (int k = AUX_MAX_K - 1; k >= 0; --k) {
                if (k < K && !decided) {
                    const u32 hk = a.half[k];
                    if (v[k] != hk) {
                        neg = v[k] > hk;
                        decided = true;
                    }
                }
            }
            u64 y = barrett_word((u64)v[0], mq);
#pragma unroll
            for (int k = 1; k < AUX_MAX_K; ++k)
                if (k < K) y = addmod(y, shoup((u64)v[k], ldg_tw(mix + k), mq.q), mq.q);
            if (neg) y = submod(y, a.pmod[j], mq.q);
            const u64 *add = h ? a.add1 : a.add0;
            if (add) y = addmod(y, add[t], mq.q);
            (h ? a.out1 : a.out0)[t] = y;
        }
    }
}
