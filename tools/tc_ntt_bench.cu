// tc_ntt_bench.cu -- "should the NTT be an int8-split GEMM on the tensor cores?"
//
// north_star: "Tensor cores are used only if an int8-split NTT-as-GEMM variant beats the CUDA-core path on ncu
// evidence."  This standalone microbenchmark produces that evidence on sm_100a with the Blackwell-native instructions
// (tcgen05.mma kind::i8 issued by one thread, accumulators in TMEM, tcgen05.ld for the epilogue):
//
//   part 1  raw rate of tcgen05.mma.kind::i8 (M=128, N=256, K=32, u8 x u8 -> s32), checked against a host GEMM;
//   part 2  a 128-point DFT over Z_q, q a 30-bit NTT prime (7 radix-2 stages' worth of a limb transform), as the
//           int8-split GEMM north_star describes: W = sum_a 2^(8a) W_a, X = sum_b 2^(8b) X_b (four byte planes each),
//           16 partial products D_ab = W_a X_b on the tensor cores (4 accumulator groups in TMEM, N = 32 columns x 4
//           planes), then on the CUDA cores: tcgen05.ld, recombination sum 2^(8(a+b)) D_ab and reduction mod q.
//           Every output is checked against the O(n^2) DFT on the host; the stages are timed separately
//           (byte-plane staging of X / MMA / epilogue) and together.
//
// The comparison figure is the library's own transform: bench.py's `ntt` record gives transforms/s of the 30-bit
// four-step NTT at N = 2^16 (16 radix-2 stages per element), i.e. element-stages per second.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/tc_ntt_bench tools/tc_ntt_bench.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef unsigned long long u64;
typedef unsigned int u32;

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e__ = (x);                                                                  \
        if (e__ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e__)); \
            exit(2);                                                                            \
        }                                                                                       \
    } while (0)

// ---- canonical K-major, no-swizzle operand layout of tcgen05.mma (cute/arch/mma_sm100_desc.hpp) ----------------
// An [R rows][K bytes] operand is cut into core matrices of 8 rows x 16 bytes (128 contiguous bytes, rows 16 B apart).
// Core matrices that are neighbours along K lie LBO = 128 B apart, neighbours along the rows SBO = (K/16)*128 B apart.
__host__ __device__ inline size_t op_off(int r, int k, int K) {
    return (size_t)(r >> 3) * (size_t)(K >> 4) * 128 + (size_t)(k >> 4) * 128 + (size_t)(r & 7) * 16 + (size_t)(k & 15);
}
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
// 64-bit shared-memory matrix descriptor: start address, leading / stride byte offsets (all >> 4), version 1 (sm_100),
// layout type 0 = no swizzle.
__device__ __forceinline__ u64 make_desc(u32 saddr, u32 lbo, u32 sbo) {
    return (u64)((saddr & 0x3FFFFu) >> 4) | ((u64)(lbo >> 4) << 16) | ((u64)(sbo >> 4) << 32) | (1ull << 46);
}
// 32-bit instruction descriptor for kind::i8: D = s32 (2 at bits 4-5), A and B unsigned 8 bit (0 at bits 7-9 / 10-12),
// both K-major (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__host__ __device__ constexpr u32 make_idesc(int M, int N) { return (2u << 4) | ((u32)(N >> 3) << 17) | ((u32)(M >> 4) << 24); }

__device__ __forceinline__ void tmem_alloc(u32 *slot, u32 ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(u32 taddr, u32 ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_i8(u32 tmem_d, u64 adesc, u64 bdesc, u32 idesc, u32 accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(u64 *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    // bounded: a descriptor mistake must end in a trap (reported by the host), never in a kernel that spins for ever
    for (unsigned spin = 0; spin < (1u << 26); ++spin) {
        u32 done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// 16 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(u32 taddr, u32 (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// =====================================================================================================================
// part 1: raw tcgen05.mma.kind::i8 rate.  One CTA per SM, A [128][32*KB] and B [256][32*KB] staged once, `iters` rounds
// of KB instructions (M=128, N=256, K=32 each), one commit per round.  out: the accumulator tile of CTA 0.
// =====================================================================================================================
template <int KB>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(const uint8_t *__restrict__ A, const uint8_t *__restrict__ B, int iters, int *__restrict__ out, int swap) {
    constexpr int M = 128, N = 256, K = 32 * KB;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem, *sB = smem + (size_t)M * K;
    __shared__ __align__(8) u64 bar;
    __shared__ u32 tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * K / 16; i += 128) reinterpret_cast<uint4 *>(sA)[i] = reinterpret_cast<const uint4 *>(A)[i];
    for (int i = tid; i < N * K / 16; i += 128) reinterpret_cast<uint4 *>(sB)[i] = reinterpret_cast<const uint4 *>(B)[i];
    if (tid == 0) mbar_init(&bar, 1);
    if (warp == 0) tmem_alloc(&tmem_base, 256);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const u32 tm = tmem_base;
    constexpr u32 idesc = make_idesc(M, N);
    // swap = 1 exchanges the two offsets: the run reports which reading of the descriptor fields reproduces the host GEMM
    const u32 lbo = swap ? (K / 16) * 128 : 128, sbo = swap ? 128 : (K / 16) * 128;
    u32 phase = 0;
    for (int it = 0; it < iters; ++it) {
        if (tid == 0) {
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
                umma_i8(tm, make_desc(smem_u32(sA) + kb * 256, lbo, sbo), make_desc(smem_u32(sB) + kb * 256, lbo, sbo), idesc, (it | kb) ? 1u : 0u);
            umma_commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
    }
    tc_fence_after();
    if (blockIdx.x == 0) {  // lane = row m, column = n
        for (int c0 = 0; c0 < N; c0 += 16) {
            u32 r[16];
            tmem_ld16(tm + ((u32)(warp * 32) << 16) + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) out[(size_t)tid * N + c0 + j] = (int)r[j];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 256);
}

// =====================================================================================================================
// part 2: 128-point DFT mod q (30-bit) as an int8-split GEMM.
//   Wp: [4 planes a][128][128] bytes in the canonical operand layout (prepared on the host, staged once per CTA)
//   X : [128 (j)][ncols] u32 values in [0, q);  Y: [128 (i)][ncols] u32 = sum_j W[i][j] X[j][col] mod q
//   A tile = 32 columns: B operand rows n = col * 4 + b (b = byte plane of X), K = j.
//   TMEM: accumulator group a occupies columns [a * 128, a * 128 + 128): D_a[i][col * 4 + b].
//   mode bits: 1 = split X into byte planes and stage it for every tile (otherwise once), 2 = issue the MMAs,
//              4 = epilogue (tcgen05.ld + recombination + reduction + store).
// =====================================================================================================================
struct DftArgs {
    const uint8_t *Wp;
    const u32 *X;
    u32 *Y;
    int ncols, tiles_per_cta, mode, swap;
    u32 q;
    u64 mu;     // floor(2^64 / q)
    u32 c[7];   // 2^(8 s) mod q
};
__global__ void __launch_bounds__(128, 1) dft128_i8_kernel(DftArgs a) {
    constexpr int M = 128, N = 128, K = 128;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem;                    // 4 planes x 16 KiB
    uint8_t *sB = smem + 4 * (size_t)M * K;  // 16 KiB
    __shared__ __align__(8) u64 bar;
    __shared__ u32 tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 4 * M * K / 16; i += 128) reinterpret_cast<uint4 *>(sA)[i] = reinterpret_cast<const uint4 *>(a.Wp)[i];
    if (tid == 0) mbar_init(&bar, 1);
    if (warp == 0) tmem_alloc(&tmem_base, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const u32 tm = tmem_base;
    constexpr u32 idesc = make_idesc(M, N);
    const u32 lbo = a.swap ? (K / 16) * 128 : 128, sbo = a.swap ? 128 : (K / 16) * 128;
    u32 phase = 0;
    for (int t = 0; t < a.tiles_per_cta; ++t) {
        const int col0 = (blockIdx.x * a.tiles_per_cta + t) * 32;
        if ((a.mode & 1) || t == 0) {
            // byte-plane split of the X tile: thread handles (col, 4 consecutive j): four 32-bit words, one per plane
            for (int e = tid; e < 32 * 32; e += 128) {
                const int col = e & 31, j4 = (e >> 5) * 4;
                u32 x0 = a.X[(size_t)(j4 + 0) * a.ncols + col0 + col], x1 = a.X[(size_t)(j4 + 1) * a.ncols + col0 + col];
                u32 x2 = a.X[(size_t)(j4 + 2) * a.ncols + col0 + col], x3 = a.X[(size_t)(j4 + 3) * a.ncols + col0 + col];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    u32 w = ((x0 >> (8 * b)) & 0xff) | (((x1 >> (8 * b)) & 0xff) << 8) | (((x2 >> (8 * b)) & 0xff) << 16) | (((x3 >> (8 * b)) & 0xff) << 24);
                    *reinterpret_cast<u32 *>(sB + op_off(col * 4 + b, j4, K)) = w;
                }
            }
            fence_async_smem();
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (a.mode & 2) {
            if (tid == 0) {
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb)
                        umma_i8(tm + p * 128, make_desc(smem_u32(sA) + p * (M * K) + kb * 256, lbo, sbo), make_desc(smem_u32(sB) + kb * 256, lbo, sbo), idesc,
                                kb ? 1u : 0u);
                umma_commit(&bar);
            }
            mbar_wait(&bar, phase);
            phase ^= 1;
            tc_fence_after();
        }
        if (a.mode & 4) {
            const u32 lane_base = tm + ((u32)(warp * 32) << 16);
            for (int cg = 0; cg < 8; ++cg) {  // 4 columns (16 TMEM columns) per accumulator group at a time
                u32 d[4][16];
#pragma unroll
                for (int p = 0; p < 4; ++p) tmem_ld16(lane_base + p * 128 + cg * 16, d[p]);
                tmem_ld_wait();
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    u32 s[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll
                    for (int p = 0; p < 4; ++p)
#pragma unroll
                        for (int b = 0; b < 4; ++b) s[p + b] += d[p][cc * 4 + b];
                    u64 acc = 0;
#pragma unroll
                    for (int k = 0; k < 7; ++k) acc += (u64)s[k] * a.c[k];  // < 7 * 2^25 * 2^30
                    u64 qe = __umul64hi(acc, a.mu);
                    u32 r = (u32)(acc - qe * a.q);
                    if (r >= a.q) r -= a.q;
                    a.Y[(size_t)tid * a.ncols + col0 + cg * 4 + cc] = r;
                }
            }
        }
        tc_fence_before();
        __syncthreads();  // TMEM and sB are free again
        tc_fence_after();
    }
    if (warp == 0) tmem_dealloc(tm, 512);
}

// ---- host ----------------------------------------------------------------------------------------------------
static u64 mulmod(u64 a, u64 b, u64 q) { return (u64)((unsigned __int128)a * b % q); }
static u64 powmod(u64 b, u64 e, u64 q) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = mulmod(r, b, q);
        b = mulmod(b, b, q);
        e >>= 1;
    }
    return r;
}
static int g_swap = 0;
static float time_ms(cudaEvent_t e0, cudaEvent_t e1) {
    float ms = 0;
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms;
}

int main(int argc, char **argv) {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const int sms = prop.multiProcessorCount;
    const double ghz = prop.clockRate * 1e-6;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_ghz\": %.3f", prop.name, sms, ghz);
    (void)argc;
    (void)argv;

    // ---------------- part 1 ----------------
    {
        constexpr int KB = 4, M = 128, N = 256, K = 32 * KB;
        std::vector<uint8_t> A((size_t)M * K), B((size_t)N * K), An((size_t)M * K), Bn((size_t)N * K);
        u64 s = 0x9E3779B97F4A7C15ull;
        auto rnd = [&]() {
            s ^= s << 13;
            s ^= s >> 7;
            s ^= s << 17;
            return (uint8_t)(s >> 24);
        };
        for (int m = 0; m < M; ++m)
            for (int k = 0; k < K; ++k) {
                An[(size_t)m * K + k] = rnd();
                A[op_off(m, k, K)] = An[(size_t)m * K + k];
            }
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) {
                Bn[(size_t)n * K + k] = rnd();
                B[op_off(n, k, K)] = Bn[(size_t)n * K + k];
            }
        uint8_t *dA, *dB;
        int *dout;
        CK(cudaMalloc(&dA, A.size()));
        CK(cudaMalloc(&dB, B.size()));
        CK(cudaMalloc(&dout, (size_t)M * N * 4));
        CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
        const size_t smem = (size_t)(M + N) * K;
        CK(cudaFuncSetAttribute(umma_rate_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        std::vector<int> got((size_t)M * N);
        size_t bad = 0, bad_conv[2] = {0, 0};
        for (int conv = 0; conv < 2; ++conv) {  // one round: exact check, for both readings of the offset fields
            CK(cudaMemset(dout, 0xff, (size_t)M * N * 4));
            umma_rate_kernel<KB><<<1, 128, smem>>>(dA, dB, 1, dout, conv);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost));
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < N; ++n) {
                    int ref = 0;
                    for (int k = 0; k < K; ++k) ref += (int)An[(size_t)m * K + k] * (int)Bn[(size_t)n * K + k];
                    if (ref != got[(size_t)m * N + n]) ++bad_conv[conv];
                }
        }
        g_swap = bad_conv[0] <= bad_conv[1] ? 0 : 1;
        bad = bad_conv[g_swap];
        const int iters = 4000;
        umma_rate_kernel<KB><<<sms, 128, smem>>>(dA, dB, 200, dout, g_swap);
        CK(cudaEventRecord(e0));
        umma_rate_kernel<KB><<<sms, 128, smem>>>(dA, dB, iters, dout, g_swap);
        CK(cudaEventRecord(e1));
        float ms = time_ms(e0, e1);
        double macs = (double)sms * iters * KB * M * N * 32;
        printf(", \"umma_i8\": {\"shape\": \"M128 N256 K32 u8*u8->s32, %d per commit\", \"mismatches\": %zu, \"mismatches_other_offset_convention\": %zu, \"tera_mac_per_s\": %.1f, "
               "\"mac_per_clk_per_sm\": %.0f, \"dense_int8_tops\": %.0f}",
               KB, bad, bad_conv[1 - g_swap], macs / (ms * 1e-3) / 1e12, macs / (ms * 1e-3) / sms / (ghz * 1e9), 2 * macs / (ms * 1e-3) / 1e12);
        CK(cudaFree(dA));
        CK(cudaFree(dB));
        CK(cudaFree(dout));
    }

    // ---------------- part 2 ----------------
    {
        const u32 q = 1073479681u;  // 30-bit, q = 1 mod 2^17 (a prime of generate_primes(30, ., 65536))
        u64 g = 2;
        u64 w = 0;
        for (g = 2; g < 1000; ++g) {  // an element of exact order 128
            w = powmod(g, (q - 1) / 128, q);
            if (powmod(w, 64, q) == q - 1) break;
        }
        const int tiles_per_cta = 64, ncols = sms * tiles_per_cta * 32;
        std::vector<u32> W(128 * 128), X((size_t)128 * ncols);
        for (int i = 0; i < 128; ++i)
            for (int j = 0; j < 128; ++j) W[i * 128 + j] = (u32)powmod(w, (u64)i * j % 128, q);
        u64 s = 88172645463325252ull;
        for (auto &x : X) {
            s ^= s << 13;
            s ^= s >> 7;
            s ^= s << 17;
            x = (u32)(s % q);
        }
        std::vector<uint8_t> Wp((size_t)4 * 128 * 128);
        for (int p = 0; p < 4; ++p)
            for (int i = 0; i < 128; ++i)
                for (int j = 0; j < 128; ++j) Wp[(size_t)p * 128 * 128 + op_off(i, j, 128)] = (uint8_t)(W[i * 128 + j] >> (8 * p));
        uint8_t *dW;
        u32 *dX, *dY;
        CK(cudaMalloc(&dW, Wp.size()));
        CK(cudaMalloc(&dX, X.size() * 4));
        CK(cudaMalloc(&dY, X.size() * 4));
        CK(cudaMemcpy(dW, Wp.data(), Wp.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(dY, 0, X.size() * 4));
        DftArgs a;
        a.Wp = dW;
        a.X = dX;
        a.Y = dY;
        a.ncols = ncols;
        a.tiles_per_cta = tiles_per_cta;
        a.q = q;
        a.mu = (u64)(((unsigned __int128)1 << 64) / q);
        for (int k = 0; k < 7; ++k) a.c[k] = (u32)powmod(2, 8 * k, q);
        const size_t smem = (size_t)4 * 128 * 128 + (size_t)128 * 128;
        CK(cudaFuncSetAttribute(dft128_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        a.mode = 7;
        a.swap = g_swap;
        dft128_i8_kernel<<<sms, 128, smem>>>(a);
        CK(cudaDeviceSynchronize());
        std::vector<u32> Y(X.size());
        CK(cudaMemcpy(Y.data(), dY, Y.size() * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0, checked = 0;
        for (int col = 0; col < ncols; col += 997) {  // a few hundred columns, every row
            for (int i = 0; i < 128; ++i) {
                u64 acc = 0;
                for (int j = 0; j < 128; ++j) acc = (acc + mulmod(W[i * 128 + j], X[(size_t)j * ncols + col], q)) % q;
                ++checked;
                if ((u32)acc != Y[(size_t)i * ncols + col]) ++bad;
            }
        }
        printf(", \"dft128_int8_gemm\": {\"q\": %u, \"columns\": %d, \"outputs_checked\": %zu, \"mismatches\": %zu", q, ncols, checked, bad);
        const double outputs = (double)128 * ncols;
        const char *names[] = {"", "stage_only", "mma_only", "stage_mma", "epilogue_only", "stage_epilogue", "mma_epilogue", "all"};
        for (int mode : {1, 2, 4, 6, 7}) {
            a.mode = mode;
            dft128_i8_kernel<<<sms, 128, smem>>>(a);
            CK(cudaEventRecord(e0));
            for (int r = 0; r < 5; ++r) dft128_i8_kernel<<<sms, 128, smem>>>(a);
            CK(cudaEventRecord(e1));
            float ms = time_ms(e0, e1) / 5;
            printf(", \"%s\": {\"ms\": %.4f, \"outputs_per_s\": %.4g, \"element_stages_per_s\": %.4g}", names[mode], ms, outputs / (ms * 1e-3),
                   7 * outputs / (ms * 1e-3));
        }
        printf("}");
        CK(cudaFree(dW));
        CK(cudaFree(dX));
        CK(cudaFree(dY));
    }
    printf("}\n");
    return 0;
}
