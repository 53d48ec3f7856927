// imad_bench.cu -- integer-pipe microbenchmarks for 64-bit Shoup multiplication variants on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/imad_bench tools/imad_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

__device__ __forceinline__ u64 shoup_exact(u64 x, u64 w, u64 ws, u64 q) {
    u64 h = __umul64hi(x, ws);
    return x * w - h * q;
}
// approximate quotient, cross terms through 64-bit products
__device__ __forceinline__ u64 shoup_apx_wide(u64 x, u64 w, u64 ws, u64 q) {
    u32 x0 = (u32)x, x1 = (u32)(x >> 32), s0 = (u32)ws, s1 = (u32)(ws >> 32);
    u64 h = (u64)x1 * s1 + (((u64)x1 * s0) >> 32) + (((u64)x0 * s1) >> 32);
    return x * w - h * q;
}
// approximate quotient, cross terms through IMAD.HI
__device__ __forceinline__ u64 shoup_apx_hi(u64 x, u64 w, u64 ws, u64 q) {
    u32 x0 = (u32)x, x1 = (u32)(x >> 32), s0 = (u32)ws, s1 = (u32)(ws >> 32);
    u64 h = (u64)x1 * s1 + (u64)__umulhi(x1, s0) + (u64)__umulhi(x0, s1);
    return x * w - h * q;
}
// same with negated modulus folded into one multiply-add chain
__device__ __forceinline__ u64 shoup_apx_hi_nq(u64 x, u64 w, u64 ws, u64 nq) {
    u32 x0 = (u32)x, x1 = (u32)(x >> 32), s0 = (u32)ws, s1 = (u32)(ws >> 32);
    u64 h = (u64)x1 * s1 + (u64)__umulhi(x1, s0) + (u64)__umulhi(x0, s1);
    return x * w + h * nq;
}
__device__ __forceinline__ u64 mulhi_only(u64 x, u64 w, u64 ws, u64 q) { return __umul64hi(x, ws) + w; }
__device__ __forceinline__ u64 mullo_only(u64 x, u64 w, u64 ws, u64 q) { return x * w + ws; }
__device__ __forceinline__ u64 wide_only(u64 x, u64 w, u64 ws, u64 q) { return (u64)(u32)x * (u32)w + ws; }
__device__ __forceinline__ u64 hi32_only(u64 x, u64 w, u64 ws, u64 q) { return (u64)__umulhi((u32)x, (u32)w) + ((u64)__umulhi((u32)(x >> 32), (u32)ws) << 32); }
__device__ __forceinline__ u64 lo32_only(u64 x, u64 w, u64 ws, u64 q) { return (u64)((u32)x * (u32)w) | ((u64)((u32)(x >> 32) * (u32)ws) << 32); }

template <int V>
__global__ void k(u64 *out, int iters, u64 q, u64 w, u64 ws) {
    u64 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (u64)threadIdx.x * 977 + i * 31 + blockIdx.x + (1ull << 60);
    u64 nq = 0 - q;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (V == 0) v[i] = shoup_exact(v[i], w, ws, q);
            if (V == 1) v[i] = shoup_apx_wide(v[i], w, ws, q);
            if (V == 2) v[i] = shoup_apx_hi(v[i], w, ws, q);
            if (V == 3) v[i] = shoup_apx_hi_nq(v[i], w, ws, nq);
            if (V == 4) v[i] = mulhi_only(v[i], w, ws, q);
            if (V == 5) v[i] = mullo_only(v[i], w, ws, q);
            if (V == 6) v[i] = wide_only(v[i], w, ws, q);
            if (V == 7) v[i] = hi32_only(v[i], w, ws, q);
            if (V == 8) v[i] = lo32_only(v[i], w, ws, q);
        }
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= v[i];
    if (s == 0x123456789abcdefull) out[0] = s;
}
template <int V>
void run(const char *name, u64 *d) {
    const u64 q = 2305843009211596801ull, w = 1234567890123456789ull % q;
    const u64 ws = (u64)(((unsigned __int128)w << 64) / q);
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<V><<<blocks, threads>>>(d, 64, q, w, ws);
    cudaEventRecord(e0);
    k<V><<<blocks, threads>>>(d, iters, q, w, ws);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * 8.0 * iters;
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cyc_per_warp_op = (ms * 1e-3) * (clk * 1e3) / (ops / 32.0 / (148.0 * 4.0));
    printf("%-18s %8.3f ms  %.3e op/s  %.1f cycles per warp-op per SMSP (at %d MHz)\n", name, ms, ops / (ms * 1e-3), cyc_per_warp_op, clk / 1000);
}
int main() {
    u64 *d;
    cudaMalloc(&d, 64);
    run<0>("shoup_exact", d);
    run<1>("shoup_apx_wide", d);
    run<2>("shoup_apx_hi", d);
    run<3>("shoup_apx_hi_nq", d);
    run<4>("mulhi64", d);
    run<5>("mullo64", d);
    run<6>("imad.wide", d);
    run<7>("2x imad.hi", d);
    run<8>("2x imad.lo", d);
    // correctness of the approximations on random inputs is checked in tests (emulated on the host)
    return cudaDeviceSynchronize() != cudaSuccess;
}
