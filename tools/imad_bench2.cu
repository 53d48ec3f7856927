// imad_bench2.cu -- hand-scheduled PTX variants of the 64-bit Shoup multiply and the lazy CT butterfly.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

__device__ __forceinline__ u64 shoup_ref(u64 x, u64 w, u64 ws, u64 q) {
    u64 h = __umul64hi(x, ws);
    return x * w - h * q;
}
// approximate quotient (drops x0*s0 and the low halves of the cross terms: h' in [h-2, h]) and a
// multiply-add chain with the negated modulus: r = x*w + h'*(-q) mod 2^64, r in [0, 4q)
__device__ __forceinline__ u64 shoup_ptx(u64 x, u64 w, u64 ws, u64 nq) {
    u32 x0, x1, w0, w1, s0, s1, n0, n1;
    asm("mov.b64 {%0,%1}, %2;" : "=r"(x0), "=r"(x1) : "l"(x));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(w0), "=r"(w1) : "l"(w));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(s0), "=r"(s1) : "l"(ws));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(n0), "=r"(n1) : "l"(nq));
    u64 t, u, h, r;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(x1), "r"(s0));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(u) : "r"(x0), "r"(s1));
    u32 tl, th, ul, uh, mid, cy;
    asm("mov.b64 {%0,%1}, %2;" : "=r"(tl), "=r"(th) : "l"(t));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(ul), "=r"(uh) : "l"(u));
    asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=r"(mid), "=r"(cy) : "r"(th), "r"(uh));
    u64 add;
    asm("mov.b64 %0, {%1,%2};" : "=l"(add) : "r"(mid), "r"(cy));
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(h) : "r"(x1), "r"(s1), "l"(add));
    u32 h0, h1;
    asm("mov.b64 {%0,%1}, %2;" : "=r"(h0), "=r"(h1) : "l"(h));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(x0), "r"(w0));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r) : "r"(h0), "r"(n0));
    u32 r0, r1;
    asm("mov.b64 {%0,%1}, %2;" : "=r"(r0), "=r"(r1) : "l"(r));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r1) : "r"(x1), "r"(w0));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r1) : "r"(x0), "r"(w1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r1) : "r"(h1), "r"(n0));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r1) : "r"(h0), "r"(n1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(r0), "r"(r1));
    return r;
}
// exact quotient, multiply-add chain with negated modulus
__device__ __forceinline__ u64 shoup_nq(u64 x, u64 w, u64 ws, u64 nq) {
    u64 h = __umul64hi(x, ws);
    return x * w + h * nq;
}
__device__ __forceinline__ u64 csub(u64 x, u64 q) { return x >= q ? x - q : x; }

template <int V>
__global__ void k(u64 *out, int iters, u64 q, u64 w, u64 ws) {
    u64 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (u64)threadIdx.x * 977 + i * 31 + blockIdx.x + (1ull << 60);
    const u64 nq = 0 - q, q2 = 2 * q, q4 = 4 * q;
    for (int it = 0; it < iters; ++it) {
        if (V == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = shoup_ref(v[i], w, ws, q);
        } else if (V == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = shoup_ptx(v[i], w, ws, nq);
        } else if (V == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = shoup_nq(v[i], w, ws, nq);
        } else if (V == 3) {  // lazy CT butterflies [0,4q), exact Shoup: 4 butterflies on 8 values
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                u64 x = csub(v[i], q2), t = shoup_ref(v[i + 4], w, ws, q);
                v[i] = x + t;
                v[i + 4] = x - t + q2;
            }
        } else if (V == 4) {  // lazy CT butterflies [0,8q), approximate Shoup (q < 2^61)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                u64 x = csub(v[i], q4), t = shoup_ptx(v[i + 4], w, ws, nq);
                v[i] = x + t;
                v[i + 4] = x - t + q4;
            }
        }
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= v[i];
    if (s == 0x123456789abcdefull) out[0] = s;
}
template <int V>
void run(const char *name, u64 *d, double per_iter) {
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
    double ops = (double)blocks * threads * per_iter * iters;
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cyc = (ms * 1e-3) * (clk * 1e3) / (ops / 32.0 / (148.0 * 4.0));
    printf("%-26s %8.3f ms  %.3e op/s  %.1f cycles per warp-op per SMSP\n", name, ms, ops / (ms * 1e-3), cyc);
}
__global__ void check(u64 *bad, u64 q, u64 w, u64 ws) {
    u64 nq = 0 - q;
    u64 x = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    for (int i = 0; i < 64; ++i) {
        x = x * 6364136223846793005ull + 1442695040888963407ull;
        u64 r = shoup_ptx(x, w, ws, nq);
        unsigned __int128 p = (unsigned __int128)x * w;
        u64 exact = (u64)(p % q);
        if (r % q != exact || r >= 4 * q) atomicAdd((unsigned long long *)bad, 1ull);
    }
}
int main() {
    u64 *d;
    cudaMalloc(&d, 64);
    cudaMemset(d, 0, 64);
    const u64 q = 2305843009211596801ull, w = 1234567890123456789ull % q;
    const u64 ws = (u64)(((unsigned __int128)w << 64) / q);
    check<<<1024, 256>>>(d, q, w, ws);
    u64 bad = 1;
    cudaMemcpy(&bad, d, 8, cudaMemcpyDeviceToHost);
    printf("shoup_ptx mismatches or range violations (want 0): %llu\n", bad);
    run<0>("shoup exact (compiler)", d, 8);
    run<1>("shoup approx (PTX chain)", d, 8);
    run<2>("shoup exact + nq chain", d, 8);
    run<3>("CT bfly [0,4q) exact", d, 4);
    run<4>("CT bfly [0,8q) approx PTX", d, 4);
    return cudaDeviceSynchronize() != cudaSuccess;
}
