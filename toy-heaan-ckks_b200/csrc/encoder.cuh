// encoder.cuh -- CkksEncoder on the device (a "next" row of the hot-path table).
//
// The reference evaluates the canonical embedding with an O(N^2) Vandermonde product
// (special_idft / special_dft, special_fft.rs:194-242; encode / decode, ckks_encoder.rs:65-156).
// The sums run over all odd exponents e of psi = exp(i pi / N):
//     encode: coeff[c] = (1/N) sum_s z[N-1-s] * psi^(e_s c),   e_s = 5^s (s < N/2), 2N - 5^(N-1-s) (s >= N/2)
//     decode: slot[s]  = sum_c coeff[c] * psi^(-e_s c), then reversed
// With e = 2m + 1 both are a twist by psi^(+-c) around one length-N complex FFT over m, so the device
// computes them in O(N log N).  Floating-point summation order differs from the reference, hence the
// result is tolerance-checked (not bit-exact), as SURVEY.md 8f.3 prescribes.
#pragma once
#include <cstdint>

struct cplx {
    double re, im;
};
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }

// values: [batch][nvals] complex; buf: [batch][N] complex in bit-reversed order for the DIT stages.
// pow5[h] = 5^h mod 2N.  Slot h contributes conj(v) at e = 5^h and v at e = 2N - 5^h.
__global__ void enc_scatter_kernel(const cplx *__restrict__ values, size_t nvals, double delta, const unsigned *__restrict__ pow5,
                                   cplx *__restrict__ buf, int logn, size_t batch) {
    const size_t half = (size_t)1 << (logn - 1), n = half * 2;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch * half) return;
    size_t b = i / half, h = i % half;
    cplx v = cplx{0.0, 0.0};
    if (h < nvals) {
        v = values[b * nvals + h];
        v.re *= delta;
        v.im *= delta;
    }
    unsigned e = pow5[h];
    unsigned m1 = (e - 1) >> 1, m2 = (unsigned)((2 * n - e - 1) >> 1);
    unsigned r1 = __brev(m1) >> (32 - logn), r2 = __brev(m2) >> (32 - logn);
    buf[b * n + r1] = cplx{v.re, -v.im};
    buf[b * n + r2] = v;
}
// One radix-2 decimation-in-time stage (input bit-reversed): len = 2 * halfsize, sign = +1 / -1.
__global__ void fft_stage_kernel(cplx *__restrict__ buf, int logn, int loghalf, double sign, size_t batch) {
    const size_t n2 = (size_t)1 << (logn - 1);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch * n2) return;
    size_t b = i / n2, j = i % n2;
    size_t half = (size_t)1 << loghalf;
    size_t blk = j >> loghalf, off = j & (half - 1);
    size_t i0 = (b << logn) + (blk << (loghalf + 1)) + off;
    double s, c;
    sincospi(sign * (double)off / (double)half, &s, &c);
    cplx u = buf[i0], t = cmul(buf[i0 + half], cplx{c, s});
    buf[i0] = cplx{u.re + t.re, u.im + t.im};
    buf[i0 + half] = cplx{u.re - t.re, u.im - t.im};
}
// coeff[c] = round( Re( psi^c * X[c] ) / N )   (f64::round: half away from zero, like C round())
__global__ void enc_finish_kernel(const cplx *__restrict__ buf, long long *__restrict__ coeffs, int logn, size_t batch) {
    const size_t n = (size_t)1 << logn;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch * n) return;
    size_t c = i & (n - 1);
    double s, co;
    sincospi((double)c / (double)n, &s, &co);
    cplx x = buf[i];
    double re = (x.re * co - x.im * s) / (double)n;
    coeffs[i] = (long long)round(re);
}
// y[brv(c)] = coeff[c] * psi^(-c)
__global__ void dec_twist_kernel(const long long *__restrict__ coeffs, cplx *__restrict__ buf, int logn, size_t batch) {
    const size_t n = (size_t)1 << logn;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch * n) return;
    size_t b = i >> logn, c = i & (n - 1);
    double s, co;
    sincospi(-(double)c / (double)n, &s, &co);
    double a = (double)coeffs[i];
    unsigned r = __brev((unsigned)c) >> (32 - logn);
    buf[(b << logn) + r] = cplx{a * co, a * s};
}
// the same for coefficients that arrive as doubles (centred CRT of a basis with Q >= 2^128, crt_wide.cuh)
__global__ void dec_twist_f64_kernel(const double *__restrict__ coeffs, cplx *__restrict__ buf, int logn, size_t batch) {
    const size_t n = (size_t)1 << logn;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch * n) return;
    size_t b = i >> logn, c = i & (n - 1);
    double s, co;
    sincospi(-(double)c / (double)n, &s, &co);
    double a = coeffs[i];
    unsigned r = __brev((unsigned)c) >> (32 - logn);
    buf[(b << logn) + r] = cplx{a * co, a * s};
}
// out[b][i] = Y[m_i] / delta,  m_i = (2N - 5^i - 1) / 2   (the reversed slot order of special_dft)
__global__ void dec_gather_kernel(const cplx *__restrict__ buf, const unsigned *__restrict__ pow5, double inv_delta, cplx *__restrict__ out,
                                  size_t nslots, int logn, size_t batch) {
    const size_t n = (size_t)1 << logn;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch * nslots) return;
    size_t b = i / nslots, s = i % nslots;
    unsigned m = (unsigned)((2 * n - pow5[s] - 1) >> 1);
    cplx y = buf[(b << logn) + m];
    out[i] = cplx{y.re * inv_delta, y.im * inv_delta};
}
