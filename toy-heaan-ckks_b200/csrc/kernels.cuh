// kernels.cuh -- CUDA kernels (sm_100a) of the RNS-NTT hot path.  Each kernel cites the reference
// code whose effect it reproduces (paths relative to the reference repository root).
#pragma once
#include <type_traits>

#include "ntt_tile.cuh"

// Data layout everywhere: [batch][limb][N] u64, limb-major like the reference's Vec<[u64; N]>
// (poly.rs:26-30).  Coefficient-domain words are in natural order.  NTT-domain words are in the
// device-internal order of the context's NTT path (see permute_ntt_kernel).

// =================================================================================================
// Four-step NTT passes
// =================================================================================================
struct PassArgs {
    const void *src;      // u64 words, or the internal word type WD for the second pass of a transform
    void *dst;            // WD for a transposed (internal) store, u64 otherwise
    const LimbConst *lc;  // [L]
    const void *tab;      // small per-limb twiddle table (TwOf<WD>), tab_stride entries per limb
    const void *elt;      // per-element table (N entries per limb), or null
    size_t tab_stride;
    int L;          // limbs per polynomial in src
    unsigned ncols;  // columns (= stride of the transform dimension, in words)
    size_t N;
    int limb0;      // first limb handled (blockIdx.y counts from here)
    int dstL;       // limbs per polynomial in dst
    int dst_limb0;  // dst limb index = limb - dst_limb0
    // MULTI stores (limb-sharded mode): the finished coefficient-domain limb is written into the gather
    // buffers of `npeer` GPUs (own + NVLink peers) at slot (m_off + m_step * limb), layout [slot][m_cs][N].
    u64 *peer[8];
    int npeer, m_first;
    int m_off, m_step;
    size_t m_cs;
    // ADDROT (rotate_ciphertext, engine.rs:417-419 + :454): the finished coefficient-domain word at position p gets
    // automorphism(rot_src)[p] added, gathered on the fly: +-rot_src[p * rot_einv mod 2N] (poly.rs:515-538).
    const u64 *rot_src;  // [batch][L][N] coefficient domain
    u64 rot_einv;        // inverse of the (odd) Galois exponent modulo 2N
    // IO32 (auxiliary-basis key-switch, aux_ks.cuh): the "limbs" of src are (ciphertext limb, auxiliary prime) pairs
    // and the tables of limb l are those of prime (l / tab_div) % tab_mod
    int tab_div, tab_mod;
};

// One pass: a 2^A-point transform along the strided dimension of a [2^A][ncols] limb, for a tile of
// C adjacent columns.  grid = (ncols / C, L, batch), block = C * 2^(A-E).
//   PREMUL   : multiply by elt[idx][col] on load   (four-step twiddle, forward)
//   POSTMUL  : multiply by elt[idx][col] on store  (four-step twiddle and 1/N, inverse)
//   TRANSPOSE: store the tile transposed, dst[col][idx]  (lazy values, consumed by the next pass)
//              otherwise store in place dst[idx][col] as canonical representatives.
// Forward  to_ntt_domain  (poly.rs:136-148, 574-580) = <NEG_FWD,TRANSPOSE> then <CYC_FWD,PREMUL>.
// Inverse  to_coeff_domain(poly.rs:154-166, 582-591) = <CYC_INV,POSTMUL,TRANSPOSE> then <NEG_INV>.
// WD = u64 (any q < 2^63) or u32 (all q < 2^31: 32-bit butterflies, 32-bit internal scratch; the words
// that cross the boundary stay u64).
// FIXLOGN != 0: the ring degree is the compile-time constant 2^FIXLOGN (the launchers pick it for N = 2^16 and 2^14):
// ncols and N fold into the load / store immediates, which removes the per-access 64-bit address arithmetic --
// 10 % of the instructions of a 64-bit pass, 10 - 28 % of a 32-bit one (cuobjdump counts).
// IO32: the words crossing the boundary are of the transform word type too (u32 in, u32 out).
template <typename WD, int KIND, int A, int E, int C, int LAZY, bool PREMUL, bool POSTMUL, bool TRANSPOSE, bool MULTI = false, bool ADDROT = false,
          int FIXLOGN = 0, bool IO32 = false>
__global__ void __launch_bounds__(C *(1 << (A - E))) ntt_pass_kernel(PassArgs a) {
    if (FIXLOGN) {
        a.ncols = 1u << (FIXLOGN > A ? FIXLOGN - A : 0);
        a.N = (size_t)1 << FIXLOGN;
    }
    typedef TileGeom<A, E> GM;
    typedef typename TwOf<WD>::type TW;
    constexpr int CP = C + 1;
    constexpr int NT = C * GM::G;
    constexpr bool FWD = (KIND == XF_NEG_FWD || KIND == XF_CYC_FWD);
    constexpr bool SRC_INTERNAL = (KIND == XF_CYC_FWD || KIND == XF_NEG_INV);
    typedef typename std::conditional<SRC_INTERNAL || IO32, WD, u64>::type SRC_T;
    typedef typename std::conditional<TRANSPOSE || IO32, WD, u64>::type DST_T;
    extern __shared__ __align__(16) unsigned char sm_raw[];
    WD *sm = reinterpret_cast<WD *>(sm_raw);
    const SRC_T *src = reinterpret_cast<const SRC_T *>(a.src);
    DST_T *dst = reinterpret_cast<DST_T *>(a.dst);
    const int tid = threadIdx.x;
    const int c = tid % C, g = tid / C;
    const int limb = blockIdx.y + a.limb0;
    const size_t c0 = (size_t)blockIdx.x * C;
    const size_t base = ((size_t)blockIdx.z * a.L + limb) * a.N;
    const size_t dbase = ((size_t)blockIdx.z * a.dstL + (limb - a.dst_limb0)) * a.N;
    const int tl = IO32 ? (limb / a.tab_div) % a.tab_mod : limb;  // whose tables
    const LimbConst m = a.lc[tl];
    const WD q = (WD)m.q, q2 = (WD)m.q2;
    const TW *tab = reinterpret_cast<const TW *>(a.tab) + (size_t)tl * a.tab_stride;
    const TW *elt = (PREMUL || POSTMUL) ? reinterpret_cast<const TW *>(a.elt) + (size_t)tl * a.N : nullptr;

    WD v[1 << E];
    constexpr int lo_in = FWD ? GM::lo(0) : GM::lo(GM::NS - 1);
    constexpr int lo_out = FWD ? GM::lo(GM::NS - 1) : GM::lo(0);
#pragma unroll
    for (int k = 0; k < (1 << E); ++k) {
        size_t off = (size_t)tile_idx<E>(g, k, lo_in) * a.ncols + c0 + c;
        WD x = (WD)src[base + off];
        if (PREMUL) x = mul_tw<LAZY>(x, ldg_tw(elt + off), q);
        v[k] = x;
    }
    xf_tile<KIND, A, E, CP, LAZY>(v, g, c, sm, tab, q, q2);
    constexpr bool CT_RANGE = (KIND == XF_NEG_FWD || KIND == XF_CYC_INV);
    if (POSTMUL) {
#pragma unroll
        for (int k = 0; k < (1 << E); ++k) {
            size_t off = (size_t)tile_idx<E>(g, k, lo_out) * a.ncols + c0 + c;
            v[k] = mul_tw<LAZY>(v[k], ldg_tw(elt + off), q);
        }
    }
    if (TRANSPOSE) {
        // (no barrier before this put: it overwrites exactly the slots this thread read in the transform's last
        // tile_get, which was at the same window lo_out)
        tile_put<E, CP>(sm, v, g, c, lo_out);
        __syncthreads();
        DST_T *d = dst + dbase + c0 * (size_t)(1 << A);
#pragma unroll 4
        for (int i = 0; i < (1 << E); ++i) {  // (C << A) / NT = 2^E words per thread (a full unroll costs 16 registers)
            const int e = tid + i * NT;
            int cc = e >> A, r = e & ((1 << A) - 1);
            d[e] = (DST_T)sm[r * CP + cc];
        }
    } else {
#pragma unroll
        for (int k = 0; k < (1 << E); ++k) {
            size_t off = (size_t)tile_idx<E>(g, k, lo_out) * a.ncols + c0 + c;
            WD x = v[k];
            if (POSTMUL || !CT_RANGE) x = canon2<LAZY>(x, q);
            else x = canon4<LAZY>(x, q, q2);
            if (ADDROT) {
                u64 sidx = ((u64)off * a.rot_einv) & (2 * a.N - 1);
                const bool neg = sidx >= a.N;  // then the source coefficient lands on p + N: negated
                if (neg) sidx -= a.N;
                WD y = (WD)a.rot_src[base + sidx];
                if (neg && y) y = q - y;
                x = csub((WD)(x + y), q);
            }
            if (MULTI) {
                // all-gather fused into the producing pass: plain stores into own and peer HBM
                // (each GPU starts with a different peer so that no destination is hit by everyone at once)
                const size_t mo = ((size_t)(a.m_off + a.m_step * limb) * a.m_cs + blockIdx.z) * a.N + off;
                for (int t = 0; t < a.npeer; ++t) {
                    int p = a.m_first + t;
                    if (p >= a.npeer) p -= a.npeer;
                    a.peer[p][mo] = (u64)x;
                }
            } else {
                dst[dbase + off] = (DST_T)x;
            }
        }
    }
}

// =================================================================================================
// Small-N NTT: one CTA per limb, whole limb in shared memory (N <= 2048).
// to_ntt_domain / to_coeff_domain (poly.rs:136-166) for the reference's test sizes (N = 8, 16, ...).
// Internal NTT order: bit-reversed (position brv(k) holds slot k).
// =================================================================================================
struct SmallArgs {
    u64 *data;  // in place
    const LimbConst *lc;
    const tw_t *psi;   // [L][N] psi^brv(i) (forward) or psi^-brv(i) (inverse)
    const tw_t *ninv;  // [L] N^-1 (inverse only)
    int L;
    int logn;
};

template <bool INVERSE, int LAZY>
__global__ void ntt_small_kernel(SmallArgs a) {
    extern __shared__ u64 sm[];
    const int n = 1 << a.logn;
    const int limb = blockIdx.x % a.L;
    u64 *d = a.data + (size_t)blockIdx.x * n;
    const LimbConst m = a.lc[limb];
    const u64 q = m.q, q2 = m.q2;
    const tw_t *P = a.psi + (size_t)limb * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = d[i];
    __syncthreads();
    if (!INVERSE) {
        for (int s = 0; s < a.logn; ++s) {
            const int t = n >> (s + 1);
            for (int j = threadIdx.x; j < n / 2; j += blockDim.x) {
                int blk = j / t, off = j % t;
                int i0 = blk * 2 * t + off;
                u64 x = sm[i0], y = sm[i0 + t];
                ct_bfly<LAZY>(x, y, ldg_tw(P + (1 << s) + blk), q, q2);
                sm[i0] = x;
                sm[i0 + t] = y;
            }
            __syncthreads();
        }
        for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = canon4<LAZY>(sm[i], q, q2);
    } else {
        for (int s = a.logn - 1; s >= 0; --s) {
            const int t = n >> (s + 1);
            for (int j = threadIdx.x; j < n / 2; j += blockDim.x) {
                int blk = j / t, off = j % t;
                int i0 = blk * 2 * t + off;
                u64 x = sm[i0], y = sm[i0 + t];
                gs_bfly<LAZY>(x, y, ldg_tw(P + (1 << s) + blk), q, q2);
                sm[i0] = x;
                sm[i0 + t] = y;
            }
            __syncthreads();
        }
        const tw_t ni = ldg_tw(a.ninv + limb);
        for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = shoup(sm[i], ni, q);
    }
}

// =================================================================================================
// Layout of NTT-domain data crossing the boundary (channels(), from_channels(.., is_ntt=true)):
// the reference's slot k = p(psi^(2k+1)) in natural order (poly.rs:136-148).
//   a1 == logn (small path): internal position brv_logn(k)
//   four-step (n1 = 2^a1, n2 = 2^a2): internal position brv_a2(k >> a1) * n1 + brv_a1(k mod n1)
// =================================================================================================
__device__ __forceinline__ unsigned ntt_pos(unsigned k, int a1, int a2) {
    unsigned k1 = k & ((1u << a1) - 1), k2 = k >> a1;
    unsigned r1 = a1 ? (__brev(k1) >> (32 - a1)) : 0;
    unsigned r2 = a2 ? (__brev(k2) >> (32 - a2)) : 0;
    return (r2 << a1) | r1;
}
__global__ void permute_ntt_kernel(const u64 *__restrict__ src, u64 *__restrict__ dst, size_t total, int logn, int a1,
                                   int a2, int to_internal) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    size_t poly = i >> logn;
    unsigned k = (unsigned)(i & (((size_t)1 << logn) - 1));
    unsigned p = ntt_pos(k, a1, a2);
    if (to_internal) dst[(poly << logn) + p] = src[i];
    else dst[i] = src[(poly << logn) + p];
}

// =================================================================================================
// Elementwise limb kernels.  i indexes [batch][L][N]; rhs may be broadcast over the batch
// (rhs_bstride == 0).
// =================================================================================================
struct EwArgs {
    const LimbConst *lc;
    size_t total;  // batch * L * N
    size_t poly;   // L * N
    int logn;
    int L;
};
__device__ __forceinline__ int ew_limb(const EwArgs &a, size_t i) { return (int)((i % a.poly) >> a.logn); }

enum { EW_ADD = 0, EW_SUB = 1, EW_MUL = 2 };
// AddAssign (poly.rs:254-275), a += -b, pointwise MulAssign in the NTT domain (poly.rs:297-306).
template <int OP>
__global__ void ew_binary_kernel(EwArgs a, u64 *__restrict__ x, const u64 *__restrict__ y, size_t y_bstride) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (size_t)gridDim.x * blockDim.x) {
        const LimbConst &m = a.lc[ew_limb(a, i)];
        size_t b = i / a.poly, r = i % a.poly;
        u64 yy = y[b * y_bstride + r];
        u64 xx = x[i];
        if (OP == EW_ADD) x[i] = addmod(xx, yy, m.q);
        if (OP == EW_SUB) x[i] = submod(xx, yy, m.q);
        if (OP == EW_MUL) x[i] = mulmod(xx, yy, m);
    }
}
// add_ciphertexts (engine.rs:131-151) out of place: z = x + y, one pass (read two polynomials, write one) instead of
// clone + AddAssign (read three, write two).
__global__ void ew_add3_kernel(EwArgs a, const u64 *__restrict__ x, const u64 *__restrict__ y, u64 *__restrict__ z) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (size_t)gridDim.x * blockDim.x)
        z[i] = addmod(x[i], y[i], a.lc[ew_limb(a, i)].q);
}
// Neg (poly.rs:370-385)
__global__ void ew_neg_kernel(EwArgs a, u64 *__restrict__ x) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (size_t)gridDim.x * blockDim.x)
        x[i] = negmod(x[i], a.lc[ew_limb(a, i)].q);
}
// from_coeffs (poly.rs:49-66): rem_euclid of an i64 per limb.  coeffs: [batch][clen].
__global__ void from_coeffs_kernel(EwArgs a, const i64 *__restrict__ coeffs, size_t clen, u64 *__restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (size_t)gridDim.x * blockDim.x) {
        const LimbConst &m = a.lc[ew_limb(a, i)];
        size_t b = i / a.poly;
        size_t k = i & (((size_t)1 << a.logn) - 1);
        i64 c = coeffs[b * clen + k];
        u64 mag = c < 0 ? (u64)0 - (u64)c : (u64)c;
        u64 r = barrett_word(mag, m);
        out[i] = (c < 0) ? negmod(r, m.q) : r;
    }
}
// from_channels reducedness scan (poly.rs:83-93): flag = 1 if any word >= its modulus.
__global__ void check_reduced_kernel(EwArgs a, const u64 *__restrict__ x, int *flag) {
    int bad = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (size_t)gridDim.x * blockDim.x)
        bad |= (x[i] >= a.lc[ew_limb(a, i)].q);
    if (bad) atomicOr(flag, 1);
}
// Tensor product of mul_ciphertexts_gadget (engine.rs:481-493) on NTT-domain operands:
// d0 = a0*b0, d1 = a0*b1 + a1*b0, d2 = a1*b1.  d0/d1/d2 may alias a0/a1/b0 element-wise.
__device__ __forceinline__ void tensor_one(const LimbConst &m, u64 x0, u64 x1, u64 y0, u64 y1, u64 &t0, u64 &t1, u64 &t2) {
    t0 = mulmod(x0, y0, m);
    t2 = mulmod(x1, y1, m);
    // x0 y1 + x1 y0 as one 128-bit sum (both products are below 2^126: q < 2^63), reduced once
    const u64 l1 = x0 * y1, l2 = x1 * y0;
    const u64 lo = l1 + l2;
    const u64 hi = __umul64hi(x0, y1) + __umul64hi(x1, y0) + (lo < l1 ? 1ull : 0ull);
    t1 = reduce128(hi, lo, m);
}
// Two adjacent words per thread and iteration (16-byte loads and stores: the kernel is HBM-bound, four streams in and three
// out); a limb has an even number of words, so both words of a pair share their modulus.
__global__ void tensor_kernel(EwArgs a, const u64 *a0, const u64 *a1, const u64 *b0, const u64 *b1, u64 *d0, u64 *d1,
                              u64 *d2) {
    if (a.logn == 0) {  // N = 1: word by word
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (size_t)gridDim.x * blockDim.x) {
            u64 t0, t1, t2;
            tensor_one(a.lc[ew_limb(a, i)], a0[i], a1[i], b0[i], b1[i], t0, t1, t2);
            d0[i] = t0;
            d1[i] = t1;
            d2[i] = t2;
        }
        return;
    }
    const ulonglong2 *A0 = reinterpret_cast<const ulonglong2 *>(a0), *A1 = reinterpret_cast<const ulonglong2 *>(a1);
    const ulonglong2 *B0 = reinterpret_cast<const ulonglong2 *>(b0), *B1 = reinterpret_cast<const ulonglong2 *>(b1);
    ulonglong2 *D0 = reinterpret_cast<ulonglong2 *>(d0), *D1 = reinterpret_cast<ulonglong2 *>(d1), *D2 = reinterpret_cast<ulonglong2 *>(d2);
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.total / 2; p += (size_t)gridDim.x * blockDim.x) {
        const LimbConst &m = a.lc[ew_limb(a, 2 * p)];
        const ulonglong2 x0 = A0[p], x1 = A1[p], y0 = B0[p], y1 = B1[p];
        ulonglong2 t0, t1, t2;
        tensor_one(m, x0.x, x1.x, y0.x, y1.x, t0.x, t1.x, t2.x);
        tensor_one(m, x0.y, x1.y, y0.y, y1.y, t0.y, t1.y, t2.y);
        D0[p] = t0;
        D1[p] = t1;
        D2[p] = t2;
    }
}
// rescale_into (poly.rs:214-225): out_i = (c_i - (c_last % q_i)) * (q_last^-1 mod q_i) mod q_i.
// src: [batch][L][N] coefficient domain; out: [batch][L-1][N]; qlinv: [L-1] Shoup pairs.
__global__ void rescale_kernel(EwArgs a /* of the OUTPUT: L-1 limbs */, const u64 *__restrict__ src,
                               u64 *__restrict__ out, const tw_t *__restrict__ qlinv) {
    const size_t n = (size_t)1 << a.logn;
    const size_t in_poly = a.poly + n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (size_t)gridDim.x * blockDim.x) {
        int limb = ew_limb(a, i);
        const LimbConst &m = a.lc[limb];
        size_t b = i / a.poly, r = i % a.poly, k = r & (n - 1);
        u64 ci = src[b * in_poly + r];
        u64 cl = barrett_word(src[b * in_poly + a.poly + k], m);
        out[i] = shoup(submod(ci, cl, m.q), ldg_tw(qlinv + limb), m.q);
    }
}
// mul_assign_naive (poly.rs:339-367): the reference's O(N^2) schoolbook product in Z_q[X]/(X^N + 1), kept there as
// the correctness yardstick of the NTT path and kept here for the same purpose.  One thread per output
// coefficient: out[k] = sum_{i<=k} a[i] b[k-i] - sum_{i>k} a[i] b[N+k-i]  (X^N = -1).  y_bstride = 0 broadcasts rhs.
__global__ void mul_naive_kernel(EwArgs a, const u64 *__restrict__ x, const u64 *__restrict__ y, size_t y_bstride, u64 *__restrict__ out) {
    const size_t n = (size_t)1 << a.logn;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < a.total; t += (size_t)gridDim.x * blockDim.x) {
        const LimbConst &m = a.lc[ew_limb(a, t)];
        const size_t k = t & (n - 1), b = t / a.poly, limb_off = (t % a.poly) - k;
        const u64 *xa = x + t - k, *yb = y + b * y_bstride + limb_off;
        u64 pos = 0, neg = 0;
        for (size_t i = 0; i <= k; ++i) pos = mulmod_add(xa[i], yb[k - i], pos, m);
        for (size_t i = k + 1; i < n; ++i) neg = mulmod_add(xa[i], yb[n + k - i], neg, m);
        out[t] = submod(pos, neg, m.q);
    }
}

// automorphism (poly.rs:515-538) for odd exponents (a signed permutation), as a gather:
// out[j] = +-in[i] with i = j * e^-1 mod 2N (sign from i*e mod 2N >= N).
__global__ void automorphism_kernel(EwArgs a, const u64 *__restrict__ src, u64 *__restrict__ out, u64 e, u64 einv) {
    const u64 n = (u64)1 << a.logn, mask2 = 2 * n - 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (size_t)gridDim.x * blockDim.x) {
        u64 q = a.lc[ew_limb(a, i)].q;
        u64 j = i & (n - 1);
        // source index s in [0, N) with s*e == j or j+N (mod 2N)
        u64 s = (j * einv) & mask2;
        bool neg = false;
        if (s >= n) {  // then (s-n)*e == j + N (mod 2N) because e is odd
            s -= n;
            neg = true;
        }
        (void)e;
        u64 v = src[i - j + s];
        out[i] = neg ? negmod(v, q) : v;
    }
}
// automorphism for even exponents (not a bijection): the reference scatters in increasing i and
// skips zero coefficients, so the largest i with a non-zero coefficient wins (poly.rs:523-537).
__global__ void automorphism_even_mark_kernel(EwArgs a, const u64 *__restrict__ src, unsigned *__restrict__ winner,
                                              u64 e) {
    const u64 n = (u64)1 << a.logn, mask2 = 2 * n - 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (size_t)gridDim.x * blockDim.x) {
        u64 k = i & (n - 1);
        if (src[i] == 0) continue;
        u64 j = ((k * e) & mask2) & (n - 1);
        atomicMax(winner + (i - k + j), (unsigned)k + 1u);
    }
}
__global__ void automorphism_even_fill_kernel(EwArgs a, const u64 *__restrict__ src, const unsigned *__restrict__ winner,
                                              u64 *__restrict__ out, u64 e) {
    const u64 n = (u64)1 << a.logn, mask2 = 2 * n - 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (size_t)gridDim.x * blockDim.x) {
        u64 q = a.lc[ew_limb(a, i)].q;
        u64 j = i & (n - 1);
        unsigned w = winner[i];
        if (!w) {
            out[i] = 0;
            continue;
        }
        u64 k = w - 1;
        u64 v = src[i - j + k];
        out[i] = (((k * e) & mask2) >= n) ? q - v : v;
    }
}

// =================================================================================================
// Gadget key-switch, unfused building blocks (engine.rs:505-528 / :429-452)
// =================================================================================================
// alpha_i: limb `digit` of src broadcast to every limb j with `% q_j` (engine.rs:507-516).
__global__ void digit_broadcast_kernel(EwArgs a, const u64 *__restrict__ src, u64 *__restrict__ out, int digit) {
    const size_t n = (size_t)1 << a.logn;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (size_t)gridDim.x * blockDim.x) {
        int limb = ew_limb(a, i);
        size_t b = i / a.poly, k = i & (n - 1);
        u64 v = src[b * a.poly + (size_t)digit * n + k];
        out[i] = (limb == digit) ? v : barrett_word(v, a.lc[limb]);
    }
}
// (perm_row: ntt_tile.cuh)
template <typename KT>
__global__ void key_permute_kernel(const u64 *__restrict__ src, KT *__restrict__ dst, size_t total, int logn, int a1, int a2, int pe) {
    const size_t n = (size_t)1 << logn;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t k = i & (n - 1), row = k >> a1, col = k & (((size_t)1 << a1) - 1);
        dst[i - k + (perm_row(row, a2, pe) << a1) + col] = (KT)src[i];
    }
}
// acc0 += x * kb, acc1 += x * ka (all NTT domain); kb/ka: one digit's key polynomial [L][N] (rows permuted
// when pe >= 0).
__global__ void ks_mac_kernel(EwArgs a, const u64 *__restrict__ x, const void *__restrict__ kb_, const void *__restrict__ ka_,
                              u64 *__restrict__ acc0, u64 *__restrict__ acc1, int a1, int a2, int pe, int k32) {
    const size_t n = (size_t)1 << a.logn;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (size_t)gridDim.x * blockDim.x) {
        const LimbConst &m = a.lc[ew_limb(a, i)];
        size_t r = i % a.poly;
        if (pe >= 0) {
            const size_t k = r & (n - 1);
            r = r - k + (perm_row(k >> a1, a2, pe) << a1) + (k & (((size_t)1 << a1) - 1));
        }
        u64 xx = x[i];
        const u64 kbv = k32 ? (u64)__ldg(reinterpret_cast<const u32 *>(kb_) + r) : __ldg(reinterpret_cast<const u64 *>(kb_) + r);
        const u64 kav = k32 ? (u64)__ldg(reinterpret_cast<const u32 *>(ka_) + r) : __ldg(reinterpret_cast<const u64 *>(ka_) + r);
        acc0[i] = mulmod_add(xx, kbv, acc0[i], m);
        acc1[i] = mulmod_add(xx, kav, acc1[i], m);
    }
}

// =================================================================================================
// Integer-pipe microbenchmark: independent Shoup modmuls (measures the IMAD roof the key-switch
// is bound by; MEASURED_PEAKS.json has no integer figure).
// =================================================================================================
__global__ void modmul_peak_kernel(u64 *out, int iters, u64 q, tw_t t) {
    u64 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (u64)threadIdx.x * 977 + k * 31 + blockIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = shoup_lazy(v[k], t, q);
    }
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s ^= v[k];
    if (s == 0x123456789abcdefull) out[0] = s;
}

// =================================================================================================
// Fused gadget key-switch for the four-step path (engine.rs:505-528 / :429-452).
//
//   ks_pass1: for every (ciphertext, digit i, target limb j != i): first pass of NTT_j(alpha_i),
//             reading limb i of the coefficient-domain polynomial directly (the `% q_j` of
//             engine.rs:507-516 is folded into the load, or skipped when 4 q_j > q_i lets the lazy
//             butterflies absorb it) and writing the transposed intermediate to `scratch`.
//   ks_pass2: for every (ciphertext, target limb j, tile): loops over the digits i, finishes
//             NTT_j(alpha_i) in registers, multiplies by key_b[i][j] and key_a[i][j] and accumulates
//             both sums in 128-bit registers (one reduction per ~L digits); digit i == j is the
//             NTT-domain limb itself.  The epilogue adds d0 / d1 (or nothing for rotations), runs the
//             first pass of the inverse transform on the sums while they are still in registers and
//             stores the transposed intermediate for ntt_inv_pass1.
// Nothing but the digit polynomial, the keys and the result crosses HBM more than once.
// =================================================================================================
struct KsArgs {
    const u64 *digits;   // [cts][L][N] coefficient domain (d2 or the rotated c1)
    const u64 *dig_ntt;  // [cts][L][N] NTT domain of the same polynomial (digit i == j shortcut)
    void *scratch;       // [cts][L(j)][L(i)][N] transposed pass-1 output, internal word type WD
    const void *key_b, *key_a;  // [L(i)][L(j)][N] NTT domain, words of the transform type (u32 on the 32-bit path)
    const u64 *add0, *add1;    // [cts][L][N] NTT domain addends (d0, d1) or null
    void *out0, *out1;         // [cts][L][N] transposed inverse-pass-2 output, internal word type WD
    const LimbConst *lc;
    const void *P1, *W2, *W2i, *TTt, *TTi;  // TwOf<WD> tables
    size_t w2_stride;
    int L;   // target limbs held here (all of them, or this GPU's share in limb-sharded mode)
    int Ld;  // digits = limbs of the whole basis (== L unless limb-sharded)
    int joff, jstep;  // basis index of local target limb j: joff + jstep * j  (0, 1 unless limb-sharded)
    int j0;           // first local target limb of this launch (the grid covers j0 .. j0 + nj - 1)
    size_t dig_ct_stride, dig_limb_stride;  // words between ciphertexts / digits in `digits`
    int a1, a2;
    int reduce_every;  // digits between 128-bit accumulator reductions
    size_t N;
};

template <typename WD, int A, int E, int C, int LAZY, bool REDUCE, bool DIAG, int FIXLOGN = 0>
__global__ void __launch_bounds__(C *(1 << (A - E))) ks_pass1_kernel(KsArgs a) {
    if (FIXLOGN) {  // compile-time ring degree (see ntt_pass_kernel)
        a.a1 = A;
        a.a2 = FIXLOGN > A ? FIXLOGN - A : 0;
        a.N = (size_t)1 << FIXLOGN;
    }
    typedef TileGeom<A, E> GM;
    typedef typename TwOf<WD>::type TW;
    constexpr int CP = C + 1;
    constexpr int NT = C * GM::G;
    extern __shared__ __align__(16) unsigned char sm_raw[];
    WD *sm = reinterpret_cast<WD *>(sm_raw);
    const int L = a.L, Ld = a.Ld;
    const int j = a.j0 + blockIdx.y / Ld, i = blockIdx.y % Ld;
    if (DIAG && i == a.joff + a.jstep * j) return;  // ks_pass2 takes the NTT-domain limb itself for this digit
    const int tid = threadIdx.x;
    const int c = tid % C, g = tid / C;
    const size_t c0 = (size_t)blockIdx.x * C;
    const unsigned ncols = 1u << a.a2;
    const LimbConst m = a.lc[j];
    const WD q = (WD)m.q, q2 = (WD)m.q2;
    const u64 *src = a.digits + (size_t)blockIdx.z * a.dig_ct_stride + (size_t)i * a.dig_limb_stride;
    WD *dst = reinterpret_cast<WD *>(a.scratch) + (((size_t)blockIdx.z * L + j) * Ld + i) * a.N;
    const TW *tab = reinterpret_cast<const TW *>(a.P1) + ((size_t)j << A);
    WD v[1 << E];
#pragma unroll
    for (int k = 0; k < (1 << E); ++k) {
        u64 x = src[(size_t)tile_idx<E>(g, k, GM::lo(0)) * ncols + c0 + c];
        if (REDUCE) x = LAZY ? barrett_word_lazy(x, m) : barrett_word(x, m);
        v[k] = (WD)x;
    }
    xf_tile<XF_NEG_FWD, A, E, CP, LAZY>(v, g, c, sm, tab, q, q2);
    {  // four-step twiddle psi^(j2 (2 k1 + 1)), table in this pass's [rho][j2] layout
        const TW *TTt = reinterpret_cast<const TW *>(a.TTt) + (size_t)j * a.N;
#pragma unroll
        for (int k = 0; k < (1 << E); ++k)
            v[k] = mul_tw<LAZY>(v[k], ldg_tw(TTt + (size_t)tile_idx<E>(g, k, GM::lo(GM::NS - 1)) * ncols + c0 + c), q);
    }
    tile_put<E, CP>(sm, v, g, c, GM::lo(GM::NS - 1));  // own slots (last tile_get was at this window): no barrier needed
    __syncthreads();
    WD *d = dst + c0 * (size_t)(1 << A);
#pragma unroll 4
    for (int i = 0; i < (1 << E); ++i) {  // (C << A) / NT = 2^E words per thread (a full unroll costs 16 registers: 78 instead of 62)
        const int e = tid + i * NT;
        int cc = e >> A, r = e & ((1 << A) - 1);
        d[e] = sm[r * CP + cc];
    }
}

// (hi:lo) += x * k for x, k < 2^63: four 32x32 products (the two cross terms cannot overflow a word when
// both high halves are below 2^31) and one 128-bit carry chain.
__device__ __forceinline__ void mac128(u64 &lo, u64 &hi, u64 x, u64 k) {
#ifdef __CUDA_ARCH__
    u32 x0, x1, k0, k1, a0, a1, a2, a3, b0, b1, m0, m1, d0, d1;
    asm("mov.b64 {%0,%1}, %2;" : "=r"(x0), "=r"(x1) : "l"(x));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(k0), "=r"(k1) : "l"(k));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(a0), "=r"(a1) : "l"(lo));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(a2), "=r"(a3) : "l"(hi));
    u64 p00, mid, p11;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p00) : "r"(x0), "r"(k0));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(mid) : "r"(x0), "r"(k1));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(mid) : "r"(x1), "r"(k0));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p11) : "r"(x1), "r"(k1));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(b0), "=r"(b1) : "l"(p00));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(m0), "=r"(m1) : "l"(mid));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(d0), "=r"(d1) : "l"(p11));
    asm("add.cc.u32 %0, %0, %4;\n\taddc.cc.u32 %1, %1, %5;\n\taddc.cc.u32 %2, %2, %6;\n\taddc.u32 %3, %3, %7;"
        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3)
        : "r"(b0), "r"(b1), "r"(d0), "r"(d1));
    asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, 0;" : "+r"(a1), "+r"(a2), "+r"(a3) : "r"(m0), "r"(m1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(lo) : "r"(a0), "r"(a1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(hi) : "r"(a2), "r"(a3));
#else
    unsigned __int128 acc = ((unsigned __int128)hi << 64) | lo;
    acc += (unsigned __int128)x * k;
    lo = (u64)acc;
    hi = (u64)(acc >> 64);
#endif
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// ---- TMA (cp.async.bulk.tensor) + mbarrier: the Blackwell-native way to stage the tiles ------------
__device__ __forceinline__ void mbar_init(u64 *bar, unsigned count) {
    unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(b), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, unsigned bytes) {
    unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, unsigned parity) {
    unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_LOOP_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(b),
        "r"(parity)
        : "memory");
}
// One [2^A][C] box of a rank-3 tensor (cols, rows, slab) -> dense shared-memory tile; completion is
// signalled on `bar` by the copy engine (SASS: UTMALDG).
__device__ __forceinline__ void tma_load_tile(void *sdst, const void *tmap, int col0, int slab, u64 *bar) {
    unsigned d = (unsigned)__cvta_generic_to_shared(sdst);
    unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(d),
        "l"(tmap), "r"(col0), "r"(0), "r"(slab), "r"(b)
        : "memory");
}

// Asynchronous copy of a [2^A][C] tile of T words (row stride `ncols` words in global memory) into a
// dense [2^A][C] shared-memory tile, 16 bytes per request (LDGSTS), issued by all NT threads.
template <typename T, int A, int C, int NT>
__device__ __forceinline__ void stage_tile(T *sdst, const T *gsrc, unsigned ncols, int tid) {
    constexpr int PER = 16 / sizeof(T);  // words per 16-byte chunk
    constexpr int CH = C / PER;          // chunks per row
    for (int e = tid; e < (CH << A); e += NT) {
        int r = e / CH, part = e % CH;
        cp_async16(sdst + r * C + part * PER, gsrc + (size_t)r * ncols + part * PER);
    }
}

// Key-switch accumulators: sum_i x_i * k_i kept unreduced.
//   u64 limbs: 128-bit (lo, hi) registers;  u32 limbs: one 64-bit register (products < 2^62).
template <typename WD>
struct KsAcc;
template <>
struct KsAcc<u64> {
    u64 lo, hi;
    __device__ __forceinline__ void clear() { lo = hi = 0; }
    __device__ __forceinline__ void mac(u64 x, u64 k) { mac128(lo, hi, x, k); }
    __device__ __forceinline__ u64 reduce(const LimbConst &m) const { return reduce128(hi, lo, m); }
    __device__ __forceinline__ void set(u64 r) {
        lo = r;
        hi = 0;
    }
};
template <>
struct KsAcc<u32> {
    u64 s;
    __device__ __forceinline__ void clear() { s = 0; }
    __device__ __forceinline__ void mac(u32 x, u32 k) { s += (u64)x * k; }
    __device__ __forceinline__ u32 reduce(const LimbConst &m) const { return (u32)barrett_word(s, m); }
    __device__ __forceinline__ void set(u32 r) { s = r; }
};

// Shared memory of ks_pass2 in bytes: two stages of the digit tile (WD), key_b and key_a tiles (WD),
// exchange tile [2^A][C+1] (WD).
template <typename WD, int A, int C>
__host__ __device__ constexpr size_t ks2_exch_bytes() {
    return (((size_t)(1 << A) * (C + 1) * sizeof(WD) + 15) / 16) * 16;
}
template <typename WD, int A, int C>
__host__ __device__ constexpr size_t ks2_smem_bytes() {
    return (size_t)(1 << A) * C * (2 * sizeof(WD) + 2 * sizeof(WD)) + ks2_exch_bytes<WD, A, C>() + 64;
}

// TMA = true: the digit tile and the two key tiles are fetched by the copy engine
// (cp.async.bulk.tensor + mbarrier) from the rank-3 tensor maps; TMA = false: LDGSTS (cp.async).
struct KsMaps {
    alignas(64) unsigned char scratch[128], key_b[128], key_a[128];  // CUtensorMap images
};
// Resident CTAs per SM the compiler must leave room for: 64-bit words need 128 registers (128-bit accumulators);
// 32-bit words need about 80, so three times as many CTAs fit and hide the barriers of the digit loop.
template <typename WD, int NT, int C>
__host__ __device__ constexpr int ks2_min_ctas() {
    return sizeof(WD) == 8 ? (C <= 4 ? 4 : (C <= 8 ? 2 : 1)) : (768 / NT > 16 ? 16 : (768 / NT < 1 ? 1 : 768 / NT));
}
template <typename WD, int A, int E, int C, int LAZY, bool ADD, bool DIAG, bool TMA>
__global__ void __launch_bounds__(C *(1 << (A - E)), ks2_min_ctas<WD, C *(1 << (A - E)), C>()) ks_pass2_kernel(KsArgs a, const __grid_constant__ KsMaps maps) {
    typedef TileGeom<A, E> GM;
    typedef typename TwOf<WD>::type TW;
    constexpr int CP = C + 1;
    constexpr int NT = C * GM::G;
    constexpr int R = 1 << E;
    constexpr int TILE = (1 << A) * C;
    // exchange tile of the digit loop: bit-weighted conflict-free layout (ntt_tile.cuh); the epilogue keeps the
    // padded layout, whose transposed read is conflict-free
    constexpr int SWZ = (E == 3 && C == 4) ? 1 : 0;
    static_assert(GM::lo(GM::NS - 1) == 0, "the key tiles are addressed through the last register window");
    extern __shared__ __align__(128) unsigned char sm_raw[];
    // (the resident key holds words of the transform type: u32 on the 32-bit path, see ksk_finalize)
    WD *stKb = reinterpret_cast<WD *>(sm_raw);  // key_b tile of the current digit
    WD *stKa = stKb + TILE;                     // key_a tile
    WD *stS = reinterpret_cast<WD *>(stKa + TILE);  // 2 stages of the digit's pass-1 output
    WD *sm = stS + 2 * TILE;                        // exchange buffer
    u64 *bars = reinterpret_cast<u64 *>(sm_raw + ks2_smem_bytes<WD, A, C>() - 64);  // [0,1]: digit stages, [2]: keys
    const int L = a.L;
    // grid = (ciphertexts, tiles, target limbs): the ciphertext index runs fastest so that the CTAs
    // resident at the same time share the key tiles of one (target limb, tile) through L2.
    const int j = a.j0 + blockIdx.z;
    const int tid = threadIdx.x;
    const int c = tid % C, g = tid / C;
    const size_t c0 = (size_t)blockIdx.y * C;
    const unsigned ncols = 1u << a.a1;  // rho runs along the contiguous dimension
    const LimbConst m = a.lc[j];
    const WD q = (WD)m.q, q2 = (WD)m.q2;
    const size_t ct = blockIdx.x;
    const TW *W = reinterpret_cast<const TW *>(a.W2) + (size_t)j * a.w2_stride;
    constexpr int lo_in = GM::lo(0), lo_out = GM::lo(GM::NS - 1);
    const int Ld = a.Ld, jg = a.joff + a.jstep * j;  // jg: index of this target limb in the whole basis
    const int nd = DIAG ? Ld - 1 : Ld;  // digits that need a transform
    auto digit_of = [&](int t) { return (DIAG && t >= jg) ? t + 1 : t; };
    const WD *scr = reinterpret_cast<const WD *>(a.scratch) + (ct * L + j) * (size_t)Ld * a.N + c0;
    const WD *kbase_b = reinterpret_cast<const WD *>(a.key_b) + (size_t)j * a.N + c0;
    const WD *kbase_a = reinterpret_cast<const WD *>(a.key_a) + (size_t)j * a.N + c0;
    const size_t kstride = (size_t)L * a.N;

    KsAcc<WD> acc0[R], acc1[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        acc0[k].clear();
        acc1[k].clear();
    }
    const int slab0 = (int)((ct * L + j) * Ld);  // scratch slabs of this (ciphertext, target limb)
    if (TMA) {
        if (tid == 0) {
            mbar_init(bars + 0, 1);
            mbar_init(bars + 1, 1);
            mbar_init(bars + 2, 1);
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        __syncthreads();
        if (tid == 0 && nd > 0) {
            mbar_expect_tx(bars + 0, TILE * sizeof(WD));
            tma_load_tile(stS, maps.scratch, (int)c0, slab0 + digit_of(0), bars + 0);
        }
    } else if (nd > 0) {
        stage_tile<WD, A, C, NT>(stS, scr + (size_t)digit_of(0) * a.N, ncols, tid);
        cp_async_commit();
    }
    // Key rows are stored permuted, row (g * 2^E + k) of a limb at (k * G + g) (key_permute_kernel): the dense
    // tile the copy engine delivers is then [k][g][c] and the lanes of a wavefront (consecutive g, fixed k)
    // read consecutive words instead of rows 2^E apart.
    if (DIAG) {  // digit i == j: the NTT-domain limb itself
        const u64 *src = a.dig_ntt + (ct * L + j) * a.N + c0 + c;
        const WD *kb = kbase_b + (size_t)jg * kstride + c, *ka = kbase_a + (size_t)jg * kstride + c;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            size_t off = (size_t)tile_idx<E>(g, k, lo_out) * ncols;
            size_t koff = (size_t)(k * GM::G + g) * ncols;
            WD x = (WD)src[off];
            acc0[k].mac(x, __ldg(kb + koff));
            acc1[k].mac(x, __ldg(ka + koff));
        }
    }
    for (int t = 0; t < nd; ++t) {
        const int i = digit_of(t);
        if (TMA) {
            mbar_wait(bars + (t & 1), (t >> 1) & 1);  // scratch(t) landed
            __syncthreads();  // MAC(t-1) finished reading the key stage; stage (t+1)&1 is free
            if (tid == 0) {
                if (t + 1 < nd) {
                    mbar_expect_tx(bars + ((t + 1) & 1), TILE * sizeof(WD));
                    tma_load_tile(stS + ((t + 1) & 1) * TILE, maps.scratch, (int)c0, slab0 + digit_of(t + 1), bars + ((t + 1) & 1));
                }
                mbar_expect_tx(bars + 2, 2 * TILE * sizeof(WD));
                tma_load_tile(stKb, maps.key_b, (int)c0, i * L + j, bars + 2);
                tma_load_tile(stKa, maps.key_a, (int)c0, i * L + j, bars + 2);
            }
        } else {
            cp_async_wait_all();
            __syncthreads();  // scratch(t) landed for everyone; MAC(t-1) finished reading the key stage
            if (t + 1 < nd) stage_tile<WD, A, C, NT>(stS + ((t + 1) & 1) * TILE, scr + (size_t)digit_of(t + 1) * a.N, ncols, tid);
            stage_tile<WD, A, C, NT>(stKb, kbase_b + (size_t)i * kstride, ncols, tid);
            stage_tile<WD, A, C, NT>(stKa, kbase_a + (size_t)i * kstride, ncols, tid);
            cp_async_commit();
        }
        WD v[R];
        const WD *S = stS + (t & 1) * TILE;
#pragma unroll
        for (int k = 0; k < R; ++k) v[k] = S[tile_idx<E>(g, k, lo_in) * C + c];
        xf_tile<XF_CYC_FWD, A, E, CP, LAZY, SWZ>(v, g, c, sm, W, q, q2);
        if (TMA) {
            mbar_wait(bars + 2, t & 1);  // keys(t) landed
        } else {
            cp_async_wait_all();
            __syncthreads();  // keys(t) (and scratch(t+1)) landed for everyone
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int so = (k * GM::G + g) * C + c;
            // lazy8: one conditional subtraction ([0,4q) -> [0,2q)) is enough for the 128-bit accumulators
            WD x = (LAZY == 2) ? csub(v[k], q2) : canon2<LAZY>(v[k], q);
            acc0[k].mac(x, stKb[so]);
            acc1[k].mac(x, stKa[so]);
        }
        if ((t + 1) % a.reduce_every == 0 && t + 1 < nd) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                acc0[k].set(acc0[k].reduce(m));
                acc1[k].set(acc1[k].reduce(m));
            }
        }
    }
    // epilogue: reduce, add d0 / d1, inverse pass 2 (cyclic DIT + four-step twiddle and 1/N), transposed store
    const TW *Wi = reinterpret_cast<const TW *>(a.W2i) + (size_t)j * a.w2_stride;
    const TW *TTi = reinterpret_cast<const TW *>(a.TTi) + (size_t)j * a.N;
#pragma unroll
    for (int comp = 0; comp < 2; ++comp) {
        WD v[R];
        const u64 *add = comp ? a.add1 : a.add0;
        WD *out = reinterpret_cast<WD *>(comp ? a.out1 : a.out0) + (ct * L + j) * a.N;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            WD r = comp ? acc1[k].reduce(m) : acc0[k].reduce(m);
            if (ADD) {
                size_t off = (size_t)tile_idx<E>(g, k, lo_out) * ncols + c0 + c;
                r = csub((WD)(r + (WD)add[(ct * L + j) * a.N + off]), q);
            }
            v[k] = r;
        }
        __syncthreads();
        xf_tile<XF_CYC_INV, A, E, CP, LAZY>(v, g, c, sm, Wi, q, q2);
#pragma unroll
        for (int k = 0; k < R; ++k) {
            size_t off = (size_t)tile_idx<E>(g, k, lo_in) * ncols + c0 + c;
            v[k] = mul_tw<LAZY>(v[k], ldg_tw(TTi + off), q);
        }
        tile_put<E, CP>(sm, v, g, c, lo_in);  // own slots (the inverse transform's last tile_get was at window lo_in)
        __syncthreads();
        WD *d = out + c0 * (size_t)(1 << A);
#pragma unroll 4
        for (int i = 0; i < R; ++i) {  // (C << A) / NT = 2^E words per thread
            const int e = tid + i * NT;
            int cc = e >> A, r = e & ((1 << A) - 1);
            d[e] = sm[r * CP + cc];
        }
    }
}

// Last pass of the inverse transform (negacyclic GS over rho) fused with rescale_into
// (poly.rs:214-225): limb i < L-1 of the result is (c_i - (c_last % q_i)) * q_last^-1 mod q_i, where
// c_last is the already finished coefficient-domain last limb.  src: [cts][L][N] transposed
// inverse-pass-2 output (WD); last: [cts][N] coefficient domain (u64); dst: [cts][L-1][N] (u64).
template <typename WD, int A, int E, int C, int LAZY, int FIXLOGN = 0>
__global__ void __launch_bounds__(C *(1 << (A - E))) inv_pass1_rescale_kernel(PassArgs a, const u64 *__restrict__ last,
                                                                               const void *__restrict__ qlinv_) {
    if (FIXLOGN) {  // compile-time ring degree (see ntt_pass_kernel)
        a.ncols = 1u << (FIXLOGN > A ? FIXLOGN - A : 0);
        a.N = (size_t)1 << FIXLOGN;
    }
    typedef TileGeom<A, E> GM;
    typedef typename TwOf<WD>::type TW;
    constexpr int CP = C + 1;
    extern __shared__ __align__(16) unsigned char sm_raw[];
    WD *sm = reinterpret_cast<WD *>(sm_raw);
    const WD *src = reinterpret_cast<const WD *>(a.src);
    u64 *dst = reinterpret_cast<u64 *>(a.dst);
    const int tid = threadIdx.x;
    const int c = tid % C, g = tid / C;
    const int limb = blockIdx.y;  // < dstL
    const size_t c0 = (size_t)blockIdx.x * C;
    const size_t base_in = ((size_t)blockIdx.z * a.L + limb) * a.N;
    const size_t base_out = ((size_t)blockIdx.z * a.dstL + limb) * a.N;
    const LimbConst m = a.lc[limb];
    const WD q = (WD)m.q, q2 = (WD)m.q2;
    const TW *tab = reinterpret_cast<const TW *>(a.tab) + (size_t)limb * a.tab_stride;
    const TW qi = ldg_tw(reinterpret_cast<const TW *>(qlinv_) + limb);
    WD v[1 << E];
#pragma unroll
    for (int k = 0; k < (1 << E); ++k) v[k] = src[base_in + (size_t)tile_idx<E>(g, k, GM::lo(GM::NS - 1)) * a.ncols + c0 + c];
    xf_tile<XF_NEG_INV, A, E, CP, LAZY>(v, g, c, sm, tab, q, q2);
#pragma unroll
    for (int k = 0; k < (1 << E); ++k) {
        size_t off = (size_t)tile_idx<E>(g, k, GM::lo(0)) * a.ncols + c0 + c;
        WD ci = canon2<LAZY>(v[k], q);
        WD cl = (WD)barrett_word(last[(size_t)blockIdx.z * a.N + off], m);
        WD d = ci >= cl ? ci - cl : ci + q - cl;
        dst[base_out + off] = (u64)shoup(d, qi, q);
    }
}

// Coefficient-domain limbs that already exist (the rotated c1 of rotate_ciphertext, engine.rs:429) pushed
// into the gather buffers of every GPU: src [cs][L][N] -> slot (m_off + m_step*limb) of [slot][m_cs][N],
// 16 bytes per store.
struct PushArgs {
    const u64 *src;
    u64 *peer[8];
    int npeer, m_first;
    int m_off, m_step;
    size_t m_cs;
    int L;
    int logn;
    size_t total2;  // cs * L * N / 2
};
__global__ void lshard_push_kernel(PushArgs a) {
    const size_t half = (size_t)1 << (a.logn - 1);
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < a.total2; t += (size_t)gridDim.x * blockDim.x) {
        const size_t k2 = t & (half - 1), row = t >> (a.logn - 1);
        const size_t ct = row / a.L, limb = row % a.L;
        const ulonglong2 v = reinterpret_cast<const ulonglong2 *>(a.src)[t];
        const size_t dst = (((size_t)(a.m_off + a.m_step * (int)limb) * a.m_cs + ct) << (a.logn - 1)) + k2;
        for (int t2 = 0; t2 < a.npeer; ++t2) {
            int p = a.m_first + t2;
            if (p >= a.npeer) p -= a.npeer;
            reinterpret_cast<ulonglong2 *>(a.peer[p])[dst] = v;
        }
    }
}

// =================================================================================================
// Limb-sharded mode (optional): barrier between the GPUs that hold the limbs of one batch.
// Launched on the stream right after the kernel whose peer stores must be visible: thread t publishes
// this rank's epoch in GPU t's flag word (release, system scope) and then waits for GPU t's epoch in
// its own flag array (acquire).  A wall-clock limit turns a lost peer into an error word instead of
// a hung device.
// =================================================================================================
struct BarArgs {
    unsigned *peer_flags[8];  // flag arrays of every GPU (own included), [world] words each
    unsigned *my_flags;
    unsigned *err;       // device error word of this rank (read by the poison guard)
    unsigned *host_err;  // the same word in mapped pinned host memory: the host sees a failure without synchronising
    int rank, world;
    unsigned epoch;
    unsigned long long timeout_ns;
};
__global__ void lshard_barrier_kernel(BarArgs a) {
#ifdef __CUDA_ARCH__
    const int t = threadIdx.x;
    if (t >= a.world) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.peer_flags[t] + a.rank), "r"(a.epoch) : "memory");
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a.my_flags + t) : "memory");
        if ((int)(v - a.epoch) >= 0) break;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > a.timeout_ns) {
            atomicExch(a.err, a.epoch);
            if (a.host_err) {
                *reinterpret_cast<volatile unsigned *>(a.host_err) = a.epoch;
                __threadfence_system();
            }
            break;
        }
        __nanosleep(200);
    }
#endif
}

// Fail-closed guard of the limb-sharded entry points: launched after the last kernel of a call.  If a barrier of
// this rank gave up waiting for a peer, the kernels behind it consumed gather / last buffers that peer never filled,
// so the outputs are overwritten with all-ones words: not canonical for any modulus (every q < 2^63), i.e. rejected
// by every reducedness scan, never mistaken for ciphertext limbs.
__global__ void lshard_poison_kernel(const unsigned *__restrict__ err, u64 *__restrict__ o0, size_t w0, u64 *__restrict__ o1, size_t w1) {
    if (*err == 0) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < w0; i += (size_t)gridDim.x * blockDim.x) o0[i] = ~0ull;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < w1; i += (size_t)gridDim.x * blockDim.x) o1[i] = ~0ull;
}

// =================================================================================================
// Fused four-step transform: both passes of to_ntt_domain / to_coeff_domain (poly.rs:136-166) in one
// CTA per limb, the limb resident in shared memory between the passes, so a limb crosses HBM exactly
// once in each direction (16 N bytes per transform = the algorithmic figure of SURVEY 8d).
// N = 2^(A1+A2) <= 2^14; N/16 threads; in place.
//   forward : load x[j1][j2] (lanes over j2) -> negacyclic pass over j1 -> four-step twiddle ->
//             transposed hand-off through shared memory -> cyclic pass over j2 (lanes over rho) -> store
//   inverse : the mirror image.
// Shared memory: max(N * 17/16, N + max(n1, n2)) words, used first as per-tile exchange regions of the
// first pass, then as the padded hand-off matrix, then as exchange regions of the second pass.
// =================================================================================================
struct FusedArgs {
    u64 *data;            // [batch][L][N], in place
    const LimbConst *lc;  // [L]
    const void *P1, *W2, *TT;  // forward: P1, W2, TTt; inverse: P1i, W2i, TTi  (TwOf<WD>)
    size_t w2_stride;
    int L;
};
template <int A1, int A2>
__host__ __device__ constexpr size_t fused_smem_words() {
    return ((size_t)1 << (A1 + A2)) / 16 * 17 > ((size_t)1 << (A1 + A2)) + ((size_t)1 << A1) + ((size_t)1 << A2)
               ? ((size_t)1 << (A1 + A2)) / 16 * 17
               : ((size_t)1 << (A1 + A2)) + ((size_t)1 << A1) + ((size_t)1 << A2);
}

template <typename WD, int A1, int A2, int LAZY, bool INV>
__global__ void __launch_bounds__((1 << (A1 + A2)) / 16) ntt_fused_kernel(FusedArgs a) {
    constexpr int E = 4, C = 16, CP = 17;
    typedef typename TwOf<WD>::type TW;
    typedef TileGeom<A1, E> G1;
    typedef TileGeom<A2, E> G2;
    constexpr int N1 = 1 << A1, N2 = 1 << A2;
    constexpr size_t N = (size_t)N1 * N2;
    extern __shared__ __align__(16) unsigned char sm_raw[];
    WD *sm = reinterpret_cast<WD *>(sm_raw);
    const int tid = threadIdx.x;
    const int limb = blockIdx.x % a.L;
    u64 *d = a.data + (size_t)blockIdx.x * N;
    const LimbConst m = a.lc[limb];
    const WD q = (WD)m.q, q2 = (WD)m.q2;
    const TW *P1 = reinterpret_cast<const TW *>(a.P1) + (size_t)limb * N1;
    const TW *W2 = reinterpret_cast<const TW *>(a.W2) + (size_t)limb * a.w2_stride;
    const TW *TT = reinterpret_cast<const TW *>(a.TT) + (size_t)limb * N;
    // pass over j1 / rho (negacyclic, length n1): tiles over j2
    const int c1 = tid % C, g1 = (tid / C) % G1::G, t1 = tid / (C * G1::G);
    // pass over j2 / gamma (cyclic, length n2): tiles over rho
    const int c2 = tid % C, g2 = (tid / C) % G2::G, t2 = tid / (C * G2::G);
    WD v[1 << E];
    if (!INV) {
        constexpr int lo_in1 = G1::lo(0), lo_out1 = G1::lo(G1::NS - 1), lo_in2 = G2::lo(0), lo_out2 = G2::lo(G2::NS - 1);
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = (WD)d[(size_t)tile_idx<E>(g1, k, lo_in1) * N2 + t1 * C + c1];
        xf_tile<XF_NEG_FWD, A1, E, CP, LAZY>(v, g1, c1, sm + (size_t)t1 * N1 * CP, P1, q, q2);
#pragma unroll
        for (int k = 0; k < 16; ++k)  // four-step twiddle, table in [rho][j2] layout
            v[k] = mul_tw<LAZY>(v[k], ldg_tw(TT + (size_t)tile_idx<E>(g1, k, lo_out1) * N2 + t1 * C + c1), q);
        __syncthreads();  // exchange regions are dead
#pragma unroll
        for (int k = 0; k < 16; ++k) sm[(size_t)(t1 * C + c1) * (N1 + 1) + tile_idx<E>(g1, k, lo_out1)] = v[k];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = sm[(size_t)tile_idx<E>(g2, k, lo_in2) * (N1 + 1) + t2 * C + c2];
        __syncthreads();  // hand-off consumed before the second pass reuses the memory
        xf_tile<XF_CYC_FWD, A2, E, CP, LAZY>(v, g2, c2, sm + (size_t)t2 * N2 * CP, W2, q, q2);
#pragma unroll
        for (int k = 0; k < 16; ++k) d[(size_t)tile_idx<E>(g2, k, lo_out2) * N1 + t2 * C + c2] = (u64)canon2<LAZY>(v[k], q);
    } else {
        constexpr int lo_in2 = G2::lo(G2::NS - 1), lo_out2 = G2::lo(0), lo_in1 = G1::lo(G1::NS - 1), lo_out1 = G1::lo(0);
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = (WD)d[(size_t)tile_idx<E>(g2, k, lo_in2) * N1 + t2 * C + c2];
        xf_tile<XF_CYC_INV, A2, E, CP, LAZY>(v, g2, c2, sm + (size_t)t2 * N2 * CP, W2, q, q2);
#pragma unroll
        for (int k = 0; k < 16; ++k)  // four-step twiddle and 1/N, table in [j2][rho] layout
            v[k] = mul_tw<LAZY>(v[k], ldg_tw(TT + (size_t)tile_idx<E>(g2, k, lo_out2) * N1 + t2 * C + c2), q);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) sm[(size_t)(t2 * C + c2) * (N2 + 1) + tile_idx<E>(g2, k, lo_out2)] = v[k];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = sm[(size_t)tile_idx<E>(g1, k, lo_in1) * (N2 + 1) + t1 * C + c1];
        __syncthreads();
        xf_tile<XF_NEG_INV, A1, E, CP, LAZY>(v, g1, c1, sm + (size_t)t1 * N1 * CP, P1, q, q2);
#pragma unroll
        for (int k = 0; k < 16; ++k) d[(size_t)tile_idx<E>(g1, k, lo_out1) * N2 + t1 * C + c1] = (u64)canon2<LAZY>(v[k], q);
    }
}
