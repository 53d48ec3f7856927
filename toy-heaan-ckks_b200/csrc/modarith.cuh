// modarith.cuh -- 64-bit modular arithmetic for RNS limbs on the sm_100a integer (IMAD) pipe.
//
// Replaces the reference's scalar helpers add_mod / sub_mod / mul_mod (poly.rs:629-653, where
// mul_mod is a u128 product followed by a software `%`).  All results that leave a kernel are the
// canonical representatives in [0, q) the reference produces; inside a kernel values may be lazy:
//   LAZY = 1 (q < 2^62): Cooley-Tukey butterflies keep values in [0, 4q), Gentleman-Sande in [0, 2q)
//                        (Harvey's bounds), constants are multiplied with Shoup's precomputed quotient;
//   LAZY = 2 (q < 2^61, 64-bit words): the Shoup quotient is approximated (three partial products instead
//                        of four, see shoup_lazy8), products land in [0, 4q), Cooley-Tukey values live
//                        in [0, 8q), Gentleman-Sande in [0, 4q): 16 % fewer issue slots per butterfly
//                        on B200 (tools/imad_bench2.cu: 36.8 vs 44.0 cycles per warp);
//   LAZY = 0 (q < 2^63): every step is reduced to [0, q)  (the reference admits 63-bit primes,
//                        src/math/utils.rs:48, for which 4q does not fit a word).
#pragma once
#include <cstdint>

typedef unsigned long long u64;
typedef long long i64;

#ifndef __CUDACC__
// Host build (tests/emul): the same arithmetic and index code compiled by g++, so the CPU test
// suite can check the transforms and tables without a GPU.  Never part of libckks_b200.so.
#define __device__
#define __host__
#define __forceinline__ inline
static inline u64 __umul64hi(u64 a, u64 b) { return (u64)(((unsigned __int128)a * b) >> 64); }
struct ulonglong2 {
    u64 x, y;
};
static inline ulonglong2 __ldg(const ulonglong2 *p) { return *p; }
struct uint2 {
    unsigned x, y;
};
static inline uint2 __ldg(const uint2 *p) { return *p; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline void __syncthreads() {}
static inline void __syncwarp() {}
#endif

typedef unsigned int u32;

// A constant w in [0, q) with its Shoup companion ws = floor(w * 2^64 / q).
struct alignas(16) tw_t {
    u64 w, ws;
};
// 32-bit limbs (all q < 2^31): w and ws = floor(w * 2^32 / q).  Words stay u64 in HBM (reference layout);
// only the arithmetic, the twiddle tables and the internal scratch are 32-bit.
struct alignas(8) tw32_t {
    u32 w, ws;
};
template <typename W>
struct TwOf;
template <>
struct TwOf<u64> {
    typedef tw_t type;
};
template <>
struct TwOf<u32> {
    typedef tw32_t type;
};

// Per-limb constants (one entry per RNS prime, resident in HBM / L2, read through the RO path).
struct LimbConst {
    u64 q;     // modulus
    u64 q2;    // 2q (wraps for q >= 2^63: never used in !LAZY mode)
    u64 bar;   // floor(2^64 / q): Barrett quotient for single-word reduction
    u64 c64;   // 2^64 mod q
    u64 c64s;  // Shoup companion of c64
    u64 pad[3];
};

// (a compare + predicated-subtract form in PTX was measured: marginally faster in the plain passes,
// 7 % slower in ks_pass2, so the compiler's compare/select form stays)
__device__ __forceinline__ u64 csub(u64 x, u64 q) { return x >= q ? x - q : x; }
__device__ __forceinline__ u32 csub(u32 x, u32 q) {
    u32 y = x - q;  // wraps above x when x < q
    return y < x ? y : x;
}

// x * w mod q for any word x: result in [0, 2q)  (q < 2^63).
__device__ __forceinline__ u64 shoup_lazy(u64 x, tw_t t, u64 q) {
    u64 h = __umul64hi(x, t.ws);
    return x * t.w - h * q;
}
__device__ __forceinline__ u64 shoup(u64 x, tw_t t, u64 q) { return csub(shoup_lazy(x, t, q), q); }
// 32-bit: x * w mod q for any 32-bit x: result in [0, 2q)  (q < 2^31).
__device__ __forceinline__ u32 shoup_lazy(u32 x, tw32_t t, u32 q) {
    u32 h = __umulhi(x, t.ws);
    return x * t.w - h * q;
}
__device__ __forceinline__ u32 shoup(u32 x, tw32_t t, u32 q) { return csub(shoup_lazy(x, t, q), q); }

// x * w mod q for any word x with an approximate quotient: result in [0, 4q)  (q < 2^62; callers keep
// q < 2^61 so that 8q fits a word).  h' = x1*s1 + hi32(x1*s0) + hi32(x0*s1) is at most 2 below the exact
// floor(x*ws / 2^64); r = x*w + h'*(-q) mod 2^64 is evaluated as one multiply-add chain:
// 5 IMAD.WIDE + 4 IMAD and two 32-bit adds, against 6 + 4 and about ten adds for the exact form.
__device__ __forceinline__ u64 shoup_lazy8(u64 x, tw_t t, u64 nq) {
#ifdef __CUDA_ARCH__
    u32 x0, x1, w0, w1, s0, s1, n0, n1;
    asm("mov.b64 {%0,%1}, %2;" : "=r"(x0), "=r"(x1) : "l"(x));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(w0), "=r"(w1) : "l"(t.w));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(s0), "=r"(s1) : "l"(t.ws));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(n0), "=r"(n1) : "l"(nq));
    u64 a, b, h, r, add;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(a) : "r"(x1), "r"(s0));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(b) : "r"(x0), "r"(s1));
    u32 al, ah, bl, bh, mid, cy;
    asm("mov.b64 {%0,%1}, %2;" : "=r"(al), "=r"(ah) : "l"(a));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(bl), "=r"(bh) : "l"(b));
    asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=r"(mid), "=r"(cy) : "r"(ah), "r"(bh));
    asm("mov.b64 %0, {%1,%2};" : "=l"(add) : "r"(mid), "r"(cy));
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(h) : "r"(x1), "r"(s1), "l"(add));
    u32 h0, h1, r0, r1;
    asm("mov.b64 {%0,%1}, %2;" : "=r"(h0), "=r"(h1) : "l"(h));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(x0), "r"(w0));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r) : "r"(h0), "r"(n0));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(r0), "=r"(r1) : "l"(r));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r1) : "r"(x1), "r"(w0));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r1) : "r"(x0), "r"(w1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r1) : "r"(h1), "r"(n0));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r1) : "r"(h0), "r"(n1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(r0), "r"(r1));
    (void)al;
    (void)bl;
    return r;
#else
    const u64 x0 = x & 0xffffffffull, x1 = x >> 32, s0 = t.ws & 0xffffffffull, s1 = t.ws >> 32;
    const u64 h = x1 * s1 + ((x1 * s0) >> 32) + ((x0 * s1) >> 32);
    return x * t.w + h * nq;
#endif
}
// 32-bit words have no lazy8 mode; the overload only keeps the generic butterflies compilable.
__device__ __forceinline__ u32 shoup_lazy8(u32 x, tw32_t t, u32 nq) {
    (void)nq;
    return x * t.w - __umulhi(x, t.ws) * (0u - nq);
}

// x mod q for any word x: result in [0, 2q).
__device__ __forceinline__ u64 barrett_word_lazy(u64 x, const LimbConst &m) {
    u64 h = __umul64hi(x, m.bar);
    return x - h * m.q;
}
__device__ __forceinline__ u64 barrett_word(u64 x, const LimbConst &m) {
    return csub(barrett_word_lazy(x, m), m.q);
}

// (hi * 2^64 + lo) mod q for any two words: result in [0, q).
__device__ __forceinline__ u64 reduce128(u64 hi, u64 lo, const LimbConst &m) {
    tw_t c;
    c.w = m.c64;
    c.ws = m.c64s;
    u64 r1 = shoup(hi, c, m.q);
    u64 r2 = barrett_word(lo, m);
    return csub(r1 + r2, m.q);
}

// a * b mod q, a and b any words with a * b representable (always): result in [0, q).
__device__ __forceinline__ u64 mulmod(u64 a, u64 b, const LimbConst &m) {
    return reduce128(__umul64hi(a, b), a * b, m);
}
// (a * b + c) mod q with a, b < 2^63 and any word c.
__device__ __forceinline__ u64 mulmod_add(u64 a, u64 b, u64 c, const LimbConst &m) {
    u64 lo = a * b;
    u64 hi = __umul64hi(a, b);
    u64 s = lo + c;
    hi += (s < lo) ? 1ull : 0ull;
    return reduce128(hi, s, m);
}
__device__ __forceinline__ u64 addmod(u64 a, u64 b, u64 q) { return csub(a + b, q); }
__device__ __forceinline__ u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
__device__ __forceinline__ u64 negmod(u64 a, u64 q) { return a ? q - a : 0; }

// ---- butterflies ---------------------------------------------------------------------------------
// Cooley-Tukey (decimation in time): (x, y) -> (x + w*y, x - w*y).
//   LAZY: in/out [0, 4q).  !LAZY: in/out [0, q).
//   LAZY == 2 with TOPBIT (the merged negacyclic forward transform, whose butterflies all multiply): the range is
//   managed by bit 63 alone.  x may be ANY word: if x >= 2^63 (>= 4q because q < 2^61) subtract 4q, which leaves
//   xx < max(2^63, 2^64 - 4q); then xx + v and xx - v + 4q stay below 2^64 for every v in [0, 4q) -- no overflow, no
//   64-bit compare, one predicated subtraction (ISETP + 2 predicated adds instead of 2 ISETP + 2 SEL + 2 adds).
//   Values are no longer below 8q, only below 2^64: every consumer of such a transform multiplies first
//   (shoup_lazy8 takes any word), so nothing depends on the tighter bound.
__device__ __forceinline__ u64 csub_topbit(u64 x, u64 q4) {
#ifdef __CUDA_ARCH__
    asm("{\n\t.reg .pred p;\n\t.reg .u32 lo, hi, ql, qh;\n\tmov.b64 {lo, hi}, %0;\n\tmov.b64 {ql, qh}, %1;\n\t"
        "setp.lt.s32 p, hi, 0;\n\t@p sub.cc.u32 lo, lo, ql;\n\t@p subc.u32 hi, hi, qh;\n\tmov.b64 %0, {lo, hi};\n\t}"
        : "+l"(x)
        : "l"(q4));
    return x;
#else
    return (x >> 63) ? x - q4 : x;
#endif
}
__device__ __forceinline__ u32 csub_topbit(u32 x, u32 q4) { return csub(x, q4); }  // (32-bit words have no lazy8 mode)
template <int LAZY, bool TOPBIT = false, typename W, typename TW>
__device__ __forceinline__ void ct_bfly(W &x, W &y, TW t, W q, W q2) {
    if (LAZY == 2) {  // in/out [0, 8q)  (TOPBIT: any word in, [0, 2^64) out)
        const W q4 = q2 + q2;
        W xx = TOPBIT ? csub_topbit(x, q4) : csub(x, q4);
        W v = shoup_lazy8(y, t, (W)(0 - q));
        x = xx + v;
        y = xx - v + q4;
    } else if (LAZY) {
        W xx = csub(x, q2);
        W v = shoup_lazy(y, t, q);
        x = xx + v;
        y = xx - v + q2;
    } else {
        W v = shoup(y, t, q);
        W xx = x;
        x = csub(xx + v, q);
        y = xx >= v ? xx - v : xx + q - v;
    }
}
// Gentleman-Sande (decimation in frequency): (x, y) -> (x + y, (x - y) * w).
//   LAZY: in/out [0, 2q).  !LAZY: in/out [0, q).
template <int LAZY, typename W, typename TW>
__device__ __forceinline__ void gs_bfly(W &x, W &y, TW t, W q, W q2) {
    if (LAZY == 2) {  // in/out [0, 4q)
        const W q4 = q2 + q2;
        W s = csub(x + y, q4);
        W d = x - y + q4;
        y = shoup_lazy8(d, t, (W)(0 - q));
        x = s;
    } else if (LAZY) {
        W s = csub(x + y, q2);
        W d = x - y + q2;
        y = shoup_lazy(d, t, q);
        x = s;
    } else {
        W s = csub(x + y, q);
        W d = x >= y ? x - y : x + q - y;
        y = shoup(d, t, q);
        x = s;
    }
}
// Gentleman-Sande with unit twiddle.
template <int LAZY, typename W>
__device__ __forceinline__ void gs_bfly_one(W &x, W &y, W q, W q2) {
    if (LAZY == 2) {
        const W q4 = q2 + q2;
        W s = csub(x + y, q4);
        W d = csub(x - y + q4, q4);
        x = s;
        y = d;
    } else if (LAZY) {
        W s = csub(x + y, q2);
        W d = csub(x - y + q2, q2);
        x = s;
        y = d;
    } else {
        W s = csub(x + y, q);
        W d = x >= y ? x - y : x + q - y;
        x = s;
        y = d;
    }
}
// Cooley-Tukey with unit twiddle.  LAZY: in [0, 4q) -> out [0, 4q).
template <int LAZY, typename W>
__device__ __forceinline__ void ct_bfly_one(W &x, W &y, W q, W q2) {
    if (LAZY == 2) {
        const W q4 = q2 + q2;
        W xx = csub(x, q4);
        W v = csub(y, q4);
        x = xx + v;
        y = xx - v + q4;
    } else if (LAZY) {
        W xx = csub(x, q2);
        W v = csub(y, q2);
        x = xx + v;
        y = xx - v + q2;
    } else {
        W xx = x, v = y;
        x = csub(xx + v, q);
        y = xx >= v ? xx - v : xx + q - v;
    }
}

// Canonicalise a lazy value.
template <int LAZY, typename W>
__device__ __forceinline__ W canon4(W x, W q, W q2) {  // from the CT range
    if (LAZY == 2) return csub(csub(csub(x, (W)(q2 + q2)), q2), q);
    if (LAZY) return csub(csub(x, q2), q);
    return x;
}
template <int LAZY, typename W>
__device__ __forceinline__ W canon2(W x, W q) {  // from the GS range
    if (LAZY == 2) return csub(csub(x, (W)(q + q)), q);
    if (LAZY) return csub(x, q);
    return x;
}
// x * t into the GS range ([0,4q) lazy8 / [0,2q) lazy / [0,q) strict) from any word.
template <int LAZY, typename W, typename TW>
__device__ __forceinline__ W mul_tw(W x, TW t, W q) {
    if (LAZY == 2) return shoup_lazy8(x, t, (W)(0 - q));
    return LAZY ? shoup_lazy(x, t, q) : shoup(x, t, q);
}

__device__ __forceinline__ tw_t ldg_tw(const tw_t *p) {
    ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p));
    tw_t t;
    t.w = v.x;
    t.ws = v.y;
    return t;
}
__device__ __forceinline__ tw32_t ldg_tw(const tw32_t *p) {
    uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
    tw32_t t;
    t.w = v.x;
    t.ws = v.y;
    return t;
}
