// ntt_tile.cuh -- register-tiled 2^A-point transforms over a tile of C independent columns.
//
// Building block of the four-step negacyclic NTT that replaces the reference's
// to_ntt_domain / to_coeff_domain (poly.rs:136-166) and forward_ntt / inverse_ntt /
// cooley_tukey_ntt (poly.rs:574-615).
//
// A CTA owns a tile of 2^A "rows" (the transform dimension, strided in memory) by C columns
// (contiguous in memory, one column per lane, so that global accesses coalesce and every twiddle
// is warp-uniform).  Thread (g, c), g < 2^(A-E), holds 2^E elements of column c in registers and
// performs up to E radix-2 stages on them without touching memory; between such steps the tile is
// exchanged through shared memory.  Step t's register window covers index bits [lo+E-1 .. lo]:
//     idx(g, k) = (g >> lo) << (lo+E) | k << lo | (g & (2^lo - 1)).
//
// Four transforms (n = 2^A):
//   neg_fwd : merged negacyclic Cooley-Tukey, natural in -> bit-reversed out, table P[m+i] = psi^brv(m+i)
//   cyc_fwd : cyclic decimation-in-frequency,  natural in -> bit-reversed out, table W[e] = omega^e
//   cyc_inv : cyclic decimation-in-time,       bit-reversed in -> natural out, table W[e] = omega^-e
//   neg_inv : merged negacyclic Gentleman-Sande, bit-reversed in -> natural out, P[m+i] = psi^-brv(m+i)
// (no 1/n scaling here; the caller folds it into its per-element table).
#pragma once
#include "modarith.cuh"

// 1 (default): the butterflies of the merged negacyclic forward transform manage their range by bit 63 alone
// (modarith.cuh, ct_bfly TOPBIT); 0: one exact conditional subtraction per butterfly ([0, 8q) invariant).
#ifndef CKKS_NEG_FWD_TOPBIT
#define CKKS_NEG_FWD_TOPBIT 1
#endif

template <int A, int E>
struct TileGeom {
    static constexpr int NS = (A + E - 1) / E;  // number of register steps
    static constexpr int G = 1 << (A > E ? A - E : 0);  // thread groups per column
    static constexpr int R = 1 << E;                     // registers (elements) per thread
    // forward (top-down) step t: window low bit
    __host__ __device__ static constexpr int lo(int t) { return (A - (t + 1) * E) < 0 ? 0 : (A - (t + 1) * E); }
    // highest unprocessed bit entering forward step t
    __host__ __device__ static constexpr int top(int t) { return A - t * E - 1; }
};

template <int E>
__device__ __forceinline__ int tile_idx(int g, int k, int lo) {
    return ((g >> lo) << (lo + E)) | (k << lo) | (g & ((1 << lo) - 1));
}

// ---- one register step of each transform (T = step number in forward order) ---------------------
template <int A, int E, int T, int LAZY, typename WD, typename TW>
__device__ __forceinline__ void neg_fwd_step(WD (&v)[1 << E], int g, const TW *__restrict__ P, WD q, WD q2) {
    constexpr int lo = TileGeom<A, E>::lo(T);
    constexpr int top = TileGeom<A, E>::top(T);
#pragma unroll
    for (int b = top; b >= lo; --b) {
        const int rb = b - lo;
        const int s = A - 1 - b;
        const int base = (1 << s) + ((g >> lo) << (lo + E - b - 1));
#pragma unroll
        for (int k = 0; k < (1 << E); ++k) {
            if (k & (1 << rb)) continue;
            TW tw = ldg_tw(P + base + (k >> (rb + 1)));
            ct_bfly<LAZY, CKKS_NEG_FWD_TOPBIT != 0>(v[k], v[k | (1 << rb)], tw, q, q2);
        }
    }
}

template <int A, int E, int T, int LAZY, typename WD, typename TW>
__device__ __forceinline__ void neg_inv_step(WD (&v)[1 << E], int g, const TW *__restrict__ P, WD q, WD q2) {
    constexpr int lo = TileGeom<A, E>::lo(T);
    constexpr int top = TileGeom<A, E>::top(T);
#pragma unroll
    for (int b = lo; b <= top; ++b) {
        const int rb = b - lo;
        const int s = A - 1 - b;
        const int base = (1 << s) + ((g >> lo) << (lo + E - b - 1));
#pragma unroll
        for (int k = 0; k < (1 << E); ++k) {
            if (k & (1 << rb)) continue;
            TW tw = ldg_tw(P + base + (k >> (rb + 1)));
            gs_bfly<LAZY>(v[k], v[k | (1 << rb)], tw, q, q2);
        }
    }
}

template <int A, int E, int T, int LAZY, typename WD, typename TW>
__device__ __forceinline__ void cyc_fwd_step(WD (&v)[1 << E], int g, const TW *__restrict__ W, WD q, WD q2) {
    constexpr int lo = TileGeom<A, E>::lo(T);
    constexpr int top = TileGeom<A, E>::top(T);
    const int glo = g & ((1 << lo) - 1);
#pragma unroll
    for (int b = top; b >= lo; --b) {
        const int rb = b - lo;
        const int sh = A - 1 - b;
#pragma unroll
        for (int k = 0; k < (1 << E); ++k) {
            if (k & (1 << rb)) continue;
            const int klow = k & ((1 << rb) - 1);
            if (lo == 0 && klow == 0) {
                gs_bfly_one<LAZY>(v[k], v[k | (1 << rb)], q, q2);
            } else {
                const int e = (((klow << lo) | glo)) << sh;
                TW tw = ldg_tw(W + e);
                gs_bfly<LAZY>(v[k], v[k | (1 << rb)], tw, q, q2);
            }
        }
    }
}

template <int A, int E, int T, int LAZY, typename WD, typename TW>
__device__ __forceinline__ void cyc_inv_step(WD (&v)[1 << E], int g, const TW *__restrict__ W, WD q, WD q2) {
    constexpr int lo = TileGeom<A, E>::lo(T);
    constexpr int top = TileGeom<A, E>::top(T);
    const int glo = g & ((1 << lo) - 1);
#pragma unroll
    for (int b = lo; b <= top; ++b) {
        const int rb = b - lo;
        const int sh = A - 1 - b;
#pragma unroll
        for (int k = 0; k < (1 << E); ++k) {
            if (k & (1 << rb)) continue;
            const int klow = k & ((1 << rb) - 1);
            if (lo == 0 && klow == 0) {
                ct_bfly_one<LAZY>(v[k], v[k | (1 << rb)], q, q2);
            } else {
                const int e = (((klow << lo) | glo)) << sh;
                TW tw = ldg_tw(W + e);
                ct_bfly<LAZY>(v[k], v[k | (1 << rb)], tw, q, q2);
            }
        }
    }
}

// ---- shared-memory exchange between two register windows -----------------------------------------
// Two tile layouts in shared memory:
//   SWZ == 0: [idx][CP] words, CP = C + 1: the pad keeps the row-lane accesses of the transposing store
//             conflict-free (kernels that end with one); the register windows pay up to 2x wavefronts.
//   SWZ != 0: bit-weighted rows for ks_pass2's tile shape (E = 3, C = 4): addr(r, c) = sum_b w_b * bit_b(r) + c
//             with w = {4, 8, 16, 36, 72, 144, 288, 576}.  A wavefront covers 4 (u64) or 8 (u32) thread groups
//             of one register window, whose rows differ in bits {0,1,2}, {3,4,5} or {0,4,5}/{0,1,5} of r; the
//             weights of each such set are {4, 8, 16} modulo the 32 banks, so every tile_put / tile_get is
//             conflict-free for both word sizes, and the tile still fits the [2^A][C+1] allocation.  Because the
//             address is a SUM over the bits of r, the thread part (g) is computed once per window and the
//             register part (k, a compile-time constant after unrolling) folds into the LDS/STS immediate:
//             no extra instructions.  (An XOR swizzle is conflict-free too but is not affine in k; applied to
//             every kernel, with a column rotation for the transposed read, it was measured SLOWER on the
//             64-bit path -- ks_pass1 +21 % -- because these kernels are issue-bound, not shared-memory-bound.)
//             ncu before: 2.4x (u64) / 3.3x (u32) excess shared-memory wavefronts in ks_pass2.
__host__ __device__ constexpr int tile_waddr(int r) {
    return 4 * (r & 1) + 8 * ((r >> 1) & 1) + 16 * ((r >> 2) & 1) + 36 * ((r >> 3) & 1) + 72 * ((r >> 4) & 1) + 144 * ((r >> 5) & 1) +
           288 * ((r >> 6) & 1) + 576 * ((r >> 7) & 1);
}
template <int E, int CP, int SWZ>
__host__ __device__ __forceinline__ int tile_addr(int r, int c) {
    static_assert(SWZ == 0 || (E == 3 && CP == 5), "the bit-weighted layout is built for E = 3, C = 4");
    if (SWZ) return tile_waddr(r) + c;
    return r * CP + c;
}
template <int E, int CP, int SWZ = 0, typename WD>
__device__ __forceinline__ void tile_put(WD *sm, const WD (&v)[1 << E], int g, int c, int lo) {
    if (SWZ) {
        const int base = tile_waddr(tile_idx<E>(g, 0, lo)) + c;  // thread part; the k part below is a constant
#pragma unroll
        for (int k = 0; k < (1 << E); ++k) sm[base + tile_waddr(k << lo)] = v[k];
        return;
    }
#pragma unroll
    for (int k = 0; k < (1 << E); ++k) sm[tile_idx<E>(g, k, lo) * CP + c] = v[k];
}
template <int E, int CP, int SWZ = 0, typename WD>
__device__ __forceinline__ void tile_get(const WD *sm, WD (&v)[1 << E], int g, int c, int lo) {
    if (SWZ) {
        const int base = tile_waddr(tile_idx<E>(g, 0, lo)) + c;
#pragma unroll
        for (int k = 0; k < (1 << E); ++k) v[k] = sm[base + tile_waddr(k << lo)];
        return;
    }
#pragma unroll
    for (int k = 0; k < (1 << E); ++k) v[k] = sm[tile_idx<E>(g, k, lo) * CP + c];
}

// Row permutation of the resident gadget keys on the four-step path: within every limb ([2^a2 rows][2^a1]
// words, internal NTT order) row (g * 2^pe + k) moves to row (k * 2^(a2-pe) + g), pe = ks_pass2's register
// window.  perm_row() maps a logical row to where it is stored; pe < 0: not permuted.
__host__ __device__ __forceinline__ size_t perm_row(size_t row, int a2, int pe) {
    if (pe < 0 || a2 <= pe) return row;
    return ((row & (((size_t)1 << pe) - 1)) << (a2 - pe)) | (row >> pe);
}

enum { XF_NEG_FWD = 0, XF_CYC_FWD = 1, XF_CYC_INV = 2, XF_NEG_INV = 3 };

template <int KIND, int A, int E, int T, int LAZY, typename WD, typename TW>
__device__ __forceinline__ void xf_step(WD (&v)[1 << E], int g, const TW *__restrict__ tab, WD q, WD q2) {
    if (KIND == XF_NEG_FWD) neg_fwd_step<A, E, T, LAZY>(v, g, tab, q, q2);
    if (KIND == XF_CYC_FWD) cyc_fwd_step<A, E, T, LAZY>(v, g, tab, q, q2);
    if (KIND == XF_CYC_INV) cyc_inv_step<A, E, T, LAZY>(v, g, tab, q, q2);
    if (KIND == XF_NEG_INV) neg_inv_step<A, E, T, LAZY>(v, g, tab, q, q2);
}

// Full transform of the tile.  On entry the thread holds the window of the FIRST step (forward
// kinds: step 0; inverse kinds: step NS-1); on exit it holds the window of the LAST step
// (forward: NS-1; inverse: 0).  `sm` is the [2^A][CP] exchange buffer (unused if NS == 1).
// Barriers of a three-step transform (A = 8 with E = 3: windows over index bits [7:5], [4:2], [2:0]).  Threads are
// numbered tid = g * C + c (all callers).  The exchange between the two LOWER windows only moves words inside groups of
// 2^lo(1) thread groups that share the index bits above lo(1) + E - 1: with lo(1) = 2 these are the thread groups
// 4h .. 4h+3, i.e. 4 * C consecutive threads -- inside one warp when 4 * C <= 32.  That exchange therefore needs
// __syncwarp(), not a CTA barrier.  And a tile_put that follows a tile_get of the SAME window overwrites exactly the
// slots the thread itself has just read, so no barrier is needed between them either.  A three-step transform is left
// with ONE CTA barrier (the exchange that involves the top window) instead of three.
template <int KIND, int A, int E, int CP, int LAZY, int SWZ = 0, typename WD, typename TW>
__device__ __forceinline__ void xf_tile(WD (&v)[1 << E], int g, int c, WD *sm, const TW *__restrict__ tab, WD q, WD q2) {
    typedef TileGeom<A, E> GM;
    constexpr bool FWD = (KIND == XF_NEG_FWD || KIND == XF_CYC_FWD);
    static_assert(GM::NS >= 1 && GM::NS <= 3, "1..3 register steps supported");
    constexpr bool WARP_LOCAL = GM::NS == 3 && (((CP - 1) << GM::lo(1)) <= 32) && (32 % ((CP - 1) << GM::lo(1)) == 0);
    if (FWD) {
        xf_step<KIND, A, E, 0, LAZY>(v, g, tab, q, q2);
        if (GM::NS >= 2) {
            tile_put<E, CP, SWZ>(sm, v, g, c, GM::lo(0));
            __syncthreads();
            tile_get<E, CP, SWZ>(sm, v, g, c, GM::lo(1));
            xf_step<KIND, A, E, (GM::NS >= 2 ? 1 : 0), LAZY>(v, g, tab, q, q2);
        }
        if (GM::NS >= 3) {
            tile_put<E, CP, SWZ>(sm, v, g, c, GM::lo(1));  // own slots: the ones tile_get(lo(1)) read
            if (WARP_LOCAL) __syncwarp();
            else __syncthreads();
            tile_get<E, CP, SWZ>(sm, v, g, c, GM::lo(2));
            xf_step<KIND, A, E, (GM::NS >= 3 ? 2 : 0), LAZY>(v, g, tab, q, q2);
        }
    } else {
        xf_step<KIND, A, E, GM::NS - 1, LAZY>(v, g, tab, q, q2);
        if (GM::NS == 2) {
            tile_put<E, CP, SWZ>(sm, v, g, c, GM::lo(1));
            __syncthreads();
            tile_get<E, CP, SWZ>(sm, v, g, c, GM::lo(0));
            xf_step<KIND, A, E, 0, LAZY>(v, g, tab, q, q2);
        }
        if (GM::NS >= 3) {
            tile_put<E, CP, SWZ>(sm, v, g, c, GM::lo(2));
            if (WARP_LOCAL) __syncwarp();
            else __syncthreads();
            tile_get<E, CP, SWZ>(sm, v, g, c, GM::lo(1));
            xf_step<KIND, A, E, (GM::NS >= 3 ? 1 : 0), LAZY>(v, g, tab, q, q2);
            tile_put<E, CP, SWZ>(sm, v, g, c, GM::lo(1));  // own slots again
            __syncthreads();
            tile_get<E, CP, SWZ>(sm, v, g, c, GM::lo(0));
            xf_step<KIND, A, E, 0, LAZY>(v, g, tab, q, q2);
        }
    }
}
