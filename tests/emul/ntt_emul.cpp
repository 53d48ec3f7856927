// ntt_emul.cpp -- TEST INFRASTRUCTURE: runs the device transform code (csrc/ntt_tile.cuh,
// csrc/modarith.cuh) and the table builder (csrc/tables_host.hpp) on the CPU, thread by thread,
// following ntt_pass_kernel's sequence, so the not-gpu test suite can check index math, twiddle
// tables and lazy ranges against the oracle.  It is not a product path and is never loaded by the
// package.
#include <array>
#include <cstdint>
#include <vector>

#include "../../toy-heaan-ckks_b200/csrc/ntt_tile.cuh"
#include "../../toy-heaan-ckks_b200/csrc/tables_host.hpp"

namespace {
constexpr int E = 4, C = 16, CP = C + 1;

template <int KIND, int A, int T, int LAZY, typename WD, typename TW>
void step_all(std::vector<std::array<WD, 16>> &regs, const TW *tab, WD q, WD q2) {
    const int G = TileGeom<A, E>::G;
    for (int g = 0; g < G; ++g)
        for (int c = 0; c < C; ++c) {
            WD(&v)[16] = *reinterpret_cast<WD(*)[16]>(regs[g * C + c].data());
            xf_step<KIND, A, E, T, LAZY>(v, g, tab, q, q2);
        }
}
template <int A, typename WD>
void exchange(std::vector<std::array<WD, 16>> &regs, std::vector<WD> &sm, int lo_from, int lo_to) {
    const int G = TileGeom<A, E>::G;
    for (int g = 0; g < G; ++g)
        for (int c = 0; c < C; ++c) {
            WD(&v)[16] = *reinterpret_cast<WD(*)[16]>(regs[g * C + c].data());
            tile_put<E, CP>(sm.data(), v, g, c, lo_from);
        }
    for (int g = 0; g < G; ++g)
        for (int c = 0; c < C; ++c) {
            WD(&v)[16] = *reinterpret_cast<WD(*)[16]>(regs[g * C + c].data());
            tile_get<E, CP>(sm.data(), v, g, c, lo_to);
        }
}

template <int KIND, int A, int LAZY, bool PRE, bool POST, bool TR, typename WD, typename TW>
void pass(const u64 *src, u64 *dst, const LimbConst &m, const TW *tab, const TW *elt, unsigned ncols, int *range_bad, WD) {
    typedef TileGeom<A, E> GM;
    constexpr bool FWD = (KIND == XF_NEG_FWD || KIND == XF_CYC_FWD);
    constexpr int lo_in = FWD ? GM::lo(0) : GM::lo(GM::NS - 1);
    constexpr int lo_out = FWD ? GM::lo(GM::NS - 1) : GM::lo(0);
    const WD q = (WD)m.q, q2 = (WD)m.q2;
    std::vector<WD> sm((size_t)(1 << A) * CP);
    for (unsigned c0 = 0; c0 < ncols; c0 += C) {
        std::vector<std::array<WD, 16>> regs(GM::G * C);
        for (int g = 0; g < GM::G; ++g)
            for (int c = 0; c < C; ++c)
                for (int k = 0; k < 16; ++k) {
                    size_t off = (size_t)tile_idx<E>(g, k, lo_in) * ncols + c0 + c;
                    WD x = (WD)src[off];
                    if (PRE) x = mul_tw<LAZY>(x, elt[off], q);
                    regs[g * C + c][k] = x;
                }
        if (FWD) {
            step_all<KIND, A, 0, LAZY>(regs, tab, q, q2);
            if (GM::NS >= 2) {
                exchange<A, WD>(regs, sm, GM::lo(0), GM::lo(1));
                step_all<KIND, A, (GM::NS >= 2 ? 1 : 0), LAZY>(regs, tab, q, q2);
            }
            if (GM::NS >= 3) {
                exchange<A, WD>(regs, sm, GM::lo(1), GM::lo(2));
                step_all<KIND, A, (GM::NS >= 3 ? 2 : 0), LAZY>(regs, tab, q, q2);
            }
        } else {
            step_all<KIND, A, GM::NS - 1, LAZY>(regs, tab, q, q2);
            if (GM::NS >= 2) {
                exchange<A, WD>(regs, sm, GM::lo(GM::NS - 1), GM::lo(GM::NS - 2));
                step_all<KIND, A, (GM::NS >= 2 ? GM::NS - 2 : 0), LAZY>(regs, tab, q, q2);
            }
            if (GM::NS >= 3) {
                exchange<A, WD>(regs, sm, GM::lo(1), GM::lo(0));
                step_all<KIND, A, 0, LAZY>(regs, tab, q, q2);
            }
        }
        constexpr bool CT_RANGE = (KIND == XF_NEG_FWD || KIND == XF_CYC_INV);
        for (int g = 0; g < GM::G; ++g)
            for (int c = 0; c < C; ++c)
                for (int k = 0; k < 16; ++k) {
                    int idx = tile_idx<E>(g, k, lo_out);
                    size_t off = (size_t)idx * ncols + c0 + c;
                    WD x = regs[g * C + c][k];
                    // lazy-range audit: CT kinds must stay below 4q, GS kinds below 2q
                    // (lazy8 negacyclic forward with bit-63 range management: any word is in range, see ct_bfly TOPBIT)
                    constexpr bool FREE_RANGE = (KIND == XF_NEG_FWD && CKKS_NEG_FWD_TOPBIT != 0);
                    if (LAZY == 2 && !FREE_RANGE && (u64)x >= (CT_RANGE ? 8 * (u64)q : 4 * (u64)q)) *range_bad = 1;
                    if (LAZY == 1 && (u64)x >= (CT_RANGE ? 4 * (u64)q : 2 * (u64)q)) *range_bad = 1;
                    if (!LAZY && x >= q) *range_bad = 1;
                    if (POST) x = mul_tw<LAZY>(x, elt[off], q);
                    if (TR) {
                        dst[(size_t)(c0 + c) * (1 << A) + idx] = (u64)x;
                    } else {
                        if (POST || !CT_RANGE) x = canon2<LAZY>(x, q);
                        else x = canon4<LAZY>(x, q, q2);
                        dst[off] = (u64)x;
                    }
                }
    }
}

template <int KIND, int LAZY, bool PRE, bool POST, bool TR, typename WD, typename TW>
void pass_a(int A, const u64 *src, u64 *dst, const LimbConst &m, const TW *tab, const TW *elt, unsigned ncols, int *bad, WD w) {
    switch (A) {
        case 4: pass<KIND, 4, LAZY, PRE, POST, TR>(src, dst, m, tab, elt, ncols, bad, w); break;
        case 5: pass<KIND, 5, LAZY, PRE, POST, TR>(src, dst, m, tab, elt, ncols, bad, w); break;
        case 6: pass<KIND, 6, LAZY, PRE, POST, TR>(src, dst, m, tab, elt, ncols, bad, w); break;
        case 7: pass<KIND, 7, LAZY, PRE, POST, TR>(src, dst, m, tab, elt, ncols, bad, w); break;
        case 8: pass<KIND, 8, LAZY, PRE, POST, TR>(src, dst, m, tab, elt, ncols, bad, w); break;
    }
}
template <int LAZY, typename WD, typename TW>
void run(const LimbConst &lc, const TW *P1, const TW *P1i, const TW *W2, const TW *W2i, const TW *TT, const TW *TTi, u64 n, int a1,
         int a2, u64 *d, int inverse, int *bad, WD w) {
    std::vector<u64> tmp(n);
    unsigned n1 = 1u << a1, n2 = 1u << a2;
    const TW *none = nullptr;
    if (!inverse) {
        pass_a<XF_NEG_FWD, LAZY, false, false, true>(a1, d, tmp.data(), lc, P1, none, n2, bad, w);
        pass_a<XF_CYC_FWD, LAZY, true, false, false>(a2, tmp.data(), d, lc, W2, TT, n1, bad, w);
    } else {
        pass_a<XF_CYC_INV, LAZY, false, true, true>(a2, d, tmp.data(), lc, W2i, TTi, n1, bad, w);
        pass_a<XF_NEG_INV, LAZY, false, false, false>(a1, tmp.data(), d, lc, P1i, none, n2, bad, w);
    }
}
}  // namespace

// One limb, four-step path.  data: N words in place (internal NTT order on the NTT side).
// force_strict != 0 runs the !LAZY code even for small moduli.  Returns 0, or 1 if a lazy-range
// invariant was violated, or -1 for an unsupported size.
extern "C" int emul_ntt_4step(uint64_t n, uint64_t q, uint64_t *data, int inverse, int force_strict) {
    int logn = 0;
    while (((u64)1 << logn) < n) ++logn;
    if (logn < 8 || logn > 16) return -1;
    int a1 = (logn + 1) / 2, a2 = logn - a1;
    std::vector<u64> mod{(u64)q}, psi{hm::find_primitive_root(q, 2 * n)};
    ht::HostTables H;
    ht::build_host_tables(n, logn, 2, a1, a2, mod, psi, H, false, true);
    int bad = 0;
    // force_strict: 0 = the mode the table builder selects (2 for q < 2^61, 1 for q < 2^62, else 0),
    // 1 = strict, 2 = Harvey lazy where the modulus allows it
    int mode = force_strict == 1 ? 0 : (force_strict == 2 ? (H.lazy ? 1 : 0) : H.lazy);
    if (mode == 2)
        run<2>(H.lc[0], H.P1.data(), H.P1i.data(), H.W2.data(), H.W2i.data(), H.TT.data(), H.TTi.data(), n, a1, a2, (u64 *)data, inverse, &bad, (u64)0);
    else if (mode == 1)
        run<1>(H.lc[0], H.P1.data(), H.P1i.data(), H.W2.data(), H.W2i.data(), H.TT.data(), H.TTi.data(), n, a1, a2, (u64 *)data, inverse, &bad, (u64)0);
    else
        run<0>(H.lc[0], H.P1.data(), H.P1i.data(), H.W2.data(), H.W2i.data(), H.TT.data(), H.TTi.data(), n, a1, a2, (u64 *)data, inverse, &bad, (u64)0);
    return bad;
}
// Same with 32-bit words (q < 2^31).  Returns -2 if the table builder does not select the 32-bit path.
extern "C" int emul_ntt_4step32(uint64_t n, uint64_t q, uint64_t *data, int inverse, int force_strict) {
    int logn = 0;
    while (((u64)1 << logn) < n) ++logn;
    if (logn < 8 || logn > 16) return -1;
    int a1 = (logn + 1) / 2, a2 = logn - a1;
    std::vector<u64> mod{(u64)q}, psi{hm::find_primitive_root(q, 2 * n)};
    ht::HostTables H;
    ht::build_host_tables(n, logn, 2, a1, a2, mod, psi, H, true);
    if (!H.w32) return -2;
    int bad = 0;
    if (H.lazy && !force_strict)
        run<1>(H.lc[0], H.P1_32.data(), H.P1i_32.data(), H.W2_32.data(), H.W2i_32.data(), H.TT_32.data(), H.TTi_32.data(), n, a1, a2, (u64 *)data, inverse, &bad, (u32)0);
    else
        run<0>(H.lc[0], H.P1_32.data(), H.P1i_32.data(), H.W2_32.data(), H.W2i_32.data(), H.TT_32.data(), H.TTi_32.data(), n, a1, a2, (u64 *)data, inverse, &bad, (u32)0);
    return bad;
}
// Internal position of natural slot k (same formula as ntt_pos in kernels.cuh).
extern "C" uint32_t emul_ntt_pos(uint32_t k, int a1, int a2) {
    unsigned k1 = k & ((1u << a1) - 1), k2 = k >> a1;
    return (hm::brv(k2, a2) << a1) | hm::brv(k1, a1);
}
extern "C" uint64_t emul_psi(uint64_t n, uint64_t q) { return hm::find_primitive_root(q, 2 * n); }
// Scalar helpers of modarith.cuh for property tests.
extern "C" uint64_t emul_mulmod(uint64_t a, uint64_t b, uint64_t q) {
    std::vector<u64> mod{(u64)q}, psi{1};
    ht::HostTables H;
    ht::build_host_tables(1, 0, 1, 0, 0, mod, psi, H, false);
    return mulmod(a, b, H.lc[0]);
}
extern "C" uint64_t emul_mulmod_add(uint64_t a, uint64_t b, uint64_t c, uint64_t q) {
    std::vector<u64> mod{(u64)q}, psi{1};
    ht::HostTables H;
    ht::build_host_tables(1, 0, 1, 0, 0, mod, psi, H, false);
    return mulmod_add(a, b, c, H.lc[0]);
}
extern "C" uint64_t emul_barrett_word(uint64_t a, uint64_t q) {
    std::vector<u64> mod{(u64)q}, psi{1};
    ht::HostTables H;
    ht::build_host_tables(1, 0, 1, 0, 0, mod, psi, H, false);
    return barrett_word(a, H.lc[0]);
}

// shoup_lazy8 (approximate quotient): value and range for property tests.
extern "C" uint64_t emul_shoup_lazy8(uint64_t x, uint64_t w, uint64_t q) {
    tw_t t = ht::mk_tw(w, q);
    return shoup_lazy8((u64)x, t, (u64)(0 - q));
}
// Shared-memory tile layouts (ntt_tile.cuh tile_addr) for the bank-conflict check of tests/test_emul.py: word
// index of (row r, column c) in ks_pass2's tile shape (E = 3, C = 4); swz = 1 for the bit-weighted layout of its
// digit loop, 0 for the padded layout.
extern "C" int emul_tile_addr(int swz, int r, int c) { return swz ? tile_addr<3, 5, 1>(r, c) : tile_addr<3, 5, 0>(r, c); }
// Row permutation of the resident gadget keys (ntt_tile.cuh perm_row) for tests/test_emul.py.
extern "C" uint64_t emul_perm_row(uint64_t row, int a2, int pe) { return (uint64_t)perm_row((size_t)row, a2, pe); }

// ---- csrc/arena.hpp: the device-memory arena's bookkeeping, driven on the CPU with fake segments -----------------
#include <set>

#include "../../toy-heaan-ckks_b200/csrc/arena.hpp"
// Random take / give stress with `ops` operations over sizes between 1 MiB and `max_mib` MiB.  Segments are fake
// address ranges (never touched).  Checks after every operation: live ranges are disjoint, aligned and inside a
// segment; free + live bytes == segment bytes; no two free ranges of one segment are adjacent (coalescing);
// every give() of a live pointer succeeds and a second one fails.  With `cap_gib` > 0 the fake device refuses segments
// beyond that total, which exercises the exact-size and release-and-retry fallbacks.  Returns 0, or a failure code;
// *peak_bytes = the largest total segment size seen, *new_segments = segments requested after the first `ops / 2`
// operations (a steady-state workload must not need any).
extern "C" int emul_arena_stress(uint64_t seed, int ops, int max_mib, int cap_gib, uint64_t *peak_bytes, int *new_segments) {
    Arena ar;
    uint64_t next_base = (uint64_t)1 << 40, seg_total = 0, peak = 0;
    std::map<char *, size_t> segs;
    int late_segs = 0, op = 0;
    auto seg_alloc = [&](size_t bytes) -> void * {
        if (cap_gib > 0 && seg_total + bytes > ((uint64_t)cap_gib << 30)) return nullptr;
        char *b = (char *)next_base;
        next_base += bytes + ((uint64_t)1 << 30);  // a gap: separate segments are never contiguous here
        segs[b] = bytes;
        seg_total += bytes;
        if (op > ops / 2) ++late_segs;
        return b;
    };
    auto seg_free = [&](void *b) {
        seg_total -= segs[(char *)b];
        segs.erase((char *)b);
    };
    std::map<char *, size_t> live;  // our own record of what is out
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + 1;
    auto rnd = [&]() {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        return s;
    };
    for (op = 0; op < ops; ++op) {
        const bool do_take = live.empty() || (rnd() % 100 < 52 && live.size() < 24);
        if (do_take) {
            size_t bytes = ((size_t)1 << 20) + rnd() % ((size_t)max_mib << 20);
            void *p = ar.take(bytes, seg_alloc, seg_free);
            if (!p) {
                if (cap_gib == 0) return 1;  // an unlimited device cannot run out
                continue;
            }
            const size_t need = (bytes + Arena::ALIGN - 1) / Arena::ALIGN * Arena::ALIGN;
            char *c = (char *)p;
            if (((uint64_t)c & (Arena::ALIGN - 1)) != 0) return 2;
            bool inside = false;
            for (auto &kv : segs) inside |= (c >= kv.first && c + need <= kv.first + kv.second);
            if (!inside) return 3;
            auto nx = live.lower_bound(c);
            if (nx != live.end() && nx->first < c + need) return 4;  // overlaps the next live range
            if (nx != live.begin()) {
                auto pv = std::prev(nx);
                if (pv->first + pv->second > c) return 5;
            }
            live[c] = need;
        } else {
            auto it = live.begin();
            std::advance(it, rnd() % live.size());
            if (!ar.give(it->first)) return 6;
            if (ar.give(it->first)) return 7;  // double free must be refused
            live.erase(it);
        }
        size_t live_bytes = 0;
        for (auto &kv : live) live_bytes += kv.second;
        if (ar.total_bytes() != seg_total) return 8;
        if (ar.free_bytes() + live_bytes != seg_total) return 9;
        if (ar.live_ranges() != live.size()) return 10;
        // coalescing: free ranges <= live ranges + segments (between two live ranges / segment ends at most one free range)
        if (ar.free_ranges() > live.size() + segs.size()) return 11;
        if (seg_total > peak) peak = seg_total;
    }
    for (auto &kv : live)
        if (!ar.give(kv.first)) return 12;
    ar.release_free_segments(seg_free);
    if (ar.total_bytes() != 0 || seg_total != 0 || ar.segments() != 0) return 13;
    if (peak_bytes) *peak_bytes = peak;
    if (new_segments) *new_segments = late_segs;
    return 0;
}
// A horner_chain-like schedule: per level four polynomials of a level-dependent size are taken and the previous level's
// given back; `passes` passes.  Returns the number of segments requested after the first pass (must be 0).
extern "C" int emul_arena_chain(int levels, int passes, uint64_t unit_bytes) {
    Arena ar;
    uint64_t next_base = (uint64_t)1 << 40;
    int segs_after_first = 0, pass = 0;
    auto seg_alloc = [&](size_t bytes) -> void * {
        char *b = (char *)next_base;
        next_base += bytes + ((uint64_t)1 << 30);
        if (pass > 0) ++segs_after_first;
        return b;
    };
    auto seg_free = [&](void *) {};
    for (pass = 0; pass < passes; ++pass) {
        std::vector<void *> prev;
        for (int l = levels; l >= 2; --l) {
            std::vector<void *> cur;
            for (int t = 0; t < 4; ++t) cur.push_back(ar.take((size_t)l * unit_bytes, seg_alloc, seg_free));
            for (void *p : prev) ar.give(p);
            prev = cur;
        }
        for (void *p : prev) ar.give(p);
    }
    return segs_after_first;
}

// ---- auxiliary-basis gadget product (csrc/aux_crt.cuh): constants and the per-coefficient reconstruction ----------------
#include "../../toy-heaan-ckks_b200/csrc/aux_crt.cuh"
// The auxiliary primes the library picks for (n, moduli): returns K and writes primes[0..K-1] (0: not possible).
extern "C" int emul_aux_primes(uint64_t n, const uint64_t *moduli, int l, uint64_t *primes) {
    int logn = 0;
    while (((u64)1 << logn) < n) ++logn;
    AuxHost H;
    if (!aux_host_build(n, logn, std::vector<u64>(moduli, moduli + l), H)) return 0;
    for (int k = 0; k < H.K; ++k) primes[k] = H.primes[k];
    return H.K;
}
template <int K>
static void crt_all(const AuxHost &H, const std::vector<LimbConst> &lc, int l, uint64_t count, const uint64_t *res, uint64_t *out, int *negs) {
    for (int j = 0; j < l; ++j)
        for (uint64_t e = 0; e < count; ++e) {
            u32 v[K];
            for (int k = 0; k < K; ++k) v[k] = (u32)res[((size_t)j * K + k) * count + e];
            u64 hps = 0;
            if constexpr (K <= 5) hps = aux_image_hps<K>(v, H.cc, H.mstar.data() + (size_t)j * K, H.tp.data() + (size_t)j * (AUX_MAX_K + 1), lc[j]);
            aux_garner<K>(v, H.cc);
            const bool neg = aux_negative<K>(v, H.cc);
            out[(size_t)j * count + e] = aux_image<K>(v, neg, H.mix.data() + (size_t)j * K, H.pmod[j], lc[j]);
            if (negs) negs[(size_t)j * count + e] = neg ? 1 : 0;
            if (K <= 5 && hps != out[(size_t)j * count + e]) negs ? (void)(negs[(size_t)j * count + e] = -1) : (void)0;  // both reconstructions must agree
        }
}
// res: [l][K][count] residues mod the auxiliary primes of one integer per (j, e); out: [l][count] its centred value
// mod moduli[j] (the device code of aux_crt_kernel, element by element); negs (optional): the sign decisions.
extern "C" int emul_aux_crt(uint64_t n, const uint64_t *moduli, int l, uint64_t count, const uint64_t *res, uint64_t *out, int *negs) {
    int logn = 0;
    while (((u64)1 << logn) < n) ++logn;
    std::vector<u64> mod(moduli, moduli + l), psi(l, 1);
    AuxHost H;
    if (!aux_host_build(n, logn, mod, H)) return -1;
    ht::HostTables T;
    ht::build_host_tables(1, 0, 1, 0, 0, mod, psi, T, false);
    switch (H.K) {
        case 2: crt_all<2>(H, T.lc, l, count, res, out, negs); break;
        case 3: crt_all<3>(H, T.lc, l, count, res, out, negs); break;
        case 4: crt_all<4>(H, T.lc, l, count, res, out, negs); break;
        case 5: crt_all<5>(H, T.lc, l, count, res, out, negs); break;
        case 6: crt_all<6>(H, T.lc, l, count, res, out, negs); break;
        case 7: crt_all<7>(H, T.lc, l, count, res, out, negs); break;
        case 8: crt_all<8>(H, T.lc, l, count, res, out, negs); break;
        default: return -2;
    }
    return H.K;
}
// aux_reduce_sum (the final `mod p` of aux_mac_kernel: quotient estimated in double precision, exact remainder).
extern "C" uint32_t emul_aux_reduce_sum(uint64_t s, uint32_t p) { return aux_reduce_sum((u64)s, (u32)p, 1.0 / (double)p); }
