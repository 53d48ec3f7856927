// tables_host.hpp -- host-side construction of every device table from (N, moduli, psi).
// The device image of NttTable::new (basis.rs:21-84): the same psi (find_primitive_root), but laid
// out for the merged / four-step transforms instead of the reference's natural-power arrays.
#pragma once
#include <cstring>
#include <vector>

#include "host_math.hpp"
#include "modarith.cuh"

namespace ht {
struct HostTables {
    int lazy = 1;  // 0 strict, 1 Harvey lazy, 2 lazy8 (approximate Shoup quotient; 64-bit words, q < 2^61)
    bool digit_reduce = true;
    bool w32 = false;  // all q < 2^31: 32-bit tables are filled instead of the 64-bit four-step ones
    std::vector<tw32_t> ql32, P1_32, P1i_32, W2_32, W2i_32, TT_32, TTi_32, TTt_32;
    size_t w2_stride = 1;
    std::vector<LimbConst> lc;
    std::vector<tw_t> ql;                   // [L][L] q_last^-1 mod q_i
    std::vector<tw_t> psi, psii, ninv;      // small path
    std::vector<tw_t> P1, P1i, W2, W2i, TT, TTi, TTt;  // four-step (TTt: TT in [rho][j2] layout)
};
inline tw_t mk_tw(u64 w, u64 q) {
    tw_t t;
    t.w = w;
    t.ws = hm::shoup_of(w, q);
    return t;
}
inline tw32_t mk_tw32(u64 w, u64 q) {
    tw32_t t;
    t.w = (u32)w;
    t.ws = (u32)((w << 32) / q);
    return t;
}
inline void build_host_tables(u64 n, int logn, int path, int a1, int a2, const std::vector<u64> &moduli,
                              const std::vector<u64> &psis, HostTables &H, bool allow_w32 = true, bool allow_lazy8 = true) {
    const size_t L = moduli.size();
    H.lc.resize(L);
    u64 qmin = ~0ull, qmax = 0;
    H.lazy = 1;
    for (size_t j = 0; j < L; ++j) {
        u64 q = moduli[j];
        LimbConst &m = H.lc[j];
        memset(&m, 0, sizeof(m));
        m.q = q;
        m.q2 = 2 * q;
        m.bar = (u64)((((hm::u128)1) << 64) / q);
        m.c64 = (u64)((((hm::u128)1) << 64) % q);
        m.c64s = hm::shoup_of(m.c64, q);
        qmin = q < qmin ? q : qmin;
        qmax = q > qmax ? q : qmax;
    }
    H.lazy = (qmax >> 61) == 0 ? (allow_lazy8 && path == 2 ? 2 : 1) : ((qmax >> 62) == 0 ? 1 : 0);
    H.w32 = allow_w32 && path == 2 && (qmax >> 31) == 0;
    if (H.w32) H.lazy = (qmax >> 30) == 0 ? 1 : 0;  // 32-bit Harvey butterflies need 4q < 2^32
    // A digit x < q_i enters the lazy forward transform mod q_j unreduced iff x < 4 q_j.
    H.digit_reduce = !(H.lazy && (qmax >> 2) < qmin);
    H.ql.resize(L * L);
    for (size_t last = 0; last < L; ++last)
        for (size_t i = 0; i < L; ++i) {
            u64 qi = moduli[i];
            H.ql[last * L + i] = (i == last) ? mk_tw(0, qi) : mk_tw(hm::inv_mod(moduli[last] % qi, qi), qi);
        }
    if (H.w32) {
        H.ql32.resize(L * L);
        for (size_t k = 0; k < L * L; ++k) H.ql32[k] = mk_tw32(H.ql[k].w, moduli[k % L]);
    }
    if (path == 1) {
        H.psi.resize(L * n);
        H.psii.resize(L * n);
        H.ninv.resize(L);
        for (size_t j = 0; j < L; ++j) {
            u64 q = moduli[j], ps = psis[j], psinv = hm::inv_mod(ps, q);
            std::vector<u64> pw(n), pwi(n);
            pw[0] = pwi[0] = 1;
            for (u64 i = 1; i < n; ++i) {
                pw[i] = hm::mul_mod(pw[i - 1], ps, q);
                pwi[i] = hm::mul_mod(pwi[i - 1], psinv, q);
            }
            for (u64 i = 0; i < n; ++i) {
                unsigned r = hm::brv((unsigned)i, logn);
                H.psi[j * n + i] = mk_tw(pw[r], q);
                H.psii[j * n + i] = mk_tw(pwi[r], q);
            }
            H.ninv[j] = mk_tw(hm::inv_mod(n % q, q), q);
        }
        return;
    }
    const u64 n1 = (u64)1 << a1, n2 = (u64)1 << a2;
    H.w2_stride = n2 / 2 ? n2 / 2 : 1;
    H.P1.resize(L * n1);
    H.P1i.resize(L * n1);
    H.W2.resize(L * H.w2_stride);
    H.W2i.resize(L * H.w2_stride);
    H.TT.resize(L * n);
    H.TTi.resize(L * n);
    H.TTt.resize(L * n);
    for (size_t j = 0; j < L; ++j) {
        u64 q = moduli[j], ps = psis[j], psinv = hm::inv_mod(ps, q);
        u64 ninv = hm::inv_mod(n % q, q);
        u64 psi1 = hm::pow_mod(ps, n2, q), psi1i = hm::pow_mod(psinv, n2, q);
        std::vector<u64> pw(n1), pwi(n1);
        pw[0] = pwi[0] = 1;
        for (u64 i = 1; i < n1; ++i) {
            pw[i] = hm::mul_mod(pw[i - 1], psi1, q);
            pwi[i] = hm::mul_mod(pwi[i - 1], psi1i, q);
        }
        for (u64 i = 0; i < n1; ++i) {
            unsigned r = hm::brv((unsigned)i, a1);
            H.P1[j * n1 + i] = mk_tw(pw[r], q);
            H.P1i[j * n1 + i] = mk_tw(pwi[r], q);
        }
        u64 w2 = hm::pow_mod(ps, 2 * n1, q), w2i = hm::pow_mod(psinv, 2 * n1, q);
        u64 c = 1, ci = 1;
        for (u64 e = 0; e < H.w2_stride; ++e) {
            H.W2[j * H.w2_stride + e] = mk_tw(c, q);
            H.W2i[j * H.w2_stride + e] = mk_tw(ci, q);
            c = hm::mul_mod(c, w2, q);
            ci = hm::mul_mod(ci, w2i, q);
        }
        // TT[j2][rho] = psi^(j2 * (2 brv(rho) + 1)); TTi = its inverse times N^-1
        for (u64 rho = 0; rho < n1; ++rho) {
            u64 k1 = hm::brv((unsigned)rho, a1);
            u64 gen = hm::pow_mod(ps, 2 * k1 + 1, q), geni = hm::pow_mod(psinv, 2 * k1 + 1, q);
            u64 v = 1, vi = ninv;
            for (u64 j2 = 0; j2 < n2; ++j2) {
                H.TT[j * n + j2 * n1 + rho] = mk_tw(v, q);
                H.TTt[j * n + rho * n2 + j2] = H.TT[j * n + j2 * n1 + rho];
                H.TTi[j * n + j2 * n1 + rho] = mk_tw(vi, q);
                v = hm::mul_mod(v, gen, q);
                vi = hm::mul_mod(vi, geni, q);
            }
        }
    }
    if (H.w32) {
        auto narrow = [&](std::vector<tw_t> &src, std::vector<tw32_t> &dst, size_t per_limb) {
            dst.resize(src.size());
            for (size_t k = 0; k < src.size(); ++k) dst[k] = mk_tw32(src[k].w, moduli[k / per_limb]);
            std::vector<tw_t>().swap(src);
        };
        narrow(H.P1, H.P1_32, n1);
        narrow(H.P1i, H.P1i_32, n1);
        narrow(H.W2, H.W2_32, H.w2_stride);
        narrow(H.W2i, H.W2i_32, H.w2_stride);
        narrow(H.TT, H.TT_32, n);
        narrow(H.TTi, H.TTi_32, n);
        narrow(H.TTt, H.TTt_32, n);
    }
}
}  // namespace ht
