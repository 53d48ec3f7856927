// aux_ks.inl -- host side of the exact multi-modular gadget key-switch (kernels and the algorithm: aux_ks.cuh).
// Included by ckks_b200.cu after the four-step launchers.
//
// Used for mul_ciphertexts_gadget (engine.rs:474-545) and rotate_ciphertext (engine.rs:412-463) on the 64-bit four-step
// path once the basis is deep enough for L*K + 2*L*K 32-bit transforms to beat L*(L-1) 64-bit ones (aux_wanted);
// shallower levels, the 32-bit word path and the limb-sharded mode keep the per-(digit, target) pipeline of
// ks_pass1 / ks_pass2.

struct AuxKs {
    std::shared_ptr<Tables> X;  // tables of the auxiliary primes (32-bit word path, Harvey lazy)
    int K = 0;
    tw_t *d_mix = nullptr;    // [L][K] prod_{m<k} p_m mod q_j
    u64 *d_pmod = nullptr;    // [L] P mod q_j
    u64 *d_mstar = nullptr;   // [L][K] (P / p_k) mod q_j
    u64 *d_tp = nullptr;      // [L][AUX_MAX_K + 1] t P mod q_j
    AuxCrtConst cc;           // primes, Garner inverses and floor(P/2) digits as a kernel parameter
};
static void destroy_aux(AuxKs *a) {
    if (!a) return;
    for (void *p : {(void *)a->d_mix, (void *)a->d_pmod, (void *)a->d_mstar, (void *)a->d_tp})
        if (p) cudaFree(p);
    delete a;
}

static int g_ks_aux = 1;         // 0 never, 1 when it pays (aux_wanted), 2 whenever it is possible (tests)
static int g_ks_aux_min_l = 10;  // measured crossover on B200 for 61-bit primes at N = 2^16 (horner_chain per-level times)
extern "C" int ckks_set_ks_aux(int mode) {
    if (mode < 0 || mode > 2) return CKKS_BAD_ARGUMENT;
    g_ks_aux = mode;
    return CKKS_OK;
}
static bool aux_possible(const Tables &T) { return T.path == 2 && !T.w32 && T.a1 >= 4 && T.a2 >= 4 && T.L <= (size_t)AUX_MAX_L; }
static bool aux_wanted(const Tables &T, size_t L) {
    if (!aux_possible(T) || g_ks_aux == 0) return false;
    return g_ks_aux == 2 || L >= (size_t)g_ks_aux_min_l;
}

// The auxiliary tables of a context tree, built on first use.
static int aux_get(const Tables &Tc, AuxKs **out) {
    Tables &T = const_cast<Tables &>(Tc);
    std::lock_guard<std::mutex> lk(T.aux_mu);
    if (T.aux) {
        *out = T.aux;
        return CKKS_OK;
    }
    AuxHost H;
    if (!aux_host_build(T.n, T.logn, T.moduli, H)) return CKKS_UNSUPPORTED;
    const int K = H.K;
    const std::vector<u64> &primes = H.primes;
    std::unique_ptr<AuxKs, void (*)(AuxKs *)> A(new AuxKs(), destroy_aux);
    A->K = K;
    A->X = std::make_shared<Tables>();
    Tables &X = *A->X;
    X.device = T.device;
    X.n = T.n;
    X.logn = T.logn;
    X.path = 2;
    X.a1 = T.a1;
    X.a2 = T.a2;
    X.L = K;
    X.moduli = primes;
    for (u64 p : primes) X.psi.push_back(hm::find_primitive_root(p, 2 * T.n));
    TRY(build_tables(X, true, false));  // the auxiliary primes always use the 32-bit word path
    if (!X.w32 || X.lazy != 1) return CKKS_UNSUPPORTED;
    A->cc = H.cc;
    TRY(upload_vec(&A->d_mix, H.mix));
    TRY(upload_vec(&A->d_pmod, H.pmod));
    TRY(upload_vec(&A->d_mstar, H.mstar));
    TRY(upload_vec(&A->d_tp, H.tp));
    T.aux = A.release();
    *out = T.aux;
    return CKKS_OK;
}

// ---- 32-bit-in / 32-bit-out passes over (limb, auxiliary prime) pairs ------------------------------------------------
enum { AUX_FWD1, AUX_FWD2, AUX_FWD2_NOPRE, AUX_INV2, AUX_INV1 };
template <int KIND, int A, bool PRE, bool POST, bool TR, int CW>
static int launch_aux_pass_c(const char *name, dim3 grid, cudaStream_t s, const PassArgs &a) {
    constexpr int E = 4, C = CW <= (1 << A) ? CW : (1 << A);
    grid.x = a.ncols / C;
    const size_t smem = (size_t)(1 << A) * (C + 1) * sizeof(u32);
    const int block = C << (A - E);
    if (FIX_MATCH(A, a.N))
        KL(name, (ntt_pass_kernel<u32, KIND, A, E, C, 1, PRE, POST, TR, false, false, FIX_OF(A), true><<<grid, block, smem, s>>>(a)));
    else
        KL(name, (ntt_pass_kernel<u32, KIND, A, E, C, 1, PRE, POST, TR, false, false, 0, true><<<grid, block, smem, s>>>(a)));
    return CKKS_OK;
}
// Column tile: 16 words = 64-byte row segments; the last inverse pass, whose canonical rows are the WRITTEN side, does better
// with 32 (128-byte segments: 16.7 -> 12.9 ms per 512 ct-mults at cfg4) wherever the limb has that many columns; the passes
// that READ canonical rows or run in place do not (inverse pass 2: 15.0 -> 16.5, forward pass 2: 7.6 -> 8.6).
template <int KIND, int A, bool PRE, bool POST, bool TR>
static int launch_aux_pass_a(const char *name, dim3 grid, cudaStream_t s, const PassArgs &a) {
    if (KIND == XF_NEG_INV && A >= 5 && a.ncols >= 32) return launch_aux_pass_c<KIND, A, PRE, POST, TR, 32>(name, grid, s, a);
    return launch_aux_pass_c<KIND, A, PRE, POST, TR, 16>(name, grid, s, a);
}
template <int KIND, bool PRE, bool POST, bool TR>
static int launch_aux_pass(const char *name, int A, dim3 grid, cudaStream_t s, const PassArgs &a) {
    DISPATCH_A(A, return (launch_aux_pass_a<KIND, AA, PRE, POST, TR>(name, grid, s, a)));
    return CKKS_UNSUPPORTED;
}
// `limbs` (limb, prime) pairs per polynomial, `nb` polynomials; the tables of pair l are those of prime (l / div) % K.
static int aux_pass(const Tables &T, const AuxKs &A, int which, size_t nb, size_t limbs, int div, const void *src, void *dst) {
    if (!nb || !limbs) return CKKS_OK;
    const Tables &X = *A.X;
    const unsigned n1 = 1u << X.a1, n2 = 1u << X.a2;
    cudaStream_t s = S(T);
    for (size_t b0 = 0; b0 < nb; b0 += 32768)
        for (size_t l0 = 0; l0 < limbs; l0 += 32768) {
            const size_t cb = nb - b0 < 32768 ? nb - b0 : 32768, cl = limbs - l0 < 32768 ? limbs - l0 : 32768;
            PassArgs a;
            memset(&a, 0, sizeof(a));
            a.lc = X.d_lc;
            a.L = (int)limbs;
            a.dstL = (int)limbs;
            a.N = X.n;
            a.limb0 = (int)l0;
            a.tab_div = div;
            a.tab_mod = A.K;
            a.src = (const char *)src + b0 * limbs * X.n * 4;
            a.dst = (char *)dst + b0 * limbs * X.n * 4;
            dim3 g(1, (unsigned)cl, (unsigned)cb);
            switch (which) {
                case AUX_FWD1:
                    a.tab = X.d_P1;
                    a.tab_stride = n1;
                    a.ncols = n2;
                    TRY((launch_aux_pass<XF_NEG_FWD, false, false, true>("aux_fwd_pass1", X.a1, g, s, a)));
                    break;
                case AUX_FWD2:
                case AUX_FWD2_NOPRE:
                    a.tab = X.d_W2;
                    a.tab_stride = X.w2_stride;
                    a.elt = X.d_TT;
                    a.ncols = n1;
                    if (which == AUX_FWD2) TRY((launch_aux_pass<XF_CYC_FWD, true, false, false>("aux_fwd_pass2", X.a2, g, s, a)));
                    else TRY((launch_aux_pass<XF_CYC_FWD, false, false, false>("aux_fwd_pass2", X.a2, g, s, a)));
                    break;
                case AUX_INV2:
                    a.tab = X.d_W2i;
                    a.tab_stride = X.w2_stride;
                    a.elt = X.d_TTi;
                    a.ncols = n1;
                    TRY((launch_aux_pass<XF_CYC_INV, false, true, true>("aux_inv_pass2", X.a2, g, s, a)));
                    break;
                case AUX_INV1:
                    a.tab = X.d_P1i;
                    a.tab_stride = n1;
                    a.ncols = n2;
                    TRY((launch_aux_pass<XF_NEG_INV, false, false, false>("aux_inv_pass1", X.a1, g, s, a)));
                    break;
            }
        }
    return CKKS_OK;
}

// ---- keys -------------------------------------------------------------------------------------------------------------
// The auxiliary form of a gadget key from its coefficient-domain polynomials a, b: [L(i)][L(j)][N] u64 (device).
static int ksk_add_aux(const Tables &T, ckks_ksk *k, const u64 *a_coeff, const u64 *b_coeff) {
    const size_t L = k->ctx->L;
    if (k->digits != L || !aux_wanted(T, L)) return CKKS_OK;
    AuxKs *A;
    if (aux_get(T, &A) != CKKS_OK) return CKKS_OK;  // no auxiliary basis for this ring: the key keeps the standard form only
    const size_t words = (size_t)A->K * L * L * T.n;
    u32 *tmp = nullptr;
    CU(pool_malloc(T, (void **)&tmp, words * 4));
    int rc = CKKS_OK;
    for (int h = 0; h < 2 && rc == CKKS_OK; ++h) {
        u32 *dst = nullptr;
        if (pool_malloc(T, (void **)&dst, words * 4) != cudaSuccess) {
            rc = cuda_fail(cudaGetLastError(), "auxiliary key");
            break;
        }
        (h ? k->xa : k->xb) = dst;
        KLV("aux_key_reduce", (aux_key_reduce_kernel<<<ew_grid(words), 256, 0, S(T)>>>(h ? a_coeff : b_coeff, dst, A->X->d_lc, (int)L, A->K, T.logn)));
        // K * L * L (prime, target, digit) rows; the prime is the slowest index
        rc = aux_pass(T, *A, AUX_FWD1, 1, (size_t)A->K * L * L, (int)(L * L), dst, tmp);
        if (rc == CKKS_OK) rc = aux_pass(T, *A, AUX_FWD2, 1, (size_t)A->K * L * L, (int)(L * L), tmp, dst);
    }
    dev_free(T, tmp);
    if (rc == CKKS_OK && cudaPeekAtLastError() != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "auxiliary key");
    if (rc == CKKS_OK) k->aux_k = A->K;
    return rc;
}

// ---- mul_ciphertexts_gadget ------------------------------------------------------------------------------------------
static size_t aux_chunk(const Tables &T, const AuxKs &A, size_t L, size_t batch) {
    const size_t per = 4 * (size_t)A.K * L * T.n * 4;  // x, rb, ra and the transposed intermediate
    size_t c = (g_ks_scratch_mib << 20) / per;
    if (c < 1) c = 1;
    if (c > 32768) c = 32768;
    return c < batch ? c : batch;
}
// out0 / out1: [cs][L][N], or with `last` ([2][cs][N] scratch) the rescaled [cs][L-1][N].
static int aux_keyswitch(const Tables &T, const AuxKs &A, size_t L, size_t cs, const u64 *digits, const ckks_ksk *key, const u64 *add0,
                         const u64 *add1, u32 *scr, u64 *out0, u64 *out1, u64 *last) {
    const Tables &X = *A.X;
    const size_t n = T.n, K = (size_t)A.K;
    const size_t W = cs * K * L * n;
    u32 *xs = scr, *rb = scr + W, *ra = scr + 2 * W, *rt = scr + 3 * W;
    cudaStream_t s = S(T);
    {  // first pass of NTT_{p_k}(alpha_i mod p_k): ks_pass1 with the auxiliary primes as the "target limbs"
        KsArgs a;
        memset(&a, 0, sizeof(a));
        a.digits = digits;
        a.scratch = xs;
        a.lc = X.d_lc;
        a.P1 = X.d_P1;
        a.TTt = X.d_TTt;
        a.L = (int)K;
        a.Ld = (int)L;
        a.joff = 0;
        a.jstep = 1;
        a.j0 = 0;
        a.dig_ct_stride = L * n;
        a.dig_limb_stride = n;
        a.a1 = X.a1;
        a.a2 = X.a2;
        a.N = n;
        dim3 g1(1u << X.a2, (unsigned)(K * L), (unsigned)cs);
        DISPATCH_A(X.a1, TRY((launch_ks1_w<u32, AA>(1, true, false, g1, s, a))));
    }
    TRY(aux_pass(T, A, AUX_FWD2_NOPRE, cs, K * L, (int)L, xs, xs));  // in place: a CTA owns its column tile
    {
        AuxMacArgs m;
        AuxMacMap map;
        memset(&map, 0, sizeof(map));
        if (!make_tile_map(map.x, xs, 4, n, L, cs * K, 32)) {
            g_err = "auxiliary-basis key-switch: cannot encode the TMA descriptor of the digit transforms";
            return CKKS_CUDA_ERROR;
        }
        m.kb = key->xb;
        m.ka = key->xa;
        m.rb = rb;
        m.ra = ra;
        m.alc = X.d_lc;
        m.L = (int)L;
        m.K = (int)K;
        m.logn = T.logn;
        m.cs = (unsigned)cs;
        // at most 12 target limbs (384 threads, 80 registers: two CTAs per SM, so the key loads at the start of one overlap
        // the arithmetic of the other -- measured 38.3 -> 36.6 ms per 512 ct-mults against 24 limbs and one CTA per SM)
#ifndef CKKS_AUX_MAC_JMAX
#define CKKS_AUX_MAC_JMAX 12
#endif
        const int jblocks = (int)((L + CKKS_AUX_MAC_JMAX - 1) / CKKS_AUX_MAC_JMAX);
        m.jb = (int)((L + jblocks - 1) / jblocks);
        dim3 g((unsigned)(n / 32), (unsigned)(K * jblocks)), blk(32, (unsigned)m.jb);
        switch ((L + 3) / 4) {
#define AUX_MAC_CASE(L4v) \
    case L4v: KL("aux_mac", (aux_mac_kernel<L4v><<<g, blk, 0, s>>>(m, map))); break
            AUX_MAC_CASE(1);
            AUX_MAC_CASE(2);
            AUX_MAC_CASE(3);
            AUX_MAC_CASE(4);
            AUX_MAC_CASE(5);
            AUX_MAC_CASE(6);
            AUX_MAC_CASE(7);
            AUX_MAC_CASE(8);
#undef AUX_MAC_CASE
            default: return CKKS_UNSUPPORTED;
        }
    }
    for (u32 *r : {rb, ra}) {
        TRY(aux_pass(T, A, AUX_INV2, cs, L * K, 1, r, rt));
        TRY(aux_pass(T, A, AUX_INV1, cs, L * K, 1, rt, r));
    }
    AuxCrtArgs c;
    memset(&c, 0, sizeof(c));
    c.rb = rb;
    c.ra = ra;
    c.add0 = add0;
    c.add1 = add1;
    c.lc = T.d_lc;
    c.mix = A.d_mix;
    c.pmod = A.d_pmod;
    c.mstar = A.d_mstar;
    c.tp = A.d_tp;
    c.L = (int)L;
    c.K = (int)K;
    c.logn = T.logn;
#ifndef CKKS_AUX_CRT_EPT
#define CKKS_AUX_CRT_EPT 4
#endif
    constexpr int EPT = CKKS_AUX_CRT_EPT;  // N >= 2^8 on this path; N = 2^8, 2^9 use one coefficient per thread
    auto launch = [&](bool rs, int ept) -> int {
        c.rows = cs * (size_t)c.nj;
        const size_t gy = c.rows < 32768 ? c.rows : 32768;
        dim3 g((unsigned)(n / (256 * (size_t)ept)), (unsigned)gy, (unsigned)((c.rows + gy - 1) / gy));
#define AUX_CRT_CASE(Kv)                                                                                      \
    case Kv:                                                                                                  \
        if (rs) {                                                                                             \
            if (ept == EPT) KL("aux_crt", (aux_crt_kernel<Kv, EPT, true><<<g, 256, 0, s>>>(c, A.cc)));        \
            else KL("aux_crt", (aux_crt_kernel<Kv, 1, true><<<g, 256, 0, s>>>(c, A.cc)));                     \
        } else {                                                                                              \
            if (ept == EPT) KL("aux_crt", (aux_crt_kernel<Kv, EPT, false><<<g, 256, 0, s>>>(c, A.cc)));       \
            else KL("aux_crt", (aux_crt_kernel<Kv, 1, false><<<g, 256, 0, s>>>(c, A.cc)));                    \
        }                                                                                                     \
        break
        switch (K) {
            AUX_CRT_CASE(2);
            AUX_CRT_CASE(3);
            AUX_CRT_CASE(4);
            AUX_CRT_CASE(5);
            AUX_CRT_CASE(6);
            AUX_CRT_CASE(7);
            AUX_CRT_CASE(8);
#undef AUX_CRT_CASE
            default: return CKKS_UNSUPPORTED;
        }
        return CKKS_OK;
    };
    const int ept = n >= 256 * EPT ? EPT : 1;
    if (!last) {  // every limb, no rescale
        c.out0 = out0;
        c.out1 = out1;
        c.j0 = 0;
        c.nj = (int)L;
        c.outL = (int)L;
        return launch(false, ept);
    }
    // rescale_ciphertext fused in: the last limb first, then the others with the rescale epilogue
    c.out0 = last;
    c.out1 = last + cs * n;
    c.j0 = (int)L - 1;
    c.nj = 1;
    c.outL = 1;
    TRY(launch(false, ept));
    c.last0 = last;
    c.last1 = last + cs * n;
    c.ql = T.d_qlinv + (L - 1) * T.L;
    c.out0 = out0;
    c.out1 = out1;
    c.j0 = 0;
    c.nj = (int)L - 1;
    c.outL = (int)L - 1;
    return launch(true, ept);
}

static int rescale_dev(const Tables &T, size_t L, size_t batch, const u64 *src, u64 *dst);
// mul_ciphertexts_gadget (+ rescale_ciphertext) with the auxiliary-basis key-switch; same contract as fused_mul_relin.
static int fused_mul_relin_aux(const Tables &T, size_t L, size_t batch, const u64 *a0, const u64 *a1, const u64 *b0, const u64 *b1,
                               const ckks_ksk *rlk, bool rescale, u64 *o0, u64 *o1) {
    AuxKs *Ap;
    TRY(aux_get(T, &Ap));
    const AuxKs &A = *Ap;
    if (rlk->aux_k != A.K) return CKKS_BAD_HANDLE;
    NvtxScope nvtx_call(rescale ? "ckks:mul_relin_rescale" : "ckks:mul_relin");
    const size_t n = T.n, cs_max = aux_chunk(T, A, L, batch);
    const size_t W = cs_max * L * n;
    u64 *A0 = nullptr, *A1 = nullptr, *B0 = nullptr, *B1 = nullptr, *TMP = nullptr, *SCR = nullptr, *LAST = nullptr;
    std::lock_guard<std::mutex> ws_lock(const_cast<Tables &>(T).ws_mu);
    int rc = ws_get(T, WS_A0, W * 8, &A0);
    if (rc == CKKS_OK) rc = ws_get(T, WS_A1, W * 8, &A1);
    if (rc == CKKS_OK) rc = ws_get(T, WS_B0, W * 8, &B0);
    if (rc == CKKS_OK) rc = ws_get(T, WS_B1, W * 8, &B1);
    if (rc == CKKS_OK) rc = ws_get(T, WS_TMP, W * 8, &TMP);
    if (rc == CKKS_OK) rc = ws_get(T, WS_SCR, 4 * cs_max * (size_t)A.K * L * n * 4, &SCR);
    if (rc == CKKS_OK && rescale) rc = ws_get(T, WS_LAST, 2 * cs_max * n * 8, &LAST);
    const size_t outL = rescale ? L - 1 : L;
    for (size_t s0 = 0; s0 < batch && rc == CKKS_OK; s0 += cs_max) {
        const size_t cs = batch - s0 < cs_max ? batch - s0 : cs_max;
        const size_t off = s0 * L * n;
        Span sp = whole(cs, L);
        auto step = [&]() -> int {
            const u64 *in[4] = {a0 + off, a1 + off, b0 + off, b1 + off};
            u64 *nt[4] = {A0, A1, B0, B1};
            for (int t = 0; t < 4; ++t) {
                TRY(run_pass(T, P_FWD1, sp, in[t], TMP));
                TRY(run_pass(T, P_FWD2, sp, TMP, nt[t]));
            }
            EwArgs e = ew_args(T, L, cs);
            KL("tensor", (tensor_kernel<<<ew_grid(e.total), 256, 0, S(T)>>>(e, A0, A1, B0, B1, A0, A1, B0)));  // d0,d1,d2
            // all three back to the coefficient domain: d2 is the digit polynomial (engine.rs:493), d0 / d1 are added
            // to the key-switch sums there instead of in the NTT domain (the same ring elements)
            TRY(run_pass(T, P_INV2, sp, B0, TMP));
            TRY(run_pass(T, P_INV1, sp, TMP, B1));
            TRY(run_pass(T, P_INV2, sp, A0, TMP));
            TRY(run_pass(T, P_INV1, sp, TMP, A0));
            TRY(run_pass(T, P_INV2, sp, A1, TMP));
            TRY(run_pass(T, P_INV1, sp, TMP, A1));
            u64 *d0 = o0 + s0 * outL * n, *d1 = o1 + s0 * outL * n;
            return aux_keyswitch(T, A, L, cs, B1, rlk, A0, A1, reinterpret_cast<u32 *>(SCR), d0, d1, rescale ? LAST : nullptr);
        };
        rc = step();
    }
    return rc;
}

// rotate_ciphertext (engine.rs:412-463) with the auxiliary-basis key-switch; same contract as fused_rotate: the digits are
// automorphism(c1), automorphism(c0) is the addend of the first sum, the second sum is the new c1 as it is.
static int fused_rotate_aux(const Tables &T, size_t L, size_t batch, const u64 *c0, const u64 *c1, u64 e, const ckks_ksk *key, u64 *o0,
                            u64 *o1) {
    AuxKs *Ap;
    TRY(aux_get(T, &Ap));
    const AuxKs &A = *Ap;
    if (key->aux_k != A.K) return CKKS_BAD_HANDLE;
    NvtxScope nvtx_call("ckks:rotate");
    const size_t n = T.n, cs_max = aux_chunk(T, A, L, batch);
    const u64 einv = inv_mod_pow2(e, 2 * n);
    u64 *D = nullptr, *R0 = nullptr, *SCR = nullptr;
    std::lock_guard<std::mutex> ws_lock(const_cast<Tables &>(T).ws_mu);
    int rc = ws_get(T, WS_B0, cs_max * L * n * 8, &D);
    if (rc == CKKS_OK) rc = ws_get(T, WS_A0, cs_max * L * n * 8, &R0);
    if (rc == CKKS_OK) rc = ws_get(T, WS_SCR, 4 * cs_max * (size_t)A.K * L * n * 4, &SCR);
    for (size_t s0 = 0; s0 < batch && rc == CKKS_OK; s0 += cs_max) {
        const size_t cs = batch - s0 < cs_max ? batch - s0 : cs_max;
        const size_t off = s0 * L * n;
        auto step = [&]() -> int {
            EwArgs ea = ew_args(T, L, cs);
            KL("automorphism", (automorphism_kernel<<<ew_grid(ea.total), 256, 0, S(T)>>>(ea, c1 + off, D, e, einv)));
            KL("automorphism", (automorphism_kernel<<<ew_grid(ea.total), 256, 0, S(T)>>>(ea, c0 + off, R0, e, einv)));
            return aux_keyswitch(T, A, L, cs, D, key, R0, nullptr, reinterpret_cast<u32 *>(SCR), o0 + off, o1 + off, nullptr);
        };
        rc = step();
    }
    return rc;
}
