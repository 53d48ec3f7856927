// limb_shard.inl -- optional limb-sharded mode of the gadget ct-mult (north_star: "an optional
// limb-sharded mode at N=2^16 / L>=24, where key-switching all-gathers the decomposed digits").
// Included at the end of ckks_b200.cu (uses its launch helpers).
//
// Ownership: limb j of the basis lives on GPU (j mod G); the local limb jl of rank r is limb r + G*jl, so
// the share stays balanced while rescale drops limbs from the end.  Every GPU holds, for every ciphertext
// of the batch, only its own limbs, and of the gadget key only the slices [digit i][own limb j] (1/G).
//
// Per chunk of ciphertexts (engine.rs:473-539 + :263-282, same minimal exact schedule as fused_mul_relin):
//   phase A (limb-local)  4 forward transforms per own limb, tensor, inverse transform of d2 whose LAST PASS
//                         STORES EVERY FINISHED DIGIT LIMB INTO THE GATHER BUFFER OF ALL G GPUs (plain
//                         stores over NVLink into peer HBM: the all-gather is fused into the producing kernel)
//   barrier               flag words in peer memory (lshard_barrier_kernel), no host involvement
//   phase B               ks_pass1/ks_pass2 over all L digits for the own target limbs, first inverse pass;
//                         the owner of the last limb finishes it and stores it into every GPU's `last` buffer
//                         (the broadcast rescale needs, fused the same way)
//   barrier
//   phase C               last inverse pass fused with rescale_into (poly.rs:214-225) for the own limbs
// The one-call entry point software-pipelines the chunks over two streams and two buffer sets: phase A of
// chunk k+1 (NVLink-bound in its last kernel) runs on an auxiliary stream while phases B/C of chunk k
// (integer-pipe bound) run on the context's stream, so the exchange hides behind the key-switch.
// With `peer_stores = 0` the stores go to the own buffers only and the caller runs the collectives itself
// (torch.distributed / NCCL all-gather and broadcast on the exported buffers) between the phases.
//
// Results are the reference's words: the schedule is the batch-sharded one restricted to the own limbs.

struct LsSet {  // one of the two buffer sets the chunk pipeline alternates between
    u64 *A0 = nullptr, *A1 = nullptr, *B0 = nullptr, *B1 = nullptr, *TMP = nullptr;
    size_t off_gather = 0, off_last = 0;  // byte offsets in the symmetric block
};
struct LsSym {  // symmetric buffers of one rank: one cudaMalloc block (one IPC handle), shared by all levels
    int device = 0, rank = 0, world = 1;
    size_t n = 0, cs_max = 0, Lg0 = 0, Ll0 = 0;
    unsigned char *block = nullptr;  // gather[0] | gather[1] | last[0] | last[1] | flags
    size_t off_flags = 0, block_bytes = 0;
    unsigned char *base_p[8] = {nullptr};  // block of every rank (own included)
    void *ipc_base[8] = {nullptr};
    bool connected = false;
    unsigned epochA = 0, epochB = 0;  // flag words [0..8) / [16..24): barriers on the auxiliary / main stream
    cudaEvent_t ev = nullptr;         // in-process groups: barrier by events (ckks_lshard_barrier_local)
    cudaStream_t aux = nullptr, aux2 = nullptr;  // aux2: the dropped limb's broadcast, under the rest of phase B
    cudaEvent_t ev_start = nullptr, ev_a[2] = {nullptr, nullptr}, ev_c[2] = {nullptr, nullptr}, ev_k = nullptr, ev_l = nullptr;
    unsigned long long timeout_ns = 20ull * 1000000000ull;
    unsigned *h_fail = nullptr, *d_fail = nullptr;  // mapped pinned word: a barrier's timeout, visible to the host at once
    // how the digits reach the peers in the pipelined entry point: 0 = stores from the producing kernel,
    // 1 = copy engines (cudaMemcpyAsync over the peer mappings: no SM is held while NVLink is busy, so the
    // exchange of chunk k+1 really runs under the key-switch of chunk k)
    int exchange = 0;
    LsSet set[2];
    u64 *SCR = nullptr;
    u64 *gather(int k = 0) const { return reinterpret_cast<u64 *>(block + set[k].off_gather); }
    u64 *last(int k = 0) const { return reinterpret_cast<u64 *>(block + set[k].off_last); }
    u64 *gather_of(int p, int k) const { return reinterpret_cast<u64 *>(base_p[p] + set[k].off_gather); }
    u64 *last_of(int p, int k) const { return reinterpret_cast<u64 *>(base_p[p] + set[k].off_last); }
    unsigned *flags() const { return reinterpret_cast<unsigned *>(block + off_flags); }  // [8] = error word
    unsigned *flags_of(int p) const { return reinterpret_cast<unsigned *>(base_p[p] + off_flags); }
    ~LsSym() {
        cudaSetDevice(device);
        if (ev) cudaEventDestroy(ev);
        for (cudaEvent_t e : {ev_start, ev_a[0], ev_a[1], ev_c[0], ev_c[1], ev_k, ev_l})
            if (e) cudaEventDestroy(e);
        if (aux) cudaStreamDestroy(aux);
        if (aux2) cudaStreamDestroy(aux2);
        if (h_fail) cudaFreeHost(h_fail);
        for (int p = 0; p < 8; ++p)
            if (ipc_base[p]) cudaIpcCloseMemHandle(ipc_base[p]);
        if (block) cudaFree(block);
        for (int k = 0; k < 2; ++k)
            for (u64 *b : {set[k].A0, set[k].A1, set[k].B0, set[k].B1, set[k].TMP})
                if (b) cudaFree(b);
        if (SCR) cudaFree(SCR);
    }
};
struct ckks_lshard {
    uint32_t magic;
    int rank, world;
    size_t Lg;  // limbs of the whole basis at this level
    size_t Ll;  // limbs held here
    std::vector<u64> moduli_g;  // whole basis at the top level
    ckks_ctx *local;            // context over the own limbs (local limb jl = basis limb rank + world*jl)
    std::shared_ptr<LsSym> sym;
    void *d_qlinv_last;  // [Ll] (q_{Lg-1})^-1 mod own q (Shoup pairs in the transform word type)
    bool digit_reduce;
};
static bool ok_lshard(const ckks_lshard *s) { return s && s->magic == MAGIC_LSHARD && ok_ctx(s->local) && s->sym; }
static size_t ls_count(size_t Lg, int rank, int world) { return Lg > (size_t)rank ? (Lg - rank + world - 1) / world : 0; }

// Level-specific tables: q_last^-1 mod every own modulus, and whether foreign digits fit the lazy input range.
static int ls_level_tables(ckks_lshard *s) {
    const Tables &T = *s->local->T;
    const u64 qlast = s->moduli_g[s->Lg - 1];
    u64 qmax = 0, qmin = ~0ull;
    for (size_t i = 0; i < s->Lg; ++i) qmax = s->moduli_g[i] > qmax ? s->moduli_g[i] : qmax;
    for (size_t jl = 0; jl < s->Ll; ++jl) qmin = T.moduli[jl] < qmin ? T.moduli[jl] : qmin;
    // a digit x < q_i enters the lazy transform mod q_j unreduced iff x < 4 q_j (and fits the word type)
    s->digit_reduce = !(T.lazy && (qmax >> 2) < qmin && (!T.w32 || (qmax >> 31) == 0));
    s->d_qlinv_last = nullptr;
    if (s->Ll == 0) return CKKS_OK;
    if (T.w32) {
        std::vector<tw32_t> h(s->Ll);
        for (size_t jl = 0; jl < s->Ll; ++jl) {
            u64 q = T.moduli[jl];
            h[jl] = (q == qlast) ? ht::mk_tw32(0, q) : ht::mk_tw32(hm::inv_mod(qlast % q, q), q);
        }
        TRY(upload_vec((tw32_t **)&s->d_qlinv_last, h));
    } else {
        std::vector<tw_t> h(s->Ll);
        for (size_t jl = 0; jl < s->Ll; ++jl) {
            u64 q = T.moduli[jl];
            h[jl] = (q == qlast) ? mk_tw(0, q) : mk_tw(hm::inv_mod(qlast % q, q), q);
        }
        TRY(upload_vec((tw_t **)&s->d_qlinv_last, h));
    }
    return CKKS_OK;
}

extern "C" int ckks_lshard_create(uint64_t n, const uint64_t *moduli, size_t l, int rank, int world, int device, size_t chunk,
                                  ckks_lshard **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    if (l == 0) return CKKS_EMPTY_BASIS;
    if (!moduli || world < 1 || world > 8 || rank < 0 || rank >= world) return CKKS_BAD_ARGUMENT;
    if (l < (size_t)world) {
        g_err = "limb-sharded mode needs at least one limb per GPU";
        return CKKS_UNSUPPORTED;
    }
    std::vector<u64> own;
    for (size_t j = rank; j < l; j += world) own.push_back(moduli[j]);
    // validate the WHOLE basis like RnsBasis::new before building the own share
    if (n == 0 || (n & (n - 1)) != 0) return CKKS_INVALID_DEGREE;
    for (size_t i = 0; i < l; ++i)
        if (!hm::is_ntt_friendly_prime(moduli[i], n)) return CKKS_NON_NTT_FRIENDLY_MODULUS;
    ckks_ctx *local = nullptr;
    TRY(ckks_ctx_create(n, reinterpret_cast<const uint64_t *>(own.data()), own.size(), device, &local));
    if (local->T->path != 2) {
        ckks_ctx_destroy(local);
        g_err = "limb-sharded mode is built on the four-step path (N >= 256)";
        return CKKS_UNSUPPORTED;
    }
    auto sym = std::make_shared<LsSym>();
    sym->device = device;
    sym->rank = rank;
    sym->world = world;
    sym->n = n;
    sym->Lg0 = l;
    sym->Ll0 = own.size();
    size_t per = ls_count(l, 0, world) * l * n * sizeof(u64);  // scratch per ciphertext on the fullest rank
    size_t cs = ((size_t)4 << 30) / per;
    if (cs < 1) cs = 1;
    if (cs > 1024) cs = 1024;  // small rings: bound the exchange buffers rather than fill 4 GiB
    if (chunk && chunk < cs) cs = chunk;
    sym->cs_max = cs;
    const size_t gather_bytes = l * cs * n * 8, last_bytes = 2 * cs * n * 8;
    sym->set[0].off_gather = 0;
    sym->set[1].off_gather = gather_bytes;
    sym->set[0].off_last = 2 * gather_bytes;
    sym->set[1].off_last = 2 * gather_bytes + last_bytes;
    sym->off_flags = 2 * gather_bytes + 2 * last_bytes;
    sym->block_bytes = sym->off_flags + 256;
    ckks_lshard *s = new ckks_lshard();
    s->magic = MAGIC_LSHARD;
    s->rank = rank;
    s->world = world;
    s->Lg = l;
    s->Ll = own.size();
    s->moduli_g.assign(moduli, moduli + l);
    s->local = local;
    s->sym = sym;
    s->d_qlinv_last = nullptr;
    auto body = [&]() -> int {
        CU(cudaSetDevice(device));
        CU(cudaMalloc((void **)&sym->block, sym->block_bytes));  // plain cudaMalloc: exportable with cudaIpcGetMemHandle
        CU(cudaMemset(sym->block + sym->off_flags, 0, 256));
        CU(cudaHostAlloc((void **)&sym->h_fail, sizeof(unsigned), cudaHostAllocMapped));
        *sym->h_fail = 0;
        CU(cudaHostGetDevicePointer((void **)&sym->d_fail, sym->h_fail, 0));
        local->T->fail_word = sym->h_fail;
        const size_t W = cs * sym->Ll0 * n * 8;
        for (int k = 0; k < 2; ++k)
            for (u64 **b : {&sym->set[k].A0, &sym->set[k].A1, &sym->set[k].B0, &sym->set[k].B1, &sym->set[k].TMP}) CU(cudaMalloc((void **)b, W));
        CU(cudaMalloc((void **)&sym->SCR, W * l));
        CU(cudaStreamCreateWithFlags(&sym->aux, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&sym->aux2, cudaStreamNonBlocking));
        for (cudaEvent_t *e : {&sym->ev_start, &sym->ev_a[0], &sym->ev_a[1], &sym->ev_c[0], &sym->ev_c[1], &sym->ev_k, &sym->ev_l})
            CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        CU(cudaDeviceSynchronize());
        sym->base_p[rank] = sym->block;
        if (world == 1) sym->connected = true;
        return ls_level_tables(s);
    };
    int rc = body();
    if (rc != CKKS_OK) {
        s->magic = 0;
        delete s;
        ckks_ctx_destroy(local);
        return rc;
    }
    *out = s;
    return CKKS_OK;
}
extern "C" int ckks_lshard_destroy(ckks_lshard *s) {
    if (!ok_lshard(s)) return CKKS_BAD_HANDLE;
    cudaSetDevice(s->sym->device);
    cudaStreamSynchronize(s->local->T->stream);
    if (s->d_qlinv_last) cudaFree(s->d_qlinv_last);
    ckks_ctx_destroy(s->local);
    s->magic = 0;
    delete s;
    return CKKS_OK;
}
// The level below (rescale_ciphertext drops the last limb of the basis, engine.rs:263-282): its owner loses
// one local limb, everyone else keeps theirs; buffers are shared with the parent.
extern "C" int ckks_lshard_drop_last(ckks_lshard *s, ckks_lshard **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    if (!ok_lshard(s)) return CKKS_BAD_HANDLE;
    if (s->Lg < 2) return CKKS_INVALID_MOD_DROP;
    if (s->Lg - 1 < (size_t)s->world) {
        g_err = "limb-sharded mode needs at least one limb per GPU";
        return CKKS_UNSUPPORTED;
    }
    const bool owner = (int)((s->Lg - 1) % s->world) == s->rank;
    ckks_ctx *local = nullptr;
    TRY(ckks_ctx_drop_last(s->local, owner ? 1 : 0, &local));
    ckks_lshard *c = new ckks_lshard();
    c->magic = MAGIC_LSHARD;
    c->rank = s->rank;
    c->world = s->world;
    c->Lg = s->Lg - 1;
    c->Ll = local->L;
    c->moduli_g = s->moduli_g;
    c->local = local;
    c->sym = s->sym;
    c->d_qlinv_last = nullptr;
    CU(cudaSetDevice(s->sym->device));
    int rc = ls_level_tables(c);
    if (rc != CKKS_OK) {
        ckks_lshard_destroy(c);
        return rc;
    }
    *out = c;
    return CKKS_OK;
}
extern "C" ckks_ctx *ckks_lshard_local_ctx(ckks_lshard *s) { return ok_lshard(s) ? s->local : nullptr; }
extern "C" size_t ckks_lshard_channel_count(const ckks_lshard *s) { return ok_lshard(s) ? s->Lg : 0; }
extern "C" size_t ckks_lshard_chunk(const ckks_lshard *s) { return ok_lshard(s) ? s->sym->cs_max : 0; }
extern "C" int ckks_lshard_set_exchange(ckks_lshard *s, int mode) {
    if (!ok_lshard(s)) return CKKS_BAD_HANDLE;
    if (mode != 0 && mode != 1) return CKKS_BAD_ARGUMENT;
    s->sym->exchange = mode;
    return CKKS_OK;
}
extern "C" int ckks_lshard_set_timeout_ms(ckks_lshard *s, uint64_t ms) {
    if (!ok_lshard(s)) return CKKS_BAD_HANDLE;
    s->sym->timeout_ns = ms * 1000000ull;
    return CKKS_OK;
}
// Device pointers of the exchange buffers, for a caller that runs the collectives itself (peer_stores = 0):
// gather [Lg0][chunk][N] (digit limb i of chunk ciphertext c at (i*chunk + c)*N), last [2][chunk][N].
extern "C" int ckks_lshard_buffers(ckks_lshard *s, uint64_t **gather, size_t *gather_words, uint64_t **last, size_t *last_words) {
    if (!ok_lshard(s)) return CKKS_BAD_HANDLE;
    const LsSym &y = *s->sym;
    if (gather) *gather = reinterpret_cast<uint64_t *>(y.gather(0));
    if (gather_words) *gather_words = y.Lg0 * y.cs_max * y.n;
    if (last) *last = reinterpret_cast<uint64_t *>(y.last(0));
    if (last_words) *last_words = 2 * y.cs_max * y.n;
    return CKKS_OK;
}

// ---- wiring the GPUs together ------------------------------------------------------------------------
extern "C" size_t ckks_lshard_ipc_size(void) { return sizeof(cudaIpcMemHandle_t); }
extern "C" int ckks_lshard_ipc_export(ckks_lshard *s, void *blob) {
    if (!ok_lshard(s) || !blob) return CKKS_BAD_HANDLE;
    CU(cudaSetDevice(s->sym->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s->sym->block));
    memcpy(blob, &h, sizeof(h));
    return CKKS_OK;
}
// blobs: world handles in rank order (one process per GPU; the caller moves them, e.g. all_gather_object).
extern "C" int ckks_lshard_ipc_import(ckks_lshard *s, const void *blobs) {
    if (!ok_lshard(s) || !blobs) return CKKS_BAD_HANDLE;
    LsSym &y = *s->sym;
    CU(cudaSetDevice(y.device));
    for (int p = 0; p < y.world; ++p) {
        if (p == y.rank || y.ipc_base[p]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const unsigned char *)blobs + (size_t)p * sizeof(h), sizeof(h));
        void *base = nullptr;
        CU(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        y.ipc_base[p] = base;
        y.base_p[p] = reinterpret_cast<unsigned char *>(base);
    }
    y.connected = true;
    return CKKS_OK;
}
// All ranks in ONE process (several GPUs, or several ranks on one GPU for tests): direct pointers.
extern "C" int ckks_lshard_connect_local(ckks_lshard **shards, int world) {
    if (!shards || world < 1 || world > 8) return CKKS_BAD_ARGUMENT;
    for (int r = 0; r < world; ++r)
        if (!ok_lshard(shards[r]) || shards[r]->world != world || shards[r]->rank != r) return CKKS_BAD_HANDLE;
    for (int r = 0; r < world; ++r) {
        LsSym &y = *shards[r]->sym;
        if (y.block_bytes != shards[0]->sym->block_bytes) return CKKS_BATCH_MISMATCH;  // different chunk / basis
        CU(cudaSetDevice(y.device));
        for (int p = 0; p < world; ++p) {
            LsSym &z = *shards[p]->sym;
            if (z.device != y.device) {
                int can = 0;
                CU(cudaDeviceCanAccessPeer(&can, y.device, z.device));
                if (!can) {
                    g_err = "no peer access between the GPUs of a limb-sharded group";
                    return CKKS_CUDA_ERROR;
                }
                cudaError_t e = cudaDeviceEnablePeerAccess(z.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
                cudaGetLastError();
            }
            y.base_p[p] = z.block;
        }
        y.connected = true;
    }
    return CKKS_OK;
}

// ---- key slices --------------------------------------------------------------------------------------
// a, b: host, [digit i = 0..Lg)[own limb jl][N], coefficient domain (the rows of the reference's
// RnsGadgetRelinKey, engine.rs:225-253, restricted to this GPU's limbs); transformed once.
extern "C" int ckks_lshard_ksk_upload(ckks_lshard *s, const uint64_t *a, const uint64_t *b, ckks_ksk **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    if (!ok_lshard(s)) return CKKS_BAD_HANDLE;
    if (!a || !b) return CKKS_BAD_ARGUMENT;
    const Tables &T = *s->local->T;
    CU(cudaSetDevice(T.device));
    ckks_ksk *k;
    TRY(ksk_new(s->local, &k, s->Lg));
    const size_t words = s->Lg * s->Ll * T.n;
    int rc = CKKS_OK;
    if (cudaMemcpyAsync(k->a, a, words * 8, cudaMemcpyHostToDevice, S(T)) != cudaSuccess ||
        cudaMemcpyAsync(k->b, b, words * 8, cudaMemcpyHostToDevice, S(T)) != cudaSuccess)
        rc = cuda_fail(cudaGetLastError(), "ksk h2d");
    if (rc == CKKS_OK) rc = scan_reduced_pair_sync(T, s->Ll, s->Lg, k->a, k->b);  // canonical words only (poly.rs:83-93)
    if (rc == CKKS_OK) rc = ntt_inplace(T, s->Ll, s->Lg, k->a, false);
    if (rc == CKKS_OK) rc = ntt_inplace(T, s->Ll, s->Lg, k->b, false);
    if (rc == CKKS_OK) rc = ksk_finalize(T, k);
    if (rc == CKKS_OK && cudaStreamSynchronize(S(T)) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "ksk sync");
    if (rc != CKKS_OK) {
        ckks_ksk_free(k);
        return rc;
    }
    *out = k;
    return CKKS_OK;
}

// ---- the three phases --------------------------------------------------------------------------------
// Last inverse pass (negacyclic GS over rho, canonical coefficient-domain words) with the result stored
// into `npeer` buffers at slot (m_off + m_step*limb), layout [slot][m_cs][N].
template <typename WD, int A>
static int launch_inv1_multi_w(int lazy, dim3 grid, cudaStream_t s, const PassArgs &a) {
    constexpr int E = 4, C = 16;  // 16 columns: 128-byte segments for the stores that cross NVLink
    grid.x = a.ncols / C;
    const size_t smem = (size_t)(1 << A) * (C + 1) * sizeof(WD);
    const int block = C << (A - E);
#define M(LZ) \
    KL("ntt_inv_pass1_allgather", (ntt_pass_kernel<WD, XF_NEG_INV, A, E, C, (sizeof(WD) == 8 ? LZ : (LZ ? 1 : 0)), false, false, false, true><<<grid, block, smem, s>>>(a)))
    LZ_SWITCH(WD, lazy, M);
#undef M
    return CKKS_OK;
}
template <int A>
static int launch_inv1_multi_a(bool w32, int lazy, dim3 grid, cudaStream_t s, const PassArgs &a) {
    if (w32) return W32_DISPATCH(launch_inv1_multi_w<u32, A>(lazy, grid, s, a));
    return launch_inv1_multi_w<u64, A>(lazy, grid, s, a);
}
static int ls_inv1_multi(const Tables &T, size_t cs, int L, int limb0, int nl, const void *src, u64 *const *peers, int npeer, int first,
                         int m_off, int m_step, size_t m_cs) {
    if (!cs || !nl) return CKKS_OK;
    PassArgs a;
    memset(&a, 0, sizeof(a));
    a.lc = T.d_lc;
    a.L = L;
    a.N = T.n;
    a.limb0 = limb0;
    a.src = src;
    a.tab = T.d_P1i;
    a.tab_stride = (size_t)1 << T.a1;
    a.ncols = 1u << T.a2;
    a.npeer = npeer;
    a.m_first = npeer > 1 ? first % npeer : 0;
    for (int p = 0; p < npeer; ++p) a.peer[p] = peers[p];
    a.m_off = m_off;
    a.m_step = m_step;
    a.m_cs = m_cs;
    dim3 g(1, (unsigned)nl, (unsigned)cs);
    DISPATCH_A(T.a1, TRY(launch_inv1_multi_a<AA>(T.w32, T.lazy, g, S(T), a)));
    return CKKS_OK;
}
// A barrier that timed out leaves this rank's epochs out of step with its peers for good and its buffers half
// filled: the group is dead.  Every entry point refuses further work (sticky CKKS_NCCL_ERROR) until all shards of the
// group are destroyed and re-created (ckks_lshard_create + connect): that rebuilds buffers, flags and epochs.
static int ls_failed(const ckks_lshard *s) {
    const LsSym &y = *s->sym;
    if (y.h_fail && *reinterpret_cast<const volatile unsigned *>(y.h_fail)) {
        g_err = "limb-sharded barrier " + std::to_string(*y.h_fail) +
                " timed out waiting for a peer GPU: the group is unusable, destroy and re-create every shard of it";
        return CKKS_NCCL_ERROR;
    }
    return CKKS_OK;
}
// Enqueued behind the last kernel of a call: all-ones outputs if a barrier of this call (or an earlier one) failed.
static int ls_poison_guard(ckks_lshard *s, ckks_poly *o0, ckks_poly *o1, cudaStream_t st) {
    LsSym &y = *s->sym;
    if (y.world == 1) return CKKS_OK;
    g_cur_stream = st;
    KL("lshard_poison_guard", (lshard_poison_kernel<<<148, 256, 0, st>>>(y.flags() + 8, o0 ? o0->d : nullptr, o0 ? poly_words(o0) : 0, o1 ? o1->d : nullptr,
                                                                         o1 ? poly_words(o1) : 0)));
    return CKKS_OK;
}
// which = 0: flag words [0,8) (auxiliary-stream sequence of the chunk pipeline); 1: words [16,24) (main stream).
// Each sequence is issued in the same order on every rank, so one monotonic epoch per sequence suffices.
static int ls_barrier(ckks_lshard *s, int which, cudaStream_t st) {
    LsSym &y = *s->sym;
    if (y.world == 1) return CKKS_OK;
    BarArgs b;
    memset(&b, 0, sizeof(b));
    const int fo = which ? 16 : 0;
    for (int p = 0; p < y.world; ++p) b.peer_flags[p] = y.flags_of(p) + fo;
    b.my_flags = y.flags() + fo;
    b.err = y.flags() + 8;
    b.host_err = y.d_fail;
    b.rank = y.rank;
    b.world = y.world;
    b.epoch = which ? ++y.epochB : ++y.epochA;
    b.timeout_ns = y.timeout_ns;
    g_cur_stream = st;
    KL("lshard_barrier", (lshard_barrier_kernel<<<1, 32, 0, st>>>(b)));
    return CKKS_OK;
}
// Synchronise the stream and report a barrier that timed out (a peer that never arrived).
extern "C" int ckks_lshard_check(ckks_lshard *s) {
    if (!ok_lshard(s)) return CKKS_BAD_HANDLE;
    LsSym &y = *s->sym;
    CU(cudaSetDevice(y.device));
    unsigned e = 0;
    CU(cudaMemcpyAsync(&e, y.flags() + 8, sizeof(e), cudaMemcpyDeviceToHost, s->local->T->stream));
    CU(cudaStreamSynchronize(s->local->T->stream));
    if (e && y.h_fail) *y.h_fail = e;  // (the kernel wrote it already; kept for a device without mapped-memory coherence)
    return ls_failed(s);  // sticky: see ls_failed
}

static int ls_check_inputs(ckks_lshard *s, const ckks_poly *a0, const ckks_poly *a1, const ckks_poly *b0, const ckks_poly *b1,
                           const ckks_ksk *rlk, ckks_lshard *child, const ckks_poly *o0, const ckks_poly *o1) {
    if (!ok_lshard(s)) return CKKS_BAD_HANDLE;
    TRY(check_ct(a0, a1));
    TRY(check_ct(b0, b1));
    TRY(check_pair(a0, b0, false));
    if (!same_basis(a0->ctx, s->local)) return CKKS_BASIS_MISMATCH;
    if (a0->ntt) return CKKS_DOMAIN_MISMATCH;
    if (!ok_ksk_slice(rlk) || rlk->digits != s->Lg) return CKKS_BAD_HANDLE;
    if (!same_basis(rlk->ctx, s->local)) return CKKS_BASIS_MISMATCH;
    if (child) {
        if (!ok_lshard(child) || child->sym.get() != s->sym.get() || child->Lg + 1 != s->Lg) return CKKS_BASIS_MISMATCH;
    }
    TRY(check_ct(o0, o1));
    if (!same_basis(o0->ctx, child ? child->local : s->local)) return CKKS_BASIS_MISMATCH;
    if (o0->batch != a0->batch) return CKKS_BATCH_MISMATCH;
    return CKKS_OK;
}

// One phase (0 = A, 1 = B, 2 = C) of mul_ciphertexts_gadget [+ rescale_ciphertext when `child`] for the
// ciphertexts [s0, s0 + cs) of the batch on buffer set `k`, enqueued on the context's current stream.
static int ls_mul_phase(ckks_lshard *s, int k, int phase, size_t s0, size_t cs, const ckks_poly *a0, const ckks_poly *a1,
                        const ckks_poly *b0, const ckks_poly *b1, const ckks_ksk *rlk, ckks_lshard *child, ckks_poly *o0, ckks_poly *o1,
                        int peer_stores) {
    LsSym &y = *s->sym;
    const LsSet &w = y.set[k];
    if (cs == 0) return CKKS_OK;
    if (cs > y.cs_max || s0 + cs > a0->batch) return CKKS_BAD_ARGUMENT;
    if (peer_stores && !y.connected) {
        g_err = "limb-sharded group is not connected (ckks_lshard_ipc_import / ckks_lshard_connect_local)";
        return CKKS_BAD_ARGUMENT;
    }
    const Tables &T = *s->local->T;
    CU(cudaSetDevice(T.device));
    g_cur_stream = S(T);
    const size_t n = T.n, Ll = s->Ll, Lg = s->Lg;
    const size_t off = s0 * Ll * n;
    const bool rescale = child != nullptr;
    const int owner = (int)((Lg - 1) % s->world);
    const int np = peer_stores ? y.world : 1;
    u64 *pg[8], *pl0[8], *pl1[8];  // where the digits / the two components of the dropped limb go
    for (int p = 0; p < np; ++p) {
        pg[p] = peer_stores ? y.gather_of(p, k) : y.gather(k);
        pl0[p] = peer_stores ? y.last_of(p, k) : y.last(k);
        pl1[p] = pl0[p] + y.cs_max * n;
    }
    Span sp = whole(cs, Ll);
    if (phase == 0) {
        const u64 *in[4] = {a0->d + off, a1->d + off, b0->d + off, b1->d + off};
        u64 *nt[4] = {w.A0, w.A1, w.B0, w.B1};
        for (int t = 0; t < 4; ++t) {
            TRY(run_pass(T, P_FWD1, sp, in[t], w.TMP));
            TRY(run_pass(T, P_FWD2, sp, w.TMP, nt[t]));
        }
        EwArgs e = ew_args(T, Ll, cs);
        KL("tensor", (tensor_kernel<<<ew_grid(e.total), 256, 0, S(T)>>>(e, w.A0, w.A1, w.B0, w.B1, w.A0, w.A1, w.B0)));  // d0,d1,d2
        TRY(run_pass(T, P_INV2, sp, w.B0, w.TMP));
        // digits: coefficient-domain limbs of d2 (engine.rs:493,507), stored where every GPU will read them
        if (peer_stores && y.exchange == 1 && y.world > 1) {
            u64 *self[1] = {y.gather(k)};
            TRY(ls_inv1_multi(T, cs, (int)Ll, 0, (int)Ll, w.TMP, self, 1, 0, s->rank, s->world, y.cs_max));
            for (int t = 1; t < y.world; ++t) {
                const int p = (s->rank + t) % y.world;
                for (size_t jl = 0; jl < Ll; ++jl) {
                    const size_t slot = ((size_t)s->rank + (size_t)s->world * jl) * y.cs_max * n;
                    CU(cudaMemcpyAsync(y.gather_of(p, k) + slot, y.gather(k) + slot, cs * n * sizeof(u64), cudaMemcpyDefault, S(T)));
                }
            }
            return CKKS_OK;
        }
        TRY(ls_inv1_multi(T, cs, (int)Ll, 0, (int)Ll, w.TMP, pg, np, s->rank + 1, s->rank, s->world, y.cs_max));
        return CKKS_OK;
    }
    if (phase == 1) {
        KsShard ks{Lg, s->rank, s->world, n, y.cs_max * n, s->digit_reduce};
        if (rescale && s->rank == owner && peer_stores && y.world > 1 && Ll > 1) {
            // The owner of the limb rescale drops key-switches that limb FIRST and sends it to the peers (one GPU
            // feeding world-1 others: the longest transfer of the step) on a side stream while the key-switch of
            // its other limbs runs; the barrier that follows phase B finds the broadcast already done.
            const cudaStream_t cur = S(T);
            TRY(ks_fused_ex(T, Ll, ks, cs, y.gather(k), w.B0, rlk, w.A0, w.A1, y.SCR, w.TMP, w.B1, true, Ll - 1, 1));
            CU(cudaEventRecord(y.ev_k, cur));
            CU(cudaStreamWaitEvent(y.aux2, y.ev_k, 0));
            {
                StreamScope side(y.aux2);  // thread-local redirection of the launch helpers (no shared state is touched)
                TRY(ls_inv1_multi(T, cs, (int)Ll, (int)Ll - 1, 1, w.TMP, pl0, np, s->rank + 1, 0, 0, y.cs_max));
                TRY(ls_inv1_multi(T, cs, (int)Ll, (int)Ll - 1, 1, w.B1, pl1, np, s->rank + 1, 0, 0, y.cs_max));
            }
            CU(cudaEventRecord(y.ev_l, y.aux2));
            TRY(ks_fused_ex(T, Ll, ks, cs, y.gather(k), w.B0, rlk, w.A0, w.A1, y.SCR, w.TMP, w.B1, true, 0, Ll - 1));
            CU(cudaStreamWaitEvent(cur, y.ev_l, 0));
            return CKKS_OK;
        }
        TRY(ks_fused_ex(T, Ll, ks, cs, y.gather(k), w.B0, rlk, w.A0, w.A1, y.SCR, w.TMP, w.B1, true));
        if (!rescale) {
            const size_t ooff = s0 * Ll * n;
            TRY(run_pass(T, P_INV1, sp, w.TMP, o0->d + ooff));
            TRY(run_pass(T, P_INV1, sp, w.B1, o1->d + ooff));
            return CKKS_OK;
        }
        if (s->rank == owner) {  // the limb rescale drops: finish it and hand it to everyone
            TRY(ls_inv1_multi(T, cs, (int)Ll, (int)Ll - 1, 1, w.TMP, pl0, np, s->rank + 1, 0, 0, y.cs_max));
            TRY(ls_inv1_multi(T, cs, (int)Ll, (int)Ll - 1, 1, w.B1, pl1, np, s->rank + 1, 0, 0, y.cs_max));
        }
        return CKKS_OK;
    }
    if (phase == 2) {
        if (!rescale) return CKKS_OK;
        const size_t outL = child->Ll;
        if (outL == 0) return CKKS_OK;
        PassArgs pa;
        memset(&pa, 0, sizeof(pa));
        pa.lc = T.d_lc;
        pa.tab = T.d_P1i;
        pa.tab_stride = (size_t)1 << T.a1;
        pa.L = (int)Ll;
        pa.ncols = 1u << T.a2;
        pa.N = n;
        pa.dstL = (int)outL;
        dim3 g(1, (unsigned)outL, (unsigned)cs);
        const size_t ooff = s0 * outL * n;
        pa.src = w.TMP;
        pa.dst = o0->d + ooff;
        DISPATCH_A(T.a1, TRY(launch_inv1_rescale_a<AA>(T.w32, T.lazy, g, S(T), pa, y.last(k), s->d_qlinv_last)));
        pa.src = w.B1;
        pa.dst = o1->d + ooff;
        DISPATCH_A(T.a1, TRY(launch_inv1_rescale_a<AA>(T.w32, T.lazy, g, S(T), pa, y.last(k) + y.cs_max * n, s->d_qlinv_last)));
        return CKKS_OK;
    }
    return CKKS_BAD_ARGUMENT;
}
// The public phase call: buffer set 0, the context's stream; cs <= ckks_lshard_chunk().  o0/o1: polynomials of
// the child's (or this level's) local context with the same batch as the inputs.
extern "C" int ckks_lshard_mul_phase(ckks_lshard *s, int phase, size_t s0, size_t cs, const ckks_poly *a0, const ckks_poly *a1,
                                     const ckks_poly *b0, const ckks_poly *b1, const ckks_ksk *rlk, ckks_lshard *child, ckks_poly *o0,
                                     ckks_poly *o1, int peer_stores) {
    TRY(ls_check_inputs(s, a0, a1, b0, b1, rlk, child, o0, o1));
    TRY(ls_failed(s));
    TRY(ls_mul_phase(s, 0, phase, s0, cs, a0, a1, b0, b1, rlk, child, o0, o1, peer_stores));
    if (phase == 2 && peer_stores) TRY(ls_poison_guard(s, o0, o1, s->local->T->stream));
    return CKKS_OK;
}
extern "C" int ckks_lshard_barrier(ckks_lshard *s) {
    if (!ok_lshard(s)) return CKKS_BAD_HANDLE;
    if (!s->sym->connected) return CKKS_BAD_ARGUMENT;
    TRY(ls_failed(s));
    CU(cudaSetDevice(s->sym->device));
    return ls_barrier(s, 1, s->local->T->stream);
}

// ---- gadget key-switch of a coefficient-domain polynomial (rotate_ciphertext, engine.rs:429-452) ------
// phase 0: this GPU's limbs of `digits` (ciphertexts [s0, s0+cs)) go to every GPU's gather buffer;
// phase 1: ks0/ks1[s0..] = sum_i alpha_i * key_b[i] / key_a[i] on the own limbs, coefficient domain.
extern "C" int ckks_lshard_ks_phase(ckks_lshard *s, int phase, size_t s0, size_t cs, const ckks_poly *digits, const ckks_ksk *key,
                                    ckks_poly *ks0, ckks_poly *ks1, int peer_stores) {
    if (!ok_lshard(s)) return CKKS_BAD_HANDLE;
    if (!ok_poly(digits) || !ok_poly(ks0) || !ok_poly(ks1)) return CKKS_BAD_HANDLE;
    if (!same_basis(digits->ctx, s->local) || !same_basis(ks0->ctx, s->local) || !same_basis(ks1->ctx, s->local)) return CKKS_BASIS_MISMATCH;
    if (digits->ntt) return CKKS_DOMAIN_MISMATCH;
    if (!ok_ksk_slice(key) || key->digits != s->Lg || !same_basis(key->ctx, s->local)) return CKKS_BAD_HANDLE;
    if (ks0->batch != digits->batch || ks1->batch != digits->batch) return CKKS_BATCH_MISMATCH;
    TRY(ls_failed(s));
    LsSym &y = *s->sym;
    if (cs == 0) return CKKS_OK;
    if (cs > y.cs_max || s0 + cs > digits->batch) return CKKS_BAD_ARGUMENT;
    if (peer_stores && !y.connected) {
        g_err = "limb-sharded group is not connected (ckks_lshard_ipc_import / ckks_lshard_connect_local)";
        return CKKS_BAD_ARGUMENT;
    }
    const Tables &T = *s->local->T;
    CU(cudaSetDevice(T.device));
    g_cur_stream = S(T);
    const size_t n = T.n, Ll = s->Ll, off = s0 * Ll * n;
    if (phase == 0) {
        PushArgs a;
        memset(&a, 0, sizeof(a));
        a.src = digits->d + off;
        a.npeer = peer_stores ? y.world : 1;
        a.m_first = a.npeer > 1 ? (s->rank + 1) % a.npeer : 0;
        for (int p = 0; p < a.npeer; ++p) a.peer[p] = peer_stores ? y.gather_of(p, 0) : y.gather(0);
        a.m_off = s->rank;
        a.m_step = s->world;
        a.m_cs = y.cs_max;
        a.L = (int)Ll;
        a.logn = T.logn;
        a.total2 = cs * Ll * n / 2;
        KL("lshard_push", (lshard_push_kernel<<<ew_grid(a.total2), 256, 0, S(T)>>>(a)));
        return CKKS_OK;
    }
    if (phase == 1) {
        KsShard ks{s->Lg, s->rank, s->world, n, y.cs_max * n, s->digit_reduce};
        TRY(ks_fused_ex(T, Ll, ks, cs, y.gather(0), nullptr, key, nullptr, nullptr, y.SCR, y.set[0].TMP, y.set[0].B1, false));
        Span sp = whole(cs, Ll);
        TRY(run_pass(T, P_INV1, sp, y.set[0].TMP, ks0->d + off));
        TRY(run_pass(T, P_INV1, sp, y.set[0].B1, ks1->d + off));
        ks0->ntt = ks1->ntt = false;
        if (peer_stores) TRY(ls_poison_guard(s, ks0, ks1, S(T)));
        return CKKS_OK;
    }
    return CKKS_BAD_ARGUMENT;
}
// rotate_ciphertext (engine.rs:412-463) on this GPU's limbs: rotate_slots(k) of c0 and c1 is limb-local; the
// rotated c1 is the digit polynomial every GPU needs.  Same calling rules as ckks_lshard_ct_mul_relin_rescale.
extern "C" int ckks_lshard_ct_rotate(ckks_lshard *s, const ckks_poly *c0, const ckks_poly *c1, const ckks_ksk *rotk, int32_t k,
                                     ckks_poly **o0, ckks_poly **o1) {
    if (!o0 || !o1) return CKKS_BAD_ARGUMENT;
    *o0 = *o1 = nullptr;
    if (!ok_lshard(s)) return CKKS_BAD_HANDLE;
    TRY(check_ct(c0, c1));
    if (!same_basis(c0->ctx, s->local)) return CKKS_BASIS_MISMATCH;
    ckks_poly *r0 = nullptr, *r1 = nullptr, *k0 = nullptr, *k1 = nullptr;
    int rc = ckks_poly_rotate_slots(c0, k, &r0);  // engine.rs:417-419
    if (rc == CKKS_OK) rc = ckks_poly_rotate_slots(c1, k, &r1);
    if (rc == CKKS_OK) rc = ckks_poly_to_coeff_domain(r0);
    if (rc == CKKS_OK) rc = ckks_poly_to_coeff_domain(r1);
    if (rc == CKKS_OK) rc = poly_new(s->local, c0->batch, false, &k0);
    if (rc == CKKS_OK) rc = poly_new(s->local, c0->batch, false, &k1);
    const size_t cs_max = s->sym->cs_max;
    for (size_t s0 = 0; rc == CKKS_OK && s0 < c0->batch; s0 += cs_max) {
        const size_t cs = c0->batch - s0 < cs_max ? c0->batch - s0 : cs_max;
        rc = ckks_lshard_ks_phase(s, 0, s0, cs, r1, rotk, k0, k1, 1);
        if (rc == CKKS_OK) rc = ls_barrier(s, 1, s->local->T->stream);
        if (rc == CKKS_OK) rc = ckks_lshard_ks_phase(s, 1, s0, cs, r1, rotk, k0, k1, 1);
        if (rc == CKKS_OK) rc = ls_barrier(s, 1, s->local->T->stream);  // gather buffers are free again
    }
    if (rc == CKKS_OK) rc = ckks_poly_add_assign(r0, k0);  // engine.rs:454-455
    if (rc == CKKS_OK) rc = ls_poison_guard(s, r0, k1, s->local->T->stream);
    free2(r1, k0);
    if (rc != CKKS_OK) {
        free2(r0, k1);
        return rc;
    }
    *o0 = r0;
    *o1 = k1;
    return CKKS_OK;
}

// Barrier for a group whose ranks all live in this process and are driven in lockstep by one host thread
// (phase by phase): an event per stream, every stream waits for every other one.  Unlike the flag barrier
// no kernel spins, so ranks may even share one GPU (tests) without depending on how the hardware
// interleaves their streams.
extern "C" int ckks_lshard_barrier_local(ckks_lshard **shards, int world) {
    if (!shards || world < 1 || world > 8) return CKKS_BAD_ARGUMENT;
    for (int r = 0; r < world; ++r)
        if (!ok_lshard(shards[r]) || shards[r]->world != world || shards[r]->rank != r) return CKKS_BAD_HANDLE;
    for (int r = 0; r < world; ++r) {
        LsSym &y = *shards[r]->sym;
        CU(cudaSetDevice(y.device));
        if (!y.ev) CU(cudaEventCreateWithFlags(&y.ev, cudaEventDisableTiming));
        CU(cudaEventRecord(y.ev, shards[r]->local->T->stream));
    }
    for (int r = 0; r < world; ++r) {
        CU(cudaSetDevice(shards[r]->sym->device));
        for (int p = 0; p < world; ++p)
            if (p != r) CU(cudaStreamWaitEvent(shards[r]->local->T->stream, shards[p]->sym->ev, 0));
    }
    return CKKS_OK;
}

// Chunk pipeline of the one-call entry point.  Stream `aux`: phase A of chunk k+1 on buffer set (k+1)%2 and the
// barrier that says "every digit of that chunk is in every gather buffer".  Context stream: phases B, barrier
// ("the dropped limb is everywhere; this set's gather buffer may be overwritten"), C of chunk k.  Events order
// the two streams; the two barrier sequences use separate flag words.
//   set reuse: A(k+2) waits for C(k) locally; a peer's A(k+2) stores come after that peer passed the main-stream
//   barrier of chunk k, which this GPU signals only after its phase B(k) has read the gather buffer.
static int ls_mul_pipeline(ckks_lshard *s, const ckks_poly *a0, const ckks_poly *a1, const ckks_poly *b0, const ckks_poly *b1,
                           const ckks_ksk *rlk, ckks_lshard *child, ckks_poly *r0, ckks_poly *r1) {
    LsSym &y = *s->sym;
    const Tables &T = *s->local->T;
    CU(cudaSetDevice(T.device));
    const cudaStream_t main = S(T), aux = y.aux;
    const size_t batch = a0->batch, cs_max = y.cs_max;
    const size_t K = (batch + cs_max - 1) / cs_max;
    if (K == 0) return CKKS_OK;
    // (a smaller first chunk, to shorten the phase A nothing can hide, was measured no faster at 8 GPUs)
    auto span = [&](size_t k, size_t &s0, size_t &cs) {
        s0 = k * cs_max;
        cs = batch - s0 < cs_max ? batch - s0 : cs_max;
    };
    // phase A of chunk k on the auxiliary stream
    auto phase_a = [&](size_t k) -> int {
        size_t s0, cs;
        span(k, s0, cs);
        int rc;
        {
            StreamScope on_aux(aux);  // thread-local: the shared Tables keep their stream
            rc = ls_mul_phase(s, (int)(k & 1), 0, s0, cs, a0, a1, b0, b1, rlk, child, r0, r1, 1);
        }
        if (rc == CKKS_OK) rc = ls_barrier(s, 0, aux);
        if (rc != CKKS_OK) return rc;
        CU(cudaEventRecord(y.ev_a[k & 1], aux));
        return CKKS_OK;
    };
    CU(cudaEventRecord(y.ev_start, main));  // inputs and earlier work on the context's stream
    CU(cudaStreamWaitEvent(aux, y.ev_start, 0));
    TRY(phase_a(0));
    for (size_t k = 0; k < K; ++k) {
        if (k + 1 < K) {
            if (k >= 1) CU(cudaStreamWaitEvent(aux, y.ev_c[(k + 1) & 1], 0));  // C(k-1) released that buffer set
            TRY(phase_a(k + 1));
        }
        size_t s0, cs;
        span(k, s0, cs);
        CU(cudaStreamWaitEvent(main, y.ev_a[k & 1], 0));
        TRY(ls_mul_phase(s, (int)(k & 1), 1, s0, cs, a0, a1, b0, b1, rlk, child, r0, r1, 1));
        TRY(ls_barrier(s, 1, main));
        TRY(ls_mul_phase(s, (int)(k & 1), 2, s0, cs, a0, a1, b0, b1, rlk, child, r0, r1, 1));
        CU(cudaEventRecord(y.ev_c[k & 1], main));
    }
    g_cur_stream = main;
    return CKKS_OK;
}

// mul_ciphertexts_gadget (+ rescale_ciphertext into `child`'s level when child != NULL) on this GPU's limbs
// of a batch, all phases and both cross-GPU exchanges enqueued on the stream without host synchronisation.
// Every GPU of the group must make the same call on its own share.
extern "C" int ckks_lshard_ct_mul_relin_rescale(ckks_lshard *s, const ckks_poly *a0, const ckks_poly *a1, const ckks_poly *b0,
                                                const ckks_poly *b1, const ckks_ksk *rlk, ckks_lshard *child, ckks_poly **o0,
                                                ckks_poly **o1) {
    if (!o0 || !o1) return CKKS_BAD_ARGUMENT;
    *o0 = *o1 = nullptr;
    if (!ok_lshard(s) || (child && !ok_lshard(child))) return CKKS_BAD_HANDLE;
    if (!ok_poly(a0)) return CKKS_BAD_HANDLE;
    ckks_poly *r0 = nullptr, *r1 = nullptr;
    ckks_ctx *octx = child ? child->local : s->local;
    int rc = poly_new(octx, a0->batch, false, &r0);
    if (rc == CKKS_OK) rc = poly_new(octx, a0->batch, false, &r1);
    if (rc == CKKS_OK) rc = ls_check_inputs(s, a0, a1, b0, b1, rlk, child, r0, r1);
    if (rc == CKKS_OK) rc = ls_failed(s);
    if (rc == CKKS_OK) rc = ls_mul_pipeline(s, a0, a1, b0, b1, rlk, child, r0, r1);
    if (rc == CKKS_OK) rc = ls_poison_guard(s, r0, r1, s->local->T->stream);
    if (rc != CKKS_OK) {
        free2(r0, r1);
        return rc;
    }
    *o0 = r0;
    *o1 = r1;
    return CKKS_OK;
}
