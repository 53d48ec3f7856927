// batch_shard.inl -- the batch-sharded multi-GPU group of the C ABI (SURVEY.md 8b "multi-GPU", 8e default mode).
// Included at the end of ckks_b200.cu (uses host_pipeline and the context helpers).
//
// The reference is single-threaded and single-device: its callers hold a `Vec<Ciphertext>` and loop over it
// (examples/horner_chain.rs:211-278, engine.rs:473).  Independent ciphertexts never interact except through the
// read-only gadget key, so ONE process can spread a host batch over the GPUs of a box with no inter-GPU traffic:
//   * ckks_comm_init builds one context (tables, stream, staging pipeline) per device;
//   * a key is uploaded once per call site and replicated to every device (576 MiB each at cfg4);
//   * every *_host call cuts the batch into `ndev` contiguous shares and runs the three-stage H2D / compute / D2H
//     pipeline of the single-GPU entry point on each device from its own host thread.
// Results are the single-GPU words (same kernels, same schedule); only the share boundaries differ.
#include <thread>

struct ckks_comm {
    uint32_t magic;
    std::vector<ckks_ctx *> ctx;  // one per device, same basis
};
struct ckks_comm_ksk {
    uint32_t magic;
    ckks_comm *comm;
    std::vector<ckks_ksk *> key;  // replica per device
    size_t L;
};
enum : uint32_t { MAGIC_COMM = 0x434b434du, MAGIC_COMM_KSK = 0x434b434bu };

static bool ok_comm(const ckks_comm *c) {
    if (!(c && c->magic == MAGIC_COMM && !c->ctx.empty())) return false;
    for (ckks_ctx *x : c->ctx)
        if (!(x && x->magic == MAGIC_CTX)) return false;
    return true;
}
static bool ok_comm_ksk(const ckks_comm_ksk *k) { return k && k->magic == MAGIC_COMM_KSK && ok_comm(k->comm) && k->key.size() == k->comm->ctx.size(); }

extern "C" int ckks_comm_destroy(ckks_comm *c) {
    if (!(c && c->magic == MAGIC_COMM)) return CKKS_BAD_HANDLE;
    for (ckks_ctx *x : c->ctx)
        if (x) ckks_ctx_destroy(x);
    c->magic = 0;
    delete c;
    return CKKS_OK;
}
extern "C" int ckks_comm_init(int ndev, const int *devices, uint64_t n, const uint64_t *moduli, size_t l, ckks_comm **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    if (ndev < 1 || ndev > 64) return CKKS_BAD_ARGUMENT;
    ckks_comm *c = new ckks_comm();
    c->magic = MAGIC_COMM;
    for (int i = 0; i < ndev; ++i) {
        ckks_ctx *x = nullptr;
        int rc = ckks_ctx_create(n, moduli, l, devices ? devices[i] : i, &x);  // validates like RnsBasis::new (basis.rs:97-106)
        if (rc != CKKS_OK) {
            ckks_comm_destroy(c);
            return rc;
        }
        c->ctx.push_back(x);
    }
    *out = c;
    return CKKS_OK;
}
// RnsBasis::drop_last (basis.rs:121-134) for every device of the group.
extern "C" int ckks_comm_drop_last(ckks_comm *c, size_t k, ckks_comm **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    if (!ok_comm(c)) return CKKS_BAD_HANDLE;
    ckks_comm *d = new ckks_comm();
    d->magic = MAGIC_COMM;
    for (ckks_ctx *x : c->ctx) {
        ckks_ctx *y = nullptr;
        int rc = ckks_ctx_drop_last(x, k, &y);
        if (rc != CKKS_OK) {
            ckks_comm_destroy(d);
            return rc;
        }
        d->ctx.push_back(y);
    }
    *out = d;
    return CKKS_OK;
}
extern "C" int ckks_comm_size(const ckks_comm *c) { return ok_comm(c) ? (int)c->ctx.size() : 0; }
extern "C" ckks_ctx *ckks_comm_ctx(ckks_comm *c, int i) { return (ok_comm(c) && i >= 0 && (size_t)i < c->ctx.size()) ? c->ctx[(size_t)i] : nullptr; }

extern "C" int ckks_comm_ksk_free(ckks_comm_ksk *k) {
    if (!(k && k->magic == MAGIC_COMM_KSK)) return CKKS_BAD_HANDLE;
    for (ckks_ksk *x : k->key)
        if (x) ckks_ksk_free(x);
    k->magic = 0;
    delete k;
    return CKKS_OK;
}
// One host key (engine.rs:225-253, layout of ckks_ksk_upload) -> a transformed replica on every device; the uploads
// and transforms of the devices run concurrently.
extern "C" int ckks_comm_ksk_upload(ckks_comm *c, const uint64_t *a, const uint64_t *b, ckks_comm_ksk **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    if (!ok_comm(c)) return CKKS_BAD_HANDLE;
    if (!a || !b) return CKKS_BAD_ARGUMENT;
    const size_t nd = c->ctx.size();
    ckks_comm_ksk *k = new ckks_comm_ksk();
    k->magic = MAGIC_COMM_KSK;
    k->comm = c;
    k->L = c->ctx[0]->L;
    k->key.assign(nd, nullptr);
    std::vector<int> rcs(nd, CKKS_OK);
    std::vector<std::string> errs(nd);
    std::vector<std::thread> pool;
    for (size_t i = 0; i < nd; ++i)
        pool.emplace_back([&, i]() {
            rcs[i] = ckks_ksk_upload(c->ctx[i], a, b, &k->key[i]);
            if (rcs[i] != CKKS_OK) errs[i] = g_err;
        });
    for (auto &t : pool) t.join();
    for (size_t i = 0; i < nd; ++i)
        if (rcs[i] != CKKS_OK) {
            g_err = errs[i];
            int rc = rcs[i];
            ckks_comm_ksk_free(k);
            return rc;
        }
    *out = k;
    return CKKS_OK;
}

// Share of device i: contiguous, sizes differ by at most one ciphertext.
static void comm_share(size_t batch, size_t nd, size_t i, size_t *s0, size_t *nb) {
    const size_t base = batch / nd, rem = batch % nd;
    *nb = base + (i < rem ? 1 : 0);
    *s0 = i * base + (i < rem ? i : rem);
}
static int comm_run(ckks_comm *c, int kind, const ckks_comm_ksk *key, int32_t rot, size_t batch, const u64 *const *hin, u64 *const *hout) {
    const size_t nd = c->ctx.size();
    const size_t L = c->ctx[0]->L, n = c->ctx[0]->T->n;
    const size_t wi = L * n, wo = (kind == 0 ? L - 1 : L) * n;
    const int n_in = kind == 0 ? 4 : 2;
    std::vector<int> rcs(nd, CKKS_OK);
    std::vector<std::string> errs(nd);
    std::vector<std::thread> pool;
    for (size_t i = 0; i < nd; ++i) {
        size_t s0, nb;
        comm_share(batch, nd, i, &s0, &nb);
        if (!nb) continue;
        pool.emplace_back([&, i, s0, nb]() {
            ckks_ctx *x = c->ctx[i];
            if (!ok_ctx(x)) {  // also binds this thread's launch accounting to the context's stream
                rcs[i] = CKKS_BAD_HANDLE;
                return;
            }
            const u64 *in[4] = {nullptr, nullptr, nullptr, nullptr};
            for (int t = 0; t < n_in; ++t) in[t] = hin[t] + s0 * wi;
            u64 *o[2] = {hout[0] + s0 * wo, hout[1] + s0 * wo};
            rcs[i] = host_pipeline(x, kind, key->key[i], rot, nb, in, o);
            if (rcs[i] != CKKS_OK) errs[i] = g_err;
        });
    }
    for (auto &t : pool) t.join();
    for (size_t i = 0; i < nd; ++i)
        if (rcs[i] != CKKS_OK) {
            g_err = errs[i];
            return rcs[i];
        }
    return CKKS_OK;
}
// mul_ciphertexts_gadget (engine.rs:473-539) + rescale_ciphertext (engine.rs:263-282) of `batch` host ciphertext
// pairs ([batch][L][N] per component, reference layout), spread over the devices of the group.
extern "C" int ckks_comm_ct_mul_relin_rescale_host(ckks_comm *c, const ckks_comm_ksk *rlk, size_t batch, const uint64_t *a0,
                                                   const uint64_t *a1, const uint64_t *b0, const uint64_t *b1, uint64_t *o0,
                                                   uint64_t *o1) {
    if (!ok_comm(c) || !ok_comm_ksk(rlk)) return CKKS_BAD_HANDLE;
    if (batch && (!a0 || !a1 || !b0 || !b1 || !o0 || !o1)) return CKKS_BAD_ARGUMENT;
    if (c->ctx[0]->L < 2) return CKKS_INVALID_MOD_DROP;
    if (rlk->comm->ctx.size() != c->ctx.size() || rlk->L != c->ctx[0]->L) return CKKS_BASIS_MISMATCH;
    for (size_t i = 0; i < c->ctx.size(); ++i)
        if (!same_basis(c->ctx[i], rlk->key[i]->ctx)) return CKKS_BASIS_MISMATCH;
    const u64 *hin[4] = {(const u64 *)a0, (const u64 *)a1, (const u64 *)b0, (const u64 *)b1};
    u64 *hout[2] = {(u64 *)o0, (u64 *)o1};
    return comm_run(c, 0, rlk, 0, batch, hin, hout);
}
// rotate_ciphertext (engine.rs:412-463) of `batch` host ciphertexts, spread over the devices of the group.
extern "C" int ckks_comm_ct_rotate_host(ckks_comm *c, const ckks_comm_ksk *rotk, int32_t k, size_t batch, const uint64_t *c0,
                                        const uint64_t *c1, uint64_t *o0, uint64_t *o1) {
    if (!ok_comm(c) || !ok_comm_ksk(rotk)) return CKKS_BAD_HANDLE;
    if (batch && (!c0 || !c1 || !o0 || !o1)) return CKKS_BAD_ARGUMENT;
    if (rotk->comm->ctx.size() != c->ctx.size() || rotk->L != c->ctx[0]->L) return CKKS_BASIS_MISMATCH;
    for (size_t i = 0; i < c->ctx.size(); ++i)
        if (!same_basis(c->ctx[i], rotk->key[i]->ctx)) return CKKS_BASIS_MISMATCH;
    const u64 *hin[4] = {(const u64 *)c0, (const u64 *)c1, nullptr, nullptr};
    u64 *hout[2] = {(u64 *)o0, (u64 *)o1};
    return comm_run(c, 1, rotk, k, batch, hin, hout);
}

// What the host side alone sustains: `n_src` H2D copies of [hsrc, hsrc + src_bytes) and `n_dst` D2H copies into
// [hdst, hdst + dst_bytes), both directions concurrently on two streams, in the pipeline's chunk size, no kernels.
// Returns seconds per iteration (0 on error).  bench.py reports it next to `e2e` as the ceiling of the host-buffer
// entry points on the box it runs on.
extern "C" double ckks_bench_host_copy(int device, const void *hsrc, size_t src_bytes, int n_src, void *hdst, size_t dst_bytes,
                                       int n_dst, int iters) {
    if (ckks_device_count() <= device || device < 0 || iters < 1) return 0.0;
    if (cudaSetDevice(device) != cudaSuccess) return 0.0;
    const size_t chunk = (size_t)g_host_chunk_mib << 20;
    cudaStream_t si = nullptr, so = nullptr;
    unsigned char *din[2] = {nullptr, nullptr}, *dout[2] = {nullptr, nullptr};
    bool ok = cudaStreamCreateWithFlags(&si, cudaStreamNonBlocking) == cudaSuccess && cudaStreamCreateWithFlags(&so, cudaStreamNonBlocking) == cudaSuccess;
    for (int b = 0; b < 2 && ok; ++b) ok = cudaMalloc((void **)&din[b], chunk) == cudaSuccess && cudaMalloc((void **)&dout[b], chunk) == cudaSuccess;
    double sec = 0.0;
    if (ok) {
        auto pass = [&]() {
            unsigned k = 0;
            for (int r = 0; r < n_src; ++r)
                for (size_t off = 0; off < src_bytes; off += chunk, ++k)
                    cudaMemcpyAsync(din[k & 1], (const unsigned char *)hsrc + off, src_bytes - off < chunk ? src_bytes - off : chunk, cudaMemcpyHostToDevice, si);
            k = 0;
            for (int r = 0; r < n_dst; ++r)
                for (size_t off = 0; off < dst_bytes; off += chunk, ++k)
                    cudaMemcpyAsync((unsigned char *)hdst + off, dout[k & 1], dst_bytes - off < chunk ? dst_bytes - off : chunk, cudaMemcpyDeviceToHost, so);
        };
        pass();  // warm-up
        cudaStreamSynchronize(si);
        cudaStreamSynchronize(so);
        auto t0 = std::chrono::steady_clock::now();
        for (int it = 0; it < iters; ++it) pass();
        cudaStreamSynchronize(si);
        cudaStreamSynchronize(so);
        sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / iters;
        if (cudaGetLastError() != cudaSuccess) sec = 0.0;
    }
    for (int b = 0; b < 2; ++b) {
        if (din[b]) cudaFree(din[b]);
        if (dout[b]) cudaFree(dout[b]);
    }
    if (si) cudaStreamDestroy(si);
    if (so) cudaStreamDestroy(so);
    return sec;
}
