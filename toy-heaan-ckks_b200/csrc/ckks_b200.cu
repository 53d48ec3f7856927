// ckks_b200.cu -- libckks_b200.so: the C ABI of include/ckks_b200.h over the sm_100a kernels.
// No torch types, no CPU fallback: every compute entry point needs a CUDA device.
#include "../../include/ckks_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <cmath>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "context.hpp"
#include "host_math.hpp"
#include "tables_host.hpp"
#include "encoder.cuh"
#include "crt_wide.cuh"
#include "kernels.cuh"
#include "aux_ks.cuh"

// -------------------------------------------------------------------------------------------------
// errors, launch accounting
// -------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};
static std::mutex g_mu;
static std::map<std::string, uint64_t> g_launch_table;
static int g_ntt_path = 0;
static int g_force_unfused = 0;
static int g_host_chunk_mib = 64;
static int g_use_fused_ntt = 1;  // test hook: 0 runs 2^12..2^14 transforms as two passes through global memory
static int g_use_tma = 1;    // test hook: 0 stages ks_pass2 tiles with cp.async instead of TMA
static int g_allow_lazy8 = 1;  // test hook: 0 keeps Harvey [0,4q) butterflies even when q < 2^61
static int g_allow_w32 = 1;  // test hook: 0 forces 64-bit words even for small moduli  // test hook: run the unfused key-switch building blocks

static int cuda_fail(cudaError_t e, const char *what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return CKKS_CUDA_ERROR;
}
#define CU(x)                                          \
    do {                                               \
        cudaError_t e__ = (x);                         \
        if (e__ != cudaSuccess) return cuda_fail(e__, #x); \
    } while (0)
#define TRY(x)                      \
    do {                            \
        int s__ = (x);              \
        if (s__ != CKKS_OK) return s__; \
    } while (0)

static void count_launch(const char *name) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    std::lock_guard<std::mutex> lk(g_mu);
    g_launch_table[name]++;
}
// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg).
struct ProfRec {
    const char *name;
    cudaEvent_t e0, e1;
};
static std::atomic<bool> g_prof{false};
static std::mutex g_prof_mu;  // guards the two vectors below
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t prof_event() {
    cudaEvent_t e;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof_pool.empty()) {
        e = g_prof_pool.back();
        g_prof_pool.pop_back();
    } else {
        cudaEventCreate(&e);
    }
    return e;
}
struct ProfScope {
    ProfRec r;
    cudaStream_t s;
    bool on;
    ProfScope(const char *name, cudaStream_t st) : s(st), on(g_prof.load(std::memory_order_relaxed)) {
        if (on) {
            r.name = name;
            r.e0 = prof_event();
            r.e1 = prof_event();
            cudaEventRecord(r.e0, s);
        }
    }
    ~ProfScope() {
        if (on) {
            cudaEventRecord(r.e1, s);
            std::lock_guard<std::mutex> lk(g_prof_mu);
            if (g_prof_recs.size() < 400000) g_prof_recs.push_back(r);
            else {
                g_prof_pool.push_back(r.e0);
                g_prof_pool.push_back(r.e1);
            }
        }
    }
};
// NVTX ranges around every kernel launch (and around the entry points that enqueue many), for Nsight Systems /
// Nsight Compute timelines: off by default, on with ckks_set_nvtx(1) or CKKS_NVTX=1 in the environment.  The
// header-only NVTX v3 is a no-op unless a profiler has injected its library.
static std::atomic<int> g_nvtx{-1};
static bool nvtx_on() {
    int v = g_nvtx.load(std::memory_order_relaxed);
    if (v < 0) {
        const char *e = getenv("CKKS_NVTX");
        v = (e && e[0] && e[0] != '0') ? 1 : 0;
        g_nvtx.store(v, std::memory_order_relaxed);
    }
    return v != 0;
}
struct NvtxScope {
    bool on;
    explicit NvtxScope(const char *name) : on(nvtx_on()) {
        if (on) nvtxRangePushA(name);
    }
    ~NvtxScope() {
        if (on) nvtxRangePop();
    }
};
extern "C" int ckks_set_nvtx(int on) {
    g_nvtx.store(on != 0 ? 1 : 0);
    return CKKS_OK;
}
static thread_local cudaStream_t g_cur_stream = nullptr;  // stream of the context whose call is running on this thread
// Every launch / allocation helper enqueues on S(T): the context's stream, unless the CALLING THREAD has redirected
// its own launches with a StreamScope (the limb-sharded chunk pipeline runs phase A on an auxiliary stream).  The
// redirection is thread-local: the shared Tables object is never modified, so other threads using the same
// context tree keep launching on the context's stream.
static thread_local cudaStream_t g_stream_override = nullptr;
static inline cudaStream_t S(const Tables &T) { return g_stream_override ? g_stream_override : T.stream; }
struct StreamScope {
    cudaStream_t prev_override, prev_cur;
    explicit StreamScope(cudaStream_t s) : prev_override(g_stream_override), prev_cur(g_cur_stream) {
        g_stream_override = s;
        g_cur_stream = s;
    }
    ~StreamScope() {
        g_stream_override = prev_override;
        g_cur_stream = prev_cur;
    }
};
#define KL(name, ...)                                            \
    do {                                                         \
        count_launch(name);                                      \
        {                                                        \
            NvtxScope nv__(name);                                \
            ProfScope ps__(name, g_cur_stream);                  \
            __VA_ARGS__;                                         \
        }                                                        \
        cudaError_t e__ = cudaPeekAtLastError();                 \
        if (e__ != cudaSuccess) return cuda_fail(e__, name);     \
    } while (0)
// same accounting for the few launches that cannot `return` from the middle of a cleanup path
#define KLV(name, ...)                           \
    do {                                         \
        count_launch(name);                      \
        NvtxScope nv__(name);                    \
        ProfScope ps__(name, g_cur_stream);      \
        __VA_ARGS__;                             \
    } while (0)

extern "C" const char *ckks_status_str(int s) {
    switch (s) {
        case CKKS_OK: return "ok";
        case CKKS_INVALID_DEGREE: return "InvalidDegree";
        case CKKS_EMPTY_BASIS: return "EmptyBasis";
        case CKKS_NON_NTT_FRIENDLY_MODULUS: return "NonNttFriendlyModulus";
        case CKKS_INVALID_MOD_DROP: return "InvalidModDrop";
        case CKKS_CHANNEL_COUNT_MISMATCH: return "ChannelCountMismatch";
        case CKKS_NON_REDUCED_COEFFICIENT: return "NonReducedCoefficient";
        case CKKS_BASIS_MISMATCH: return "BasisMismatch";
        case CKKS_DOMAIN_MISMATCH: return "DomainMismatch";
        case CKKS_BATCH_MISMATCH: return "BatchMismatch";
        case CKKS_LEVEL_MISMATCH: return "LevelMismatch";
        case CKKS_SHORT_INPUT: return "ShortInput";
        case CKKS_BAD_HANDLE: return "BadHandle";
        case CKKS_BAD_ARGUMENT: return "BadArgument";
        case CKKS_UNSUPPORTED: return "Unsupported";
        case CKKS_CUDA_ERROR: return "CudaError";
        case CKKS_NCCL_ERROR: return "NcclError";
    }
    return "unknown";
}
extern "C" const char *ckks_last_error(void) { return g_err.c_str(); }
extern "C" int ckks_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
extern "C" uint64_t ckks_launch_count(void) { return g_launches.load(); }
extern "C" size_t ckks_launch_table(char *buf, size_t cap) {
    std::lock_guard<std::mutex> lk(g_mu);
    std::string s;
    for (auto &kv : g_launch_table) s += kv.first + "=" + std::to_string(kv.second) + "\n";
    if (buf && cap) {
        size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
        memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return s.size() + 1;
}
extern "C" int ckks_set_host_chunk_mib(int mib) {
    if (mib < 1) return CKKS_BAD_ARGUMENT;
    g_host_chunk_mib = mib;
    return CKKS_OK;
}
extern "C" int ckks_set_fused_ntt(int on) {
    g_use_fused_ntt = on != 0;
    return CKKS_OK;
}
extern "C" int ckks_set_tma(int on) {
    g_use_tma = on != 0;
    return CKKS_OK;
}
extern "C" int ckks_set_lazy8(int on) {
    g_allow_lazy8 = on != 0;
    return CKKS_OK;
}
extern "C" int ckks_set_word32(int on) {
    g_allow_w32 = on != 0;
    return CKKS_OK;
}
extern "C" int ckks_set_unfused(int on) {
    g_force_unfused = on != 0;
    return CKKS_OK;
}
extern "C" int ckks_set_ntt_path(int p) {
    if (p < 0 || p > 2) return CKKS_BAD_ARGUMENT;
    g_ntt_path = p;
    return CKKS_OK;
}

// -------------------------------------------------------------------------------------------------
// host number theory (src/math)
// -------------------------------------------------------------------------------------------------
extern "C" int ckks_is_prime(uint64_t n) { return hm::is_prime(n) ? 1 : 0; }
extern "C" int ckks_is_ntt_friendly_prime(uint64_t p, uint64_t n) { return hm::is_ntt_friendly_prime(p, n) ? 1 : 0; }
extern "C" int ckks_generate_primes(int bits, int count, uint64_t degree, uint64_t *out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    return hm::generate_primes(bits, count, degree, (u64 *)out) ? CKKS_OK : CKKS_BAD_ARGUMENT;
}

// -------------------------------------------------------------------------------------------------
// context
// -------------------------------------------------------------------------------------------------
static void destroy_host_pipe(struct HostPipe *p);
static void destroy_aux(struct AuxKs *a);
static void ws_release(Tables &T);
static void cache_release(Tables &T);
Tables::~Tables() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    void *ptrs[] = {d_lc, d_psi, d_psi_inv, d_ninv, d_P1, d_P1i, d_W2, d_W2i, d_TT, d_TTi, d_TTt, d_qlinv, d_qlinv_w == d_qlinv ? nullptr : d_qlinv_w};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    destroy_host_pipe(pipe);
    destroy_aux(aux);
    ws_release(*this);
    cache_release(*this);
    if (stream) cudaStreamSynchronize(stream);
    if (pool) cudaMemPoolDestroy(pool);
    if (own_stream && stream) cudaStreamDestroy(stream);
}

static bool ok_ctx(const ckks_ctx *c) {
    if (!(c && c->magic == MAGIC_CTX)) return false;
    g_cur_stream = S(*c->T);
    return true;
}
static bool ok_poly(const ckks_poly *p) { return p && p->magic == MAGIC_POLY && ok_ctx(p->ctx); }
static bool ok_ksk(const ckks_ksk *k) { return k && k->magic == MAGIC_KSK && ok_ctx(k->ctx) && k->digits == k->ctx->L; }
static bool ok_ksk_slice(const ckks_ksk *k) { return k && k->magic == MAGIC_KSK && ok_ctx(k->ctx); }
static void ctx_ref(ckks_ctx *c) { c->refs.fetch_add(1); }
static void ctx_unref(ckks_ctx *c) {
    if (c->refs.fetch_sub(1) == 1) {
        c->magic = 0;
        delete c;
    }
}
static bool same_basis(const ckks_ctx *a, const ckks_ctx *b) { return a->T.get() == b->T.get() && a->L == b->L; }

// Every stream-ordered allocation of the library comes from the context's PRIVATE memory pool (created in
// ckks_ctx_create, released with the tables): freed scratch stays cached there for the next call instead of in the
// device's default pool, which other libraries in the process (e.g. PyTorch) share; ckks_ctx_trim returns it.
// allocator statistics (ckks_alloc_stats): requests served by the block cache, requests that went to the driver's
// pool, and the host time the latter took (the pool growing or re-mapping shows up here)
static std::atomic<uint64_t> g_alloc_hits{0}, g_alloc_pool{0}, g_alloc_pool_us{0}, g_alloc_pool_max_us{0};
static cudaError_t raw_pool_malloc(const Tables &T, void **p, size_t bytes) {
    const bool big = bytes >= ((size_t)1 << 20);
    auto t0 = std::chrono::steady_clock::now();
    cudaError_t e = T.pool ? cudaMallocFromPoolAsync(p, bytes, T.pool, S(T)) : cudaMallocAsync(p, bytes, S(T));
    if (big) {
        uint64_t us = (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        g_alloc_pool.fetch_add(1);
        g_alloc_pool_us.fetch_add(us);
        uint64_t m = g_alloc_pool_max_us.load();
        while (us > m && !g_alloc_pool_max_us.compare_exchange_weak(m, us)) {
        }
    }
    return e;
}
extern "C" int ckks_alloc_stats(uint64_t *cache_hits, uint64_t *pool_allocs, uint64_t *pool_us, uint64_t *pool_max_us, int reset) {
    if (cache_hits) *cache_hits = g_alloc_hits.load();
    if (pool_allocs) *pool_allocs = g_alloc_pool.load();
    if (pool_us) *pool_us = g_alloc_pool_us.load();
    if (pool_max_us) *pool_max_us = g_alloc_pool_max_us.load();
    if (reset) {
        g_alloc_hits = 0;
        g_alloc_pool = 0;
        g_alloc_pool_us = 0;
        g_alloc_pool_max_us = 0;
    }
    return CKKS_OK;
}
static cudaError_t pool_malloc(const Tables &Tc, void **p, size_t bytes) {
    // small requests, and calls whose launches are redirected to another stream, go straight to the pool
    if (bytes < ((size_t)1 << 20) || g_stream_override) return raw_pool_malloc(Tc, p, bytes);
    Tables &T = const_cast<Tables &>(Tc);  // the arena is internal state guarded by cache_mu
    std::lock_guard<std::mutex> lk(T.cache_mu);
    cudaError_t last = cudaSuccess;
    bool hit = false;
    *p = T.arena.take(
        bytes,
        [&](size_t seg) -> void * {
            void *base = nullptr;
            last = raw_pool_malloc(T, &base, seg);
            if (last != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
            return base;
        },
        [&](void *base) {
            cudaFreeAsync(base, T.stream);
            cudaStreamSynchronize(T.stream);
            if (T.pool) cudaMemPoolTrimTo(T.pool, 0);
        },
        &hit);
    if (!*p) return last != cudaSuccess ? last : cudaErrorMemoryAllocation;
    if (hit) g_alloc_hits.fetch_add(1);
    return cudaSuccess;
}
// Give the segments that are entirely free back to the pool (ckks_ctx_trim, destruction).
static void cache_release(Tables &T) {
    std::lock_guard<std::mutex> lk(T.cache_mu);
    T.arena.release_free_segments([&](void *base) { cudaFreeAsync(base, T.stream); });
}

static void dev_free(const Tables &Tc, void *p);
enum { WS_A0 = 0, WS_A1, WS_B0, WS_B1, WS_TMP, WS_SCR, WS_LAST, WS_COUNT };
static int ws_get(const Tables &Tc, int slot, size_t bytes, u64 **out) {
    Tables &T = const_cast<Tables &>(Tc);  // the cache is internal state guarded by ws_mu
    Tables::WsSlot &w = T.ws[slot];
    if (w.bytes < bytes || !w.p) {
        if (w.p) dev_free(T, w.p);  // stream-ordered: the kernels that still use it run first
        w.p = nullptr;
        w.bytes = 0;
        CU(pool_malloc(T, &w.p, bytes ? bytes : 8));
        w.bytes = bytes;
    }
    *out = reinterpret_cast<u64 *>(w.p);
    return CKKS_OK;
}
static void ws_release(Tables &T) {
    for (auto &w : T.ws) {
        if (w.p) dev_free(T, w.p);
        w.p = nullptr;
        w.bytes = 0;
    }
}

template <class T>
static int upload_vec(T **dst, const std::vector<T> &v) {
    CU(cudaMalloc((void **)dst, v.size() * sizeof(T)));
    CU(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return CKKS_OK;
}
static tw_t mk_tw(u64 w, u64 q) { return ht::mk_tw(w, q); }

static int build_tables(Tables &T, bool allow_w32, bool allow_lazy8) {
    ht::HostTables H;
    ht::build_host_tables(T.n, T.logn, T.path, T.a1, T.a2, T.moduli, T.psi, H, allow_w32, allow_lazy8);
    T.lazy = H.lazy;
    T.w32 = H.w32;
    T.digit_reduce = H.digit_reduce;
    T.w2_stride = H.w2_stride;
    TRY(upload_vec(&T.d_lc, H.lc));
    TRY(upload_vec(&T.d_qlinv, H.ql));
    if (T.path == 1) {
        TRY(upload_vec(&T.d_psi, H.psi));
        TRY(upload_vec(&T.d_psi_inv, H.psii));
        TRY(upload_vec(&T.d_ninv, H.ninv));
        return CKKS_OK;
    }
    if (H.w32) {
        TRY(upload_vec((tw32_t **)&T.d_P1, H.P1_32));
        TRY(upload_vec((tw32_t **)&T.d_P1i, H.P1i_32));
        TRY(upload_vec((tw32_t **)&T.d_W2, H.W2_32));
        TRY(upload_vec((tw32_t **)&T.d_W2i, H.W2i_32));
        TRY(upload_vec((tw32_t **)&T.d_TT, H.TT_32));
        TRY(upload_vec((tw32_t **)&T.d_TTi, H.TTi_32));
        TRY(upload_vec((tw32_t **)&T.d_TTt, H.TTt_32));
        TRY(upload_vec((tw32_t **)&T.d_qlinv_w, H.ql32));
        return CKKS_OK;
    }
    TRY(upload_vec((tw_t **)&T.d_P1, H.P1));
    TRY(upload_vec((tw_t **)&T.d_P1i, H.P1i));
    TRY(upload_vec((tw_t **)&T.d_W2, H.W2));
    TRY(upload_vec((tw_t **)&T.d_W2i, H.W2i));
    TRY(upload_vec((tw_t **)&T.d_TT, H.TT));
    TRY(upload_vec((tw_t **)&T.d_TTi, H.TTi));
    TRY(upload_vec((tw_t **)&T.d_TTt, H.TTt));
    T.d_qlinv_w = T.d_qlinv;
    return CKKS_OK;
}

extern "C" int ckks_ctx_create(uint64_t n, const uint64_t *moduli, size_t l, int device, ckks_ctx **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    // RnsBasis::new: EmptyBasis first (basis.rs:98-100), then per-table checks (basis.rs:22-31).
    if (l == 0) return CKKS_EMPTY_BASIS;
    if (!moduli) return CKKS_BAD_ARGUMENT;
    if (n == 0 || (n & (n - 1)) != 0) return CKKS_INVALID_DEGREE;
    for (size_t i = 0; i < l; ++i)
        if (!hm::is_ntt_friendly_prime(moduli[i], n)) return CKKS_NON_NTT_FRIENDLY_MODULUS;
    int logn = 0;
    while (((u64)1 << logn) < n) ++logn;
    if (logn > 16) return CKKS_UNSUPPORTED;
    for (size_t i = 0; i < l; ++i)
        if (moduli[i] >> 63) return CKKS_UNSUPPORTED;  // the reference's add_mod wraps there too
    if (ckks_device_count() <= device || device < 0) {
        g_err = "no CUDA device " + std::to_string(device) + " (this library has no CPU fallback)";
        return CKKS_CUDA_ERROR;
    }
    CU(cudaSetDevice(device));
    auto T = std::make_shared<Tables>();
    T->device = device;
    T->n = n;
    T->logn = logn;
    T->L = l;
    T->moduli.assign(moduli, moduli + l);
    int path = g_ntt_path;
    if (path == 0) path = (logn >= 8) ? 2 : 1;
    if (path == 2 && logn < 8) return CKKS_UNSUPPORTED;
    if (path == 1 && logn > 11) return CKKS_UNSUPPORTED;
    T->path = path;
    if (path == 2) {
        T->a1 = (logn + 1) / 2;
        T->a2 = logn - T->a1;
    } else {
        T->a1 = logn;
        T->a2 = 0;
    }
    for (size_t i = 0; i < l; ++i) T->psi.push_back(hm::find_primitive_root(moduli[i], 2 * n));
    CU(cudaStreamCreateWithFlags(&T->stream, cudaStreamNonBlocking));
    T->own_stream = true;
    {
        cudaMemPoolProps props;
        memset(&props, 0, sizeof(props));
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        CU(cudaMemPoolCreate(&T->pool, &props));
        uint64_t thr = ~0ull;  // keep freed scratch cached in OUR pool; the device's default pool is left untouched
        CU(cudaMemPoolSetAttribute(T->pool, cudaMemPoolAttrReleaseThreshold, &thr));
    }
    TRY(build_tables(*T, g_allow_w32 != 0, g_allow_lazy8 != 0));
    ckks_ctx *c = new ckks_ctx();
    c->magic = MAGIC_CTX;
    c->T = T;
    c->L = l;
    c->refs = 1;
    *out = c;
    return CKKS_OK;
}
extern "C" int ckks_ctx_drop_last(ckks_ctx *ctx, size_t k, ckks_ctx **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    if (!ok_ctx(ctx)) return CKKS_BAD_HANDLE;
    if (k >= ctx->L) return CKKS_INVALID_MOD_DROP;  // basis.rs:122-127
    ckks_ctx *c = new ckks_ctx();
    c->magic = MAGIC_CTX;
    c->T = ctx->T;
    c->L = ctx->L - k;
    c->refs = 1;
    *out = c;
    return CKKS_OK;
}
extern "C" int ckks_ctx_destroy(ckks_ctx *ctx) {
    if (!ok_ctx(ctx)) return CKKS_BAD_HANDLE;
    ctx_unref(ctx);
    return CKKS_OK;
}
// A context that is the local share of a limb-sharded group: non-zero once one of its barriers lost a peer.
static int group_failed(const Tables &T) {
    if (T.fail_word && *T.fail_word) {
        g_err = "a limb-sharded barrier on this context timed out waiting for a peer GPU: results are poisoned";
        return CKKS_NCCL_ERROR;
    }
    return CKKS_OK;
}
extern "C" int ckks_ctx_sync(ckks_ctx *ctx) {
    if (!ok_ctx(ctx)) return CKKS_BAD_HANDLE;
    CU(cudaStreamSynchronize(ctx->T->stream));
    return group_failed(*ctx->T);
}
extern "C" int ckks_ctx_set_stream(ckks_ctx *ctx, void *s) {
    if (!ok_ctx(ctx)) return CKKS_BAD_HANDLE;
    Tables &T = *ctx->T;
    CU(cudaStreamSynchronize(S(T)));
    if (T.own_stream) cudaStreamDestroy(S(T));
    T.stream = (cudaStream_t)s;
    T.own_stream = false;
    return CKKS_OK;
}
extern "C" int ckks_ctx_trim(ckks_ctx *ctx) {
    if (!ok_ctx(ctx)) return CKKS_BAD_HANDLE;
    Tables &T = *ctx->T;
    CU(cudaSetDevice(T.device));
    CU(cudaStreamSynchronize(S(T)));
    if (T.pool) CU(cudaMemPoolTrimTo(T.pool, 0));
    return CKKS_OK;
}
extern "C" uint64_t ckks_ctx_degree(const ckks_ctx *c) { return ok_ctx(c) ? c->T->n : 0; }
extern "C" size_t ckks_ctx_channel_count(const ckks_ctx *c) { return ok_ctx(c) ? c->L : 0; }
extern "C" int ckks_ctx_moduli(const ckks_ctx *c, uint64_t *out) {
    if (!ok_ctx(c) || !out) return CKKS_BAD_HANDLE;
    for (size_t i = 0; i < c->L; ++i) out[i] = c->T->moduli[i];
    return CKKS_OK;
}
extern "C" uint32_t ckks_ctx_total_bits(const ckks_ctx *c) {
    if (!ok_ctx(c)) return 0;
    uint32_t s = 0;
    for (size_t i = 0; i < c->L; ++i) s += 63 - (uint32_t)__builtin_clzll(c->T->moduli[i]);
    return s;
}
extern "C" uint64_t ckks_ctx_psi(const ckks_ctx *c, size_t ch) { return (ok_ctx(c) && ch < c->L) ? c->T->psi[ch] : 0; }
// NttTable<N> of one channel in the REFERENCE's layout (basis.rs:6-17, 61-83), rebuilt on the host from psi for
// callers that read `basis.ntt_table(ch)`: which = 0 forward_roots[i] = omega^i, 1 inverse_roots[i] = omega^-i,
// 2 twist_factors[j] = psi^j, 3 untwist_factors[j] = psi^-j (N words each), 4 n_inv (one word).  The device
// tables are laid out differently (tables_host.hpp); this accessor is the drop-in view.
extern "C" int ckks_ctx_ntt_table(const ckks_ctx *c, size_t channel, int which, uint64_t *out) {
    if (!ok_ctx(c)) return CKKS_BAD_HANDLE;
    if (channel >= c->L || !out || which < 0 || which > 4) return CKKS_BAD_ARGUMENT;
    const u64 q = c->T->moduli[channel], n = c->T->n, psi = c->T->psi[channel];
    if (which == 4) {
        out[0] = hm::inv_mod(n % q, q);
        return CKKS_OK;
    }
    u64 base = which < 2 ? hm::mul_mod(psi, psi, q) : psi;  // omega = psi^2 (basis.rs:36)
    if (which & 1) base = hm::inv_mod(base, q);
    u64 v = 1;
    for (u64 i = 0; i < n; ++i) {
        out[i] = v;
        v = hm::mul_mod(v, base, q);
    }
    return CKKS_OK;
}
extern "C" int ckks_ctx_reconstruct_centered_coeff(const ckks_ctx *c, const uint64_t *res, int64_t *out) {
    if (!ok_ctx(c) || !res || !out) return CKKS_BAD_HANDLE;
    std::vector<u64> m(c->T->moduli.begin(), c->T->moduli.begin() + c->L);
    *out = hm::reconstruct_centered(m, (const u64 *)res);
    return CKKS_OK;
}

// -------------------------------------------------------------------------------------------------
// launch helpers
// -------------------------------------------------------------------------------------------------
static EwArgs ew_args(const Tables &T, size_t L, size_t batch) {
    EwArgs a;
    a.lc = T.d_lc;
    a.poly = L * T.n;
    a.total = batch * a.poly;
    a.logn = T.logn;
    a.L = (int)L;
    return a;
}
static unsigned ew_grid(size_t total) {
    size_t g = (total + 255) / 256;
    size_t cap = 148 * 32;
    return (unsigned)(g < cap ? (g ? g : 1) : cap);
}

// from_channels' reducedness scan (poly.rs:83-93) of [batch][L][N] device words: ORs 1 into *flag on a word >= q.
static int scan_reduced(const Tables &T, size_t L, size_t batch, const u64 *d, int *flag, cudaStream_t s) {
    EwArgs a = ew_args(T, L, batch);
    if (!a.total) return CKKS_OK;
    KL("check_reduced", (check_reduced_kernel<<<ew_grid(a.total), 256, 0, s>>>(a, d, flag)));
    return CKKS_OK;
}

// The same scan for two freshly uploaded buffers, synchronous: CKKS_NON_REDUCED_COEFFICIENT on a hit.
static int scan_reduced_pair_sync(const Tables &T, size_t L, size_t batch, const u64 *x, const u64 *y) {
    int *flag = nullptr, hflag = 0;
    CU(pool_malloc(T, (void **)&flag, sizeof(int)));
    cudaMemsetAsync(flag, 0, sizeof(int), S(T));
    int rc = scan_reduced(T, L, batch, x, flag, S(T));
    if (rc == CKKS_OK && y) rc = scan_reduced(T, L, batch, y, flag, S(T));
    cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, S(T));
    if (cudaStreamSynchronize(S(T)) != cudaSuccess && rc == CKKS_OK) rc = cuda_fail(cudaGetLastError(), "reducedness scan");
    cudaFreeAsync(flag, S(T));
    if (rc == CKKS_OK && hflag) rc = CKKS_NON_REDUCED_COEFFICIENT;  // poly.rs:83-93
    return rc;
}

// Dispatch on the lazy mode (0 strict, 1 Harvey, 2 lazy8); 32-bit words never use mode 2.
#ifdef CKKS_ONLY_CFG4
#define W32_DISPATCH(...) CKKS_UNSUPPORTED
#else
#define W32_DISPATCH(...) (__VA_ARGS__)
#endif
#ifdef CKKS_ONLY_CFG4  // development builds (tools/variants.sh): only the cfg4 instantiations, compiles in under a minute
#define LZ_SWITCH(WDT, lazy, M)                                  \
    do {                                                         \
        if ((lazy) == 2 && sizeof(WDT) == 8) { M(2); }           \
        else return CKKS_UNSUPPORTED;                            \
    } while (0)
#else
#define LZ_SWITCH(WDT, lazy, M)                                  \
    do {                                                         \
        if ((lazy) == 2 && sizeof(WDT) == 8) { M(2); }           \
        else if (lazy) { M(1); }                                 \
        else { M(0); }                                           \
    } while (0)
#endif

// Column tile of the plain passes: 8 for 64-bit words (more, smaller CTAs hide each other's barriers on the
// IMAD-bound path), 16 for 32-bit words (HBM-bound: wider coalesced segments).
template <typename WD>
constexpr int pass_c() {
    return sizeof(WD) == 8 ? 8 : 16;
}
template <typename WD, int KIND, int A, bool PRE, bool POST, bool TR>
static int launch_pass_w(const char *name, int lazy, dim3 grid, cudaStream_t s, const PassArgs &a) {
    constexpr int E = 4, C = pass_c<WD>();
    grid.x = a.ncols / C;
    const size_t smem = (size_t)(1 << A) * (C + 1) * sizeof(WD);
    const int block = C << (A - E);
// FIX_OF(A): the compile-time ring degree this pass size is specialised for (N = 2^16 with 2^8-point passes, N = 2^14
// with 2^7-point passes); the specialisation is used when the context's N matches, the generic kernel otherwise.
#define FIX_OF(Aval) ((Aval) == 8 ? 16 : ((Aval) == 7 ? 14 : 0))
#define FIX_MATCH(Aval, Nval) (FIX_OF(Aval) != 0 && (Nval) == ((size_t)1 << FIX_OF(Aval)))
#define M(LZ)                                                                                                                                  \
    do {                                                                                                                                       \
        if (FIX_MATCH(A, a.N))                                                                                                                 \
            KL(name, (ntt_pass_kernel<WD, KIND, A, E, C, (sizeof(WD) == 8 ? LZ : (LZ ? 1 : 0)), PRE, POST, TR, false, false, FIX_OF(A)><<<grid, block, smem, s>>>(a))); \
        else                                                                                                                                   \
            KL(name, (ntt_pass_kernel<WD, KIND, A, E, C, (sizeof(WD) == 8 ? LZ : (LZ ? 1 : 0)), PRE, POST, TR><<<grid, block, smem, s>>>(a)));       \
    } while (0)
    LZ_SWITCH(WD, lazy, M);
#undef M
    return CKKS_OK;
}
template <int KIND, int A, bool PRE, bool POST, bool TR>
static int launch_pass_a(const char *name, bool w32, int lazy, dim3 grid, cudaStream_t s, const PassArgs &a) {
    if (w32) return W32_DISPATCH(launch_pass_w<u32, KIND, A, PRE, POST, TR>(name, lazy, grid, s, a));
    return launch_pass_w<u64, KIND, A, PRE, POST, TR>(name, lazy, grid, s, a);
}
template <int KIND, bool PRE, bool POST, bool TR>
static int launch_pass(const char *name, int A, bool w32, int lazy, dim3 grid, cudaStream_t s, const PassArgs &a) {
    switch (A) {
#ifndef CKKS_ONLY_CFG4
        case 4: return launch_pass_a<KIND, 4, PRE, POST, TR>(name, w32, lazy, grid, s, a);
        case 5: return launch_pass_a<KIND, 5, PRE, POST, TR>(name, w32, lazy, grid, s, a);
        case 6: return launch_pass_a<KIND, 6, PRE, POST, TR>(name, w32, lazy, grid, s, a);
        case 7: return launch_pass_a<KIND, 7, PRE, POST, TR>(name, w32, lazy, grid, s, a);
#endif
        case 8: return launch_pass_a<KIND, 8, PRE, POST, TR>(name, w32, lazy, grid, s, a);
    }
    return CKKS_UNSUPPORTED;
}

// ---- four-step pass launchers --------------------------------------------------------------------
// A "limb range" [limb0, limb0 + nl) of [nb][L][N] words; dst may have a different limb count.
struct Span {
    size_t nb;      // polynomials
    int L;          // limbs per polynomial in src
    int limb0, nl;  // limbs processed
    int dstL, dst_limb0;
};
static Span whole(size_t nb, size_t L) { return Span{nb, (int)L, 0, (int)L, (int)L, 0}; }

enum { P_FWD1, P_FWD2, P_INV2, P_INV1 };
// src/dst are u64 words except the intermediate between the two passes of a transform (dst of
// P_FWD1 / P_INV2, src of P_FWD2 / P_INV1), which holds the transform word type (u32 when T.w32).
static int run_pass(const Tables &T, int which, Span sp, const void *src_, void *dst_) {
    if (sp.nb == 0 || sp.nl == 0) return CKKS_OK;
    const size_t wsz = T.w32 ? 4 : 8;
    const size_t ssz = (which == P_FWD2 || which == P_INV1) ? wsz : 8, dsz = (which == P_FWD1 || which == P_INV2) ? wsz : 8;
    const char *src = (const char *)src_;
    char *dst = (char *)dst_;
    const unsigned n1 = 1u << T.a1, n2 = 1u << T.a2;
    cudaStream_t s = S(T);
    for (size_t b0 = 0; b0 < sp.nb; b0 += 32768) {
        size_t nb = sp.nb - b0 < 32768 ? sp.nb - b0 : 32768;
        PassArgs a;
        memset(&a, 0, sizeof(a));
        a.lc = T.d_lc;
        a.L = sp.L;
        a.N = T.n;
        a.limb0 = sp.limb0;
        a.dstL = sp.dstL;
        a.dst_limb0 = sp.dst_limb0;
        a.src = src + b0 * sp.L * T.n * ssz;
        a.dst = dst + b0 * sp.dstL * T.n * dsz;
        a.elt = nullptr;
        switch (which) {
            case P_FWD1:
                a.tab = T.d_P1;
                a.tab_stride = n1;
                a.ncols = n2;
                TRY((launch_pass<XF_NEG_FWD, false, false, true>("ntt_fwd_pass1", T.a1, T.w32, T.lazy, dim3(n2 / 16, sp.nl, (unsigned)nb), s, a)));
                break;
            case P_FWD2:
                a.tab = T.d_W2;
                a.tab_stride = T.w2_stride;
                a.elt = T.d_TT;
                a.ncols = n1;
                TRY((launch_pass<XF_CYC_FWD, true, false, false>("ntt_fwd_pass2", T.a2, T.w32, T.lazy, dim3(n1 / 16, sp.nl, (unsigned)nb), s, a)));
                break;
            case P_INV2:
                a.tab = T.d_W2i;
                a.tab_stride = T.w2_stride;
                a.elt = T.d_TTi;
                a.ncols = n1;
                TRY((launch_pass<XF_CYC_INV, false, true, true>("ntt_inv_pass2", T.a2, T.w32, T.lazy, dim3(n1 / 16, sp.nl, (unsigned)nb), s, a)));
                break;
            case P_INV1:
                a.tab = T.d_P1i;
                a.tab_stride = n1;
                a.ncols = n2;
                TRY((launch_pass<XF_NEG_INV, false, false, false>("ntt_inv_pass1", T.a1, T.w32, T.lazy, dim3(n2 / 16, sp.nl, (unsigned)nb), s, a)));
                break;
        }
    }
    return CKKS_OK;
}

// Fused single-CTA transform (2^12 <= N <= 2^14): the limb stays in shared memory between the passes.
template <typename WD, int A1, int A2>
static int launch_fused_w(const Tables &T, size_t L, size_t batch, u64 *d, bool inverse) {
    FusedArgs a;
    a.data = d;
    a.lc = T.d_lc;
    a.P1 = inverse ? T.d_P1i : T.d_P1;
    a.W2 = inverse ? T.d_W2i : T.d_W2;
    a.TT = inverse ? T.d_TTi : T.d_TTt;
    a.w2_stride = T.w2_stride;
    a.L = (int)L;
    const size_t smem = fused_smem_words<A1, A2>() * sizeof(WD);
    const int block = (1 << (A1 + A2)) / 16;
    const unsigned grid = (unsigned)(batch * L);
#define FUSED(LZ, IV)                                                                                                  \
    do {                                                                                                               \
        if (smem > 48 * 1024)                                                                                          \
            CU(cudaFuncSetAttribute(ntt_fused_kernel<WD, A1, A2, LZ, IV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        KL(IV ? "ntt_fused_inv" : "ntt_fused_fwd", (ntt_fused_kernel<WD, A1, A2, LZ, IV><<<grid, block, smem, S(T)>>>(a)));       \
    } while (0)
#define M(LZ)                                                           \
    do {                                                                \
        constexpr int LZZ = sizeof(WD) == 8 ? LZ : (LZ ? 1 : 0);        \
        if (inverse) FUSED(LZZ, true);                                  \
        else FUSED(LZZ, false);                                         \
    } while (0)
    LZ_SWITCH(WD, T.lazy, M);
#undef M
#undef FUSED
    return CKKS_OK;
}
static int ntt_fused_run(const Tables &T, size_t L, size_t batch, u64 *d, bool inverse) {
#define FW(A1v, A2v) (T.w32 ? W32_DISPATCH(launch_fused_w<u32, A1v, A2v>(T, L, batch, d, inverse)) : launch_fused_w<u64, A1v, A2v>(T, L, batch, d, inverse))
    if (T.a1 == 6 && T.a2 == 6) return FW(6, 6);
    if (T.a1 == 7 && T.a2 == 6) return FW(7, 6);
    if (T.a1 == 7 && T.a2 == 7) return FW(7, 7);
#undef FW
    return CKKS_UNSUPPORTED;
}

// Forward / inverse transform of [batch][L][N] words in place (tmp: same size, four-step only).
static int ntt_run(const Tables &T, size_t L, size_t batch, u64 *d, u64 *tmp, bool inverse) {
    if (batch == 0) return CKKS_OK;
    cudaStream_t s = S(T);
    if (T.path == 1) {
        SmallArgs a;
        a.data = d;
        a.lc = T.d_lc;
        a.psi = inverse ? T.d_psi_inv : T.d_psi;
        a.ninv = T.d_ninv;
        a.L = (int)L;
        a.logn = T.logn;
        int block = (int)(T.n / 2 < 32 ? 32 : (T.n / 2 > 256 ? 256 : T.n / 2));
        size_t smem = T.n * sizeof(u64);
        unsigned grid = (unsigned)(batch * L);
        if (!inverse) {
            if (T.lazy) KL("ntt_small_fwd", (ntt_small_kernel<false, 1><<<grid, block, smem, s>>>(a)));
            else KL("ntt_small_fwd", (ntt_small_kernel<false, 0><<<grid, block, smem, s>>>(a)));
        } else {
            if (T.lazy) KL("ntt_small_inv", (ntt_small_kernel<true, 1><<<grid, block, smem, s>>>(a)));
            else KL("ntt_small_inv", (ntt_small_kernel<true, 0><<<grid, block, smem, s>>>(a)));
        }
        return CKKS_OK;
    }
    // Measured on B200 (profiles/): with 32-bit words the single-kernel transform wins up to N = 2^14;
    // with 64-bit words (register- and IMAD-bound) only at N = 2^12.
    // (at N = 2^14 the 1024-thread fused kernel wins forward, 3.73 vs 3.36 TB/s, and loses inverse, 3.21 vs 3.63)
    if (g_use_fused_ntt && ((T.w32 && T.logn >= 12 && (T.logn <= 13 || (T.logn == 14 && !inverse))) || (!T.w32 && T.logn == 12)))
        return ntt_fused_run(T, L, batch, d, inverse);
    // Two passes through an intermediate of the transform word type.  (Walking the batch in chunks with
    // the intermediate pinned in L2 by an access-policy window was measured slower on B200 -- the carve-out
    // and the small launches cost more than the saved HBM round trip -- and is not used.)
    Span sp = whole(batch, L);
    if (!inverse) {
        TRY(run_pass(T, P_FWD1, sp, d, tmp));
        TRY(run_pass(T, P_FWD2, sp, tmp, d));
    } else {
        TRY(run_pass(T, P_INV2, sp, d, tmp));
        TRY(run_pass(T, P_INV1, sp, tmp, d));
    }
    return CKKS_OK;
}

static int dev_alloc(const Tables &T, size_t words, u64 **out) {
    *out = nullptr;
    if (words == 0) words = 1;
    CU(pool_malloc(T, (void **)out, words * sizeof(u64)));
    return CKKS_OK;
}
static void dev_free(const Tables &Tc, void *p) {
    if (!p) return;
    Tables &T = const_cast<Tables &>(Tc);
    {
        std::lock_guard<std::mutex> lk(T.cache_mu);
        if (T.arena.give(p)) return;  // back into the arena, merged with its free neighbours
    }
    cudaFreeAsync(p, S(T));
}
static int ntt_inplace(const Tables &T, size_t L, size_t batch, u64 *d, bool inverse) {
    u64 *tmp = nullptr;
    if (T.path == 2) TRY(dev_alloc(T, batch * L * T.n, &tmp));
    int rc = ntt_run(T, L, batch, d, tmp, inverse);
    dev_free(T, tmp);
    return rc;
}

// -------------------------------------------------------------------------------------------------
// polynomials
// -------------------------------------------------------------------------------------------------
static int poly_new(ckks_ctx *ctx, size_t batch, bool ntt, ckks_poly **out) {
    const Tables &T = *ctx->T;
    CU(cudaSetDevice(T.device));
    u64 *d;
    TRY(dev_alloc(T, batch * ctx->L * T.n, &d));
    ckks_poly *p = new ckks_poly();
    p->magic = MAGIC_POLY;
    p->ctx = ctx;
    ctx_ref(ctx);
    p->batch = batch;
    p->d = d;
    p->ntt = ntt;
    *out = p;
    return CKKS_OK;
}
static size_t poly_words(const ckks_poly *p) { return p->batch * p->ctx->L * p->ctx->T->n; }

extern "C" int ckks_poly_free(ckks_poly *p) {
    if (!ok_poly(p)) return CKKS_BAD_HANDLE;
    cudaSetDevice(p->ctx->T->device);
    dev_free(*p->ctx->T, p->d);
    ckks_ctx *c = p->ctx;
    p->magic = 0;
    delete p;
    ctx_unref(c);
    return CKKS_OK;
}
extern "C" int ckks_poly_alloc(ckks_ctx *ctx, size_t batch, ckks_poly **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    if (!ok_ctx(ctx)) return CKKS_BAD_HANDLE;
    TRY(poly_new(ctx, batch, false, out));
    CU(cudaMemsetAsync((*out)->d, 0, poly_words(*out) * sizeof(u64), ctx->T->stream));
    return CKKS_OK;
}
extern "C" int ckks_poly_clone(ckks_poly *p, ckks_poly **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    if (!ok_poly(p)) return CKKS_BAD_HANDLE;
    TRY(poly_new(p->ctx, p->batch, p->ntt, out));
    CU(cudaMemcpyAsync((*out)->d, p->d, poly_words(p) * sizeof(u64), cudaMemcpyDeviceToDevice, p->ctx->T->stream));
    return CKKS_OK;
}
extern "C" size_t ckks_poly_batch(const ckks_poly *p) { return ok_poly(p) ? p->batch : 0; }
extern "C" size_t ckks_poly_channel_count(const ckks_poly *p) { return ok_poly(p) ? p->ctx->L : 0; }
extern "C" int ckks_poly_is_ntt_domain(const ckks_poly *p) { return ok_poly(p) ? (p->ntt ? 1 : 0) : -1; }

extern "C" int ckks_poly_from_coeffs(ckks_ctx *ctx, size_t batch, const int64_t *coeffs, size_t clen, ckks_poly **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    if (!ok_ctx(ctx)) return CKKS_BAD_HANDLE;
    const Tables &T = *ctx->T;
    if (clen < T.n) return CKKS_SHORT_INPUT;
    if (!coeffs && batch) return CKKS_BAD_ARGUMENT;
    TRY(poly_new(ctx, batch, false, out));
    if (batch == 0) return CKKS_OK;
    i64 *dc;
    CU(pool_malloc(T, (void **)&dc, batch * clen * sizeof(i64)));
    CU(cudaMemcpyAsync(dc, coeffs, batch * clen * sizeof(i64), cudaMemcpyHostToDevice, S(T)));
    EwArgs a = ew_args(T, ctx->L, batch);
    KL("from_coeffs", (from_coeffs_kernel<<<ew_grid(a.total), 256, 0, S(T)>>>(a, dc, clen, (*out)->d)));
    dev_free(T, dc);
    CU(cudaStreamSynchronize(S(T)));  // `coeffs` may be pageable and reused by the caller
    return CKKS_OK;
}

static int permute(const Tables &T, const u64 *src, u64 *dst, size_t words, bool to_internal) {
    if (!words) return CKKS_OK;
    KL("permute_ntt", (permute_ntt_kernel<<<(unsigned)((words + 255) / 256), 256, 0, S(T)>>>(src, dst, words, T.logn, T.a1,
                                                                                               T.a2, to_internal ? 1 : 0)));
    return CKKS_OK;
}

extern "C" int ckks_poly_from_channels(ckks_ctx *ctx, size_t batch, const uint64_t *ch, size_t nch, int in_ntt,
                                       ckks_poly **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    if (!ok_ctx(ctx)) return CKKS_BAD_HANDLE;
    if (nch != ctx->L) return CKKS_CHANNEL_COUNT_MISMATCH;  // poly.rs:77-82
    if (!ch && batch) return CKKS_BAD_ARGUMENT;
    const Tables &T = *ctx->T;
    ckks_poly *p;
    TRY(poly_new(ctx, batch, in_ntt != 0, &p));
    size_t words = poly_words(p);
    int rc = CKKS_OK;
    if (words) {
        int *flag = nullptr;
        int hflag = 0;
        u64 *stage = nullptr;
        do {
            if (pool_malloc(T, (void **)&flag, sizeof(int)) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "malloc"); break; }
            cudaMemsetAsync(flag, 0, sizeof(int), S(T));
            u64 *land = p->d;
            if (in_ntt) {
                if ((rc = dev_alloc(T, words, &stage)) != CKKS_OK) break;
                land = stage;
            }
            if (cudaMemcpyAsync(land, ch, words * sizeof(u64), cudaMemcpyHostToDevice, S(T)) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "h2d"); break; }
            EwArgs a = ew_args(T, ctx->L, batch);
            KLV("check_reduced", (check_reduced_kernel<<<ew_grid(a.total), 256, 0, S(T)>>>(a, land, flag)));
            if (in_ntt && (rc = permute(T, stage, p->d, words, true)) != CKKS_OK) break;
            cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, S(T));
            if (cudaStreamSynchronize(S(T)) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "sync"); break; }
            if (hflag) rc = CKKS_NON_REDUCED_COEFFICIENT;  // poly.rs:83-93
        } while (0);
        dev_free(T, flag);
        dev_free(T, stage);
    }
    if (rc != CKKS_OK) {
        ckks_poly_free(p);
        return rc;
    }
    *out = p;
    return CKKS_OK;
}

extern "C" int ckks_poly_download(ckks_poly *p, uint64_t *out) {
    if (!ok_poly(p)) return CKKS_BAD_HANDLE;
    const Tables &T = *p->ctx->T;
    size_t words = poly_words(p);
    if (!words) return CKKS_OK;
    if (!out) return CKKS_BAD_ARGUMENT;
    CU(cudaSetDevice(T.device));
    if (p->ntt) {
        u64 *stage;
        TRY(dev_alloc(T, words, &stage));
        TRY(permute(T, p->d, stage, words, false));
        CU(cudaMemcpyAsync(out, stage, words * sizeof(u64), cudaMemcpyDeviceToHost, S(T)));
        dev_free(T, stage);
    } else {
        CU(cudaMemcpyAsync(out, p->d, words * sizeof(u64), cudaMemcpyDeviceToHost, S(T)));
    }
    CU(cudaStreamSynchronize(S(T)));
    return group_failed(T);
}

extern "C" int ckks_poly_to_ntt_domain(ckks_poly *p) {
    if (!ok_poly(p)) return CKKS_BAD_HANDLE;
    if (p->ntt) return CKKS_OK;  // poly.rs:137-139
    CU(cudaSetDevice(p->ctx->T->device));
    TRY(ntt_inplace(*p->ctx->T, p->ctx->L, p->batch, p->d, false));
    p->ntt = true;
    return CKKS_OK;
}
extern "C" int ckks_poly_to_coeff_domain(ckks_poly *p) {
    if (!ok_poly(p)) return CKKS_BAD_HANDLE;
    if (!p->ntt) return CKKS_OK;  // poly.rs:155-157
    CU(cudaSetDevice(p->ctx->T->device));
    TRY(ntt_inplace(*p->ctx->T, p->ctx->L, p->batch, p->d, true));
    p->ntt = false;
    return CKKS_OK;
}

// rhs may have batch 1 (broadcast) or a->batch.
static int check_pair(const ckks_poly *a, const ckks_poly *b, bool allow_bcast) {
    if (!ok_poly(a) || !ok_poly(b)) return CKKS_BAD_HANDLE;
    if (!same_basis(a->ctx, b->ctx)) return CKKS_BASIS_MISMATCH;
    if (a->ntt != b->ntt) return CKKS_DOMAIN_MISMATCH;
    if (a->batch != b->batch && !(allow_bcast && b->batch == 1)) return CKKS_BATCH_MISMATCH;
    return CKKS_OK;
}
template <int OP>
static int ew_binary(const char *name, ckks_poly *a, const ckks_poly *b) {
    const Tables &T = *a->ctx->T;
    CU(cudaSetDevice(T.device));
    EwArgs e = ew_args(T, a->ctx->L, a->batch);
    if (!e.total) return CKKS_OK;
    size_t bs = (b->batch == a->batch) ? e.poly : 0;
    KL(name, (ew_binary_kernel<OP><<<ew_grid(e.total), 256, 0, S(T)>>>(e, a->d, b->d, bs)));
    return CKKS_OK;
}
extern "C" int ckks_poly_add_assign(ckks_poly *a, const ckks_poly *b) {
    TRY(check_pair(a, b, true));
    return ew_binary<EW_ADD>("ew_add", a, b);
}
extern "C" int ckks_poly_sub_assign(ckks_poly *a, const ckks_poly *b) {
    TRY(check_pair(a, b, true));
    return ew_binary<EW_SUB>("ew_sub", a, b);
}
extern "C" int ckks_poly_neg(ckks_poly *a) {
    if (!ok_poly(a)) return CKKS_BAD_HANDLE;
    const Tables &T = *a->ctx->T;
    CU(cudaSetDevice(T.device));
    EwArgs e = ew_args(T, a->ctx->L, a->batch);
    if (!e.total) return CKKS_OK;
    KL("ew_neg", (ew_neg_kernel<<<ew_grid(e.total), 256, 0, S(T)>>>(e, a->d)));
    return CKKS_OK;
}
extern "C" int ckks_poly_mul_assign(ckks_poly *a, const ckks_poly *b) {
    TRY(check_pair(a, b, true));
    const Tables &T = *a->ctx->T;
    CU(cudaSetDevice(T.device));
    if (a->ntt) return ew_binary<EW_MUL>("ew_mul", a, b);  // poly.rs:297-306
    // poly.rs:307-329: forward both, pointwise, inverse; result in the coefficient domain.
    ckks_poly *r;
    TRY(ckks_poly_clone(const_cast<ckks_poly *>(b), &r));
    int rc = ckks_poly_to_ntt_domain(r);
    if (rc == CKKS_OK) rc = ckks_poly_to_ntt_domain(a);
    if (rc == CKKS_OK) rc = ew_binary<EW_MUL>("ew_mul", a, r);
    if (rc == CKKS_OK) rc = ckks_poly_to_coeff_domain(a);
    ckks_poly_free(r);
    return rc;
}

// mul_assign_naive (poly.rs:339-367): coefficient-domain operands only (the reference debug_asserts it).
extern "C" int ckks_poly_mul_assign_naive(ckks_poly *a, const ckks_poly *b) {
    TRY(check_pair(a, b, true));
    if (a->ntt) return CKKS_DOMAIN_MISMATCH;
    const Tables &T = *a->ctx->T;
    CU(cudaSetDevice(T.device));
    EwArgs e = ew_args(T, a->ctx->L, a->batch);
    if (!e.total) return CKKS_OK;
    u64 *tmp = nullptr;
    TRY(dev_alloc(T, e.total, &tmp));
    KLV("mul_naive", (mul_naive_kernel<<<ew_grid(e.total), 256, 0, S(T)>>>(e, a->d, b->d, b->batch == 1 && a->batch != 1 ? 0 : e.poly, tmp)));
    cudaMemcpyAsync(a->d, tmp, e.total * 8, cudaMemcpyDeviceToDevice, S(T));
    dev_free(T, tmp);
    if (cudaPeekAtLastError() != cudaSuccess) return cuda_fail(cudaGetLastError(), "mul_naive");
    return CKKS_OK;
}

extern "C" int ckks_poly_mod_drop_last(const ckks_poly *p, ckks_ctx *child, ckks_poly **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    if (!ok_poly(p) || !ok_ctx(child)) return CKKS_BAD_HANDLE;
    if (child->T.get() != p->ctx->T.get() || child->L > p->ctx->L) return CKKS_BASIS_MISMATCH;
    if (child->L == 0) return CKKS_INVALID_MOD_DROP;
    const Tables &T = *child->T;
    CU(cudaSetDevice(T.device));
    TRY(poly_new(child, p->batch, p->ntt, out));
    if (p->batch)
        CU(cudaMemcpy2DAsync((*out)->d, child->L * T.n * sizeof(u64), p->d, p->ctx->L * T.n * sizeof(u64),
                             child->L * T.n * sizeof(u64), p->batch, cudaMemcpyDeviceToDevice, S(T)));
    return CKKS_OK;
}

static int rescale_dev(const Tables &T, size_t L, size_t batch, const u64 *src, u64 *dst) {
    EwArgs e = ew_args(T, L - 1, batch);
    if (!e.total) return CKKS_OK;
    KL("rescale", (rescale_kernel<<<ew_grid(e.total), 256, 0, S(T)>>>(e, src, dst, T.d_qlinv + (L - 1) * T.L)));
    return CKKS_OK;
}

extern "C" int ckks_poly_rescale_into(const ckks_poly *p, ckks_ctx *child, ckks_poly **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    if (!ok_poly(p) || !ok_ctx(child)) return CKKS_BAD_HANDLE;
    if (p->ctx->L < 2) return CKKS_INVALID_MOD_DROP;  // poly.rs:191-197
    if (child->T.get() != p->ctx->T.get() || child->L + 1 != p->ctx->L) return CKKS_BASIS_MISMATCH;
    const Tables &T = *child->T;
    CU(cudaSetDevice(T.device));
    const u64 *src = p->d;
    ckks_poly *tmp = nullptr;
    if (p->ntt) {  // poly.rs:199-208
        TRY(ckks_poly_clone(const_cast<ckks_poly *>(p), &tmp));
        int rc = ckks_poly_to_coeff_domain(tmp);
        if (rc != CKKS_OK) {
            ckks_poly_free(tmp);
            return rc;
        }
        src = tmp->d;
    }
    int rc = poly_new(child, p->batch, false, out);
    if (rc == CKKS_OK) rc = rescale_dev(T, p->ctx->L, p->batch, src, (*out)->d);
    if (tmp) ckks_poly_free(tmp);
    if (rc != CKKS_OK && *out) {
        ckks_poly_free(*out);
        *out = nullptr;
    }
    return rc;
}

static u64 inv_mod_pow2(u64 e, u64 m) {  // e odd, m a power of two
    u64 x = e;
    for (int i = 0; i < 6; ++i) x *= 2 - e * x;
    return x & (m - 1);
}

extern "C" int ckks_poly_automorphism(const ckks_poly *p, uint64_t exponent, ckks_poly **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    if (!ok_poly(p)) return CKKS_BAD_HANDLE;
    const Tables &T = *p->ctx->T;
    CU(cudaSetDevice(T.device));
    const u64 two_n = 2 * T.n;
    const u64 e = exponent % two_n;
    if (e == 0) return ckks_poly_clone(const_cast<ckks_poly *>(p), out);  // poly.rs:508-511
    const u64 *src = p->d;
    ckks_poly *tmp = nullptr;
    if (p->ntt) {  // poly.rs:494-503
        TRY(ckks_poly_clone(const_cast<ckks_poly *>(p), &tmp));
        int rc = ckks_poly_to_coeff_domain(tmp);
        if (rc != CKKS_OK) {
            ckks_poly_free(tmp);
            return rc;
        }
        src = tmp->d;
    }
    int rc = poly_new(p->ctx, p->batch, false, out);
    EwArgs a = ew_args(T, p->ctx->L, p->batch);
    if (rc == CKKS_OK && a.total) {
        if (e & 1) {
            KLV("automorphism", (automorphism_kernel<<<ew_grid(a.total), 256, 0, S(T)>>>(a, src, (*out)->d, e, inv_mod_pow2(e, two_n))));
        } else {
            unsigned *win = nullptr;
            if (pool_malloc(T, (void **)&win, a.total * sizeof(unsigned)) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "malloc");
            if (rc == CKKS_OK) {
                cudaMemsetAsync(win, 0, a.total * sizeof(unsigned), S(T));
                KLV("automorphism_even_mark", (automorphism_even_mark_kernel<<<ew_grid(a.total), 256, 0, S(T)>>>(a, src, win, e)));
                KLV("automorphism_even_fill", (automorphism_even_fill_kernel<<<ew_grid(a.total), 256, 0, S(T)>>>(a, src, win, (*out)->d, e)));
                dev_free(T, win);
            }
        }
        if (rc == CKKS_OK && cudaPeekAtLastError() != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "automorphism");
    }
    if (tmp) ckks_poly_free(tmp);
    if (rc != CKKS_OK && *out) {
        ckks_poly_free(*out);
        *out = nullptr;
    }
    return rc;
}

static u64 rot_exponent(u64 n, int32_t k) {
    u64 r = k >= 0 ? (u64)k : (u64)(-(int64_t)k);
    return hm::pow_mod(5, r, 2 * n);  // poly.rs:549-552
}
extern "C" int ckks_poly_rotate_slots(const ckks_poly *p, int32_t k, ckks_poly **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    if (!ok_poly(p)) return CKKS_BAD_HANDLE;
    const u64 n = p->ctx->T->n;
    u64 e = rot_exponent(n, k);
    if (k >= 0) return ckks_poly_automorphism(p, e, out);
    ckks_poly *mid;  // poly.rs:556-566: automorphism(5^|k|) then automorphism(2N-1)
    TRY(ckks_poly_automorphism(p, e, &mid));
    int rc = ckks_poly_automorphism(mid, 2 * n - 1, out);
    ckks_poly_free(mid);
    return rc;
}

extern "C" int ckks_poly_to_coeffs(const ckks_poly *p, int64_t *out) {
    if (!ok_poly(p)) return CKKS_BAD_HANDLE;
    const Tables &T = *p->ctx->T;
    const size_t L = p->ctx->L, n = T.n;
    if (!p->batch) return CKKS_OK;
    if (!out) return CKKS_BAD_ARGUMENT;
    ckks_poly *c;
    TRY(ckks_poly_clone(const_cast<ckks_poly *>(p), &c));
    int rc = ckks_poly_to_coeff_domain(c);
    std::vector<u64> h(poly_words(p));
    if (rc == CKKS_OK) rc = ckks_poly_download(c, (uint64_t *)h.data());
    ckks_poly_free(c);
    TRY(rc);
    // The centred CRT (basis.rs:158-180) is the decode side of the path; it runs on the host.
    std::vector<u64> mod(T.moduli.begin(), T.moduli.begin() + L), res(L);
    for (size_t b = 0; b < p->batch; ++b)
        for (size_t i = 0; i < n; ++i) {
            for (size_t l = 0; l < L; ++l) res[l] = h[(b * L + l) * n + i];
            out[b * n + i] = hm::reconstruct_centered(mod, res.data());
        }
    return CKKS_OK;
}

// Multi-word floor(Q / 2) in Garner's mixed radix: digit j = (floor(Q/2) div q_0 ... q_{j-1}) mod q_j.
using hm::half_q_digits;  // mixed-radix digits of floor(Q / 2) (host_math.hpp)
// Centred CRT of every coefficient for a basis of any size (device, Garner mixed radix: crt_wide.cuh).
// d_i64 / d_f64: device buffers [batch][N] or null; *overflow: some |x| >= 2^63.
static int crt_wide_dev(const ckks_poly *p, long long *d_i64, double *d_f64, int *overflow) {
    const Tables &T = *p->ctx->T;
    const size_t L = p->ctx->L;
    if (L > (size_t)CRT_MAX_L) return CKKS_UNSUPPORTED;
    std::vector<u64> mod(T.moduli.begin(), T.moduli.begin() + L);
    std::vector<tw_t> inv(L * L, mk_tw(0, mod[0]));
    for (size_t i = 0; i < L; ++i)
        for (size_t j = i + 1; j < L; ++j) inv[i * L + j] = mk_tw(hm::inv_mod(mod[i] % mod[j], mod[j]), mod[j]);
    std::vector<u64> half = half_q_digits(mod);
    ckks_poly *c = nullptr;
    TRY(ckks_poly_clone(const_cast<ckks_poly *>(p), &c));
    tw_t *dinv = nullptr;
    u64 *dhalf = nullptr;
    int *dflag = nullptr, hflag = 0;
    auto body = [&]() -> int {
        TRY(ckks_poly_to_coeff_domain(c));  // poly.rs:405-411
        CU(pool_malloc(T, (void **)&dinv, inv.size() * sizeof(tw_t)));
        CU(pool_malloc(T, (void **)&dhalf, half.size() * sizeof(u64)));
        CU(pool_malloc(T, (void **)&dflag, sizeof(int)));
        CU(cudaMemcpyAsync(dinv, inv.data(), inv.size() * sizeof(tw_t), cudaMemcpyHostToDevice, S(T)));
        CU(cudaMemcpyAsync(dhalf, half.data(), half.size() * sizeof(u64), cudaMemcpyHostToDevice, S(T)));
        CU(cudaMemsetAsync(dflag, 0, sizeof(int), S(T)));
        CrtWideArgs a;
        a.src = c->d;
        a.lc = T.d_lc;
        a.inv = dinv;
        a.half = dhalf;
        a.out_i64 = d_i64;
        a.out_f64 = d_f64;
        a.overflow = dflag;
        a.total = p->batch * T.n;
        a.L = (int)L;
        a.logn = T.logn;
        KL("crt_wide", (crt_wide_kernel<<<ew_grid(a.total), 256, 0, S(T)>>>(a)));
        CU(cudaMemcpyAsync(&hflag, dflag, sizeof(int), cudaMemcpyDeviceToHost, S(T)));
        CU(cudaStreamSynchronize(S(T)));  // inv / half leave scope
        return CKKS_OK;
    };
    int rc = body();
    dev_free(T, dinv);
    dev_free(T, dhalf);
    dev_free(T, dflag);
    ckks_poly_free(c);
    if (overflow) *overflow = hflag;
    return rc;
}
extern "C" int ckks_poly_to_coeffs_wide(const ckks_poly *p, int64_t *out_i64, double *out_f64, int *overflow) {
    if (overflow) *overflow = 0;
    if (!ok_poly(p)) return CKKS_BAD_HANDLE;
    const Tables &T = *p->ctx->T;
    if (!p->batch) return CKKS_OK;
    if (!out_i64 && !out_f64) return CKKS_BAD_ARGUMENT;
    CU(cudaSetDevice(T.device));
    const size_t words = p->batch * T.n;
    long long *di = nullptr;
    double *dd = nullptr;
    auto body = [&]() -> int {
        if (out_i64) CU(pool_malloc(T, (void **)&di, words * sizeof(long long)));
        if (out_f64) CU(pool_malloc(T, (void **)&dd, words * sizeof(double)));
        TRY(crt_wide_dev(p, di, dd, overflow));
        if (out_i64) CU(cudaMemcpyAsync(out_i64, di, words * sizeof(long long), cudaMemcpyDeviceToHost, S(T)));
        if (out_f64) CU(cudaMemcpyAsync(out_f64, dd, words * sizeof(double), cudaMemcpyDeviceToHost, S(T)));
        CU(cudaStreamSynchronize(S(T)));
        return CKKS_OK;
    };
    int rc = body();
    dev_free(T, di);
    dev_free(T, dd);
    return rc;
}

// -------------------------------------------------------------------------------------------------
// gadget keys
// -------------------------------------------------------------------------------------------------
constexpr int KS_E2 = 3, KS_C2 = 4;  // ks_pass2: 8 elements per thread leave room for the 128-bit accumulators
static int ksk_new(ckks_ctx *ctx, ckks_ksk **out, size_t digits = 0) {
    const Tables &T = *ctx->T;
    if (!digits) digits = ctx->L;
    size_t words = digits * ctx->L * T.n;
    u64 *a, *b;
    TRY(dev_alloc(T, words, &a));
    TRY(dev_alloc(T, words, &b));
    ckks_ksk *k = new ckks_ksk();
    k->magic = MAGIC_KSK;
    k->ctx = ctx;
    ctx_ref(ctx);
    k->a = a;
    k->b = b;
    k->digits = digits;
    k->perm_e = -1;
    k->k32 = false;
    k->xa = k->xb = nullptr;
    k->aux_k = 0;
    *out = k;
    return CKKS_OK;
}
extern "C" int ckks_ksk_free(ckks_ksk *k) {
    if (!ok_ksk_slice(k)) return CKKS_BAD_HANDLE;
    cudaSetDevice(k->ctx->T->device);
    dev_free(*k->ctx->T, k->a);
    dev_free(*k->ctx->T, k->b);
    if (k->xa) dev_free(*k->ctx->T, k->xa);
    if (k->xb) dev_free(*k->ctx->T, k->xb);
    ckks_ctx *c = k->ctx;
    k->magic = 0;
    delete k;
    ctx_unref(c);
    return CKKS_OK;
}
static int ksk_add_aux(const Tables &T, ckks_ksk *k, const u64 *a_coeff, const u64 *b_coeff);  // aux_ks.inl
// Four-step path: store the transformed key with the rows of every limb in ks_pass2's order (perm_row) and in
// the transform word type (u32 words on the 32-bit path: the key is the largest stream ks_pass2 stages).
static int ksk_finalize(const Tables &T, ckks_ksk *k) {
    if (T.path != 2 || T.a2 <= KS_E2) return CKKS_OK;
    const size_t words = k->digits * k->ctx->L * T.n;
    if (!words) return CKKS_OK;
    u64 *tmp = nullptr;
    TRY(dev_alloc(T, words, &tmp));
    for (u64 *p : {k->a, k->b}) {
        if (T.w32) {
            KLV("key_permute", (key_permute_kernel<u32><<<ew_grid(words), 256, 0, S(T)>>>(p, reinterpret_cast<u32 *>(tmp), words, T.logn, T.a1, T.a2, KS_E2)));
            cudaMemcpyAsync(p, tmp, words * 4, cudaMemcpyDeviceToDevice, S(T));
        } else {
            KLV("key_permute", (key_permute_kernel<u64><<<ew_grid(words), 256, 0, S(T)>>>(p, tmp, words, T.logn, T.a1, T.a2, KS_E2)));
            cudaMemcpyAsync(p, tmp, words * 8, cudaMemcpyDeviceToDevice, S(T));
        }
    }
    dev_free(T, tmp);
    if (cudaPeekAtLastError() != cudaSuccess) return cuda_fail(cudaGetLastError(), "key_permute");
    k->perm_e = KS_E2;
    k->k32 = T.w32;
    return CKKS_OK;
}
extern "C" int ckks_ksk_upload(ckks_ctx *ctx, const uint64_t *a, const uint64_t *b, ckks_ksk **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    if (!ok_ctx(ctx)) return CKKS_BAD_HANDLE;
    if (!a || !b) return CKKS_BAD_ARGUMENT;
    const Tables &T = *ctx->T;
    CU(cudaSetDevice(T.device));
    ckks_ksk *k;
    TRY(ksk_new(ctx, &k));
    size_t words = ctx->L * ctx->L * T.n;
    int rc = CKKS_OK;
    if (cudaMemcpyAsync(k->a, a, words * 8, cudaMemcpyHostToDevice, S(T)) != cudaSuccess ||
        cudaMemcpyAsync(k->b, b, words * 8, cudaMemcpyHostToDevice, S(T)) != cudaSuccess)
        rc = cuda_fail(cudaGetLastError(), "ksk h2d");
    // the key polynomials are RnsPoly values on the reference side (engine.rs:225-253): same canonical-word rule
    if (rc == CKKS_OK) rc = scan_reduced_pair_sync(T, ctx->L, ctx->L, k->a, k->b);
    if (rc == CKKS_OK) rc = ksk_add_aux(T, k, k->a, k->b);  // from the coefficient-domain words
    if (rc == CKKS_OK) rc = ntt_inplace(T, ctx->L, ctx->L, k->a, false);
    if (rc == CKKS_OK) rc = ntt_inplace(T, ctx->L, ctx->L, k->b, false);
    if (rc == CKKS_OK) rc = ksk_finalize(T, k);
    if (rc == CKKS_OK && cudaStreamSynchronize(S(T)) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "ksk sync");
    if (rc != CKKS_OK) {
        ckks_ksk_free(k);
        return rc;
    }
    *out = k;
    return CKKS_OK;
}
extern "C" int ckks_ksk_from_polys(const ckks_poly *a, const ckks_poly *b, ckks_ksk **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    if (!ok_poly(a) || !ok_poly(b)) return CKKS_BAD_HANDLE;
    if (!same_basis(a->ctx, b->ctx)) return CKKS_BASIS_MISMATCH;
    ckks_ctx *ctx = a->ctx;
    if (a->batch != ctx->L || b->batch != ctx->L) return CKKS_BATCH_MISMATCH;
    const Tables &T = *ctx->T;
    CU(cudaSetDevice(T.device));
    ckks_ksk *k;
    TRY(ksk_new(ctx, &k));
    size_t words = ctx->L * ctx->L * T.n;
    cudaMemcpyAsync(k->a, a->d, words * 8, cudaMemcpyDeviceToDevice, S(T));
    cudaMemcpyAsync(k->b, b->d, words * 8, cudaMemcpyDeviceToDevice, S(T));
    int rc = CKKS_OK;
    // the auxiliary form is derived from coefficient-domain words
    if (a->ntt) rc = ntt_inplace(T, ctx->L, ctx->L, k->a, true);
    if (rc == CKKS_OK && b->ntt) rc = ntt_inplace(T, ctx->L, ctx->L, k->b, true);
    if (rc == CKKS_OK) rc = ksk_add_aux(T, k, k->a, k->b);
    if (rc == CKKS_OK) rc = ntt_inplace(T, ctx->L, ctx->L, k->a, false);
    if (rc == CKKS_OK) rc = ntt_inplace(T, ctx->L, ctx->L, k->b, false);
    if (rc == CKKS_OK) rc = ksk_finalize(T, k);
    if (rc != CKKS_OK) {
        ckks_ksk_free(k);
        return rc;
    }
    *out = k;
    return CKKS_OK;
}

// Copy limb i of `target` (batch 1) into limb i of polynomial i of `dst` (batch L), adding.
__global__ void add_gadget_target_kernel(EwArgs a /* batch L */, u64 *__restrict__ dst, const u64 *__restrict__ target) {
    const size_t n = (size_t)1 << a.logn;
    size_t total = (size_t)a.L * n;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        size_t i = t >> a.logn, k = t & (n - 1);
        u64 q = a.lc[i].q;
        size_t off = i * a.poly + i * n + k;
        dst[off] = addmod(dst[off], target[i * n + k], q);
    }
}
extern "C" int ckks_gen_gadget_key_b(const ckks_poly *s, const ckks_poly *target, const ckks_poly *a, const ckks_poly *e,
                                     ckks_poly **out_b) {
    if (!out_b) return CKKS_BAD_ARGUMENT;
    *out_b = nullptr;
    if (!ok_poly(s) || !ok_poly(target) || !ok_poly(a) || !ok_poly(e)) return CKKS_BAD_HANDLE;
    ckks_ctx *ctx = s->ctx;
    if (!same_basis(ctx, target->ctx) || !same_basis(ctx, a->ctx) || !same_basis(ctx, e->ctx)) return CKKS_BASIS_MISMATCH;
    if (s->ntt || target->ntt || a->ntt || e->ntt) return CKKS_DOMAIN_MISMATCH;
    if (s->batch != 1 || target->batch != 1 || a->batch != ctx->L || e->batch != ctx->L) return CKKS_BATCH_MISMATCH;
    const Tables &T = *ctx->T;
    // engine.rs:318-327: b_i = a_i.clone(); b_i *= s; b_i = -b_i; b_i += e_i; b_i += plain_poly
    ckks_poly *b;
    TRY(ckks_poly_clone(const_cast<ckks_poly *>(a), &b));
    int rc = ckks_poly_mul_assign(b, s);
    if (rc == CKKS_OK) rc = ckks_poly_neg(b);
    if (rc == CKKS_OK) rc = ckks_poly_add_assign(b, e);
    if (rc == CKKS_OK) {
        EwArgs ea = ew_args(T, ctx->L, ctx->L);
        KLV("add_gadget_target", (add_gadget_target_kernel<<<ew_grid((size_t)ctx->L * T.n), 256, 0, S(T)>>>(ea, b->d, target->d)));
        if (cudaPeekAtLastError() != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "add_gadget_target");
    }
    if (rc != CKKS_OK) {
        ckks_poly_free(b);
        return rc;
    }
    *out_b = b;
    return CKKS_OK;
}

// -------------------------------------------------------------------------------------------------
// ciphertext operations
// -------------------------------------------------------------------------------------------------
// Gadget key-switch accumulate (engine.rs:505-528 / :429-452):
//   acc0 += sum_i NTT(alpha_i) * key_b[i],  acc1 += sum_i NTT(alpha_i) * key_a[i]   (NTT domain)
// `digits`: [batch][L][N] coefficient domain (d2 or the rotated c1).
static int keyswitch_accumulate(const Tables &T, size_t L, size_t batch, const u64 *digits, const ckks_ksk *key, u64 *acc0,
                                u64 *acc1) {
    EwArgs e = ew_args(T, L, batch);
    if (!e.total) return CKKS_OK;
    u64 *alpha = nullptr, *tmp = nullptr;
    int rc = dev_alloc(T, e.total, &alpha);
    if (rc == CKKS_OK && T.path == 2) rc = dev_alloc(T, e.total, &tmp);
    for (size_t i = 0; i < L && rc == CKKS_OK; ++i) {
        KLV("digit_broadcast", (digit_broadcast_kernel<<<ew_grid(e.total), 256, 0, S(T)>>>(e, digits, alpha, (int)i)));
        rc = ntt_run(T, L, batch, alpha, tmp, false);
        if (rc != CKKS_OK) break;
        KLV("ks_mac", (ks_mac_kernel<<<ew_grid(e.total), 256, 0, S(T)>>>(e, alpha, key->k32 ? (const void *)((const u32 *)key->b + i * e.poly) : (const void *)(key->b + i * e.poly),
                                                                                   key->k32 ? (const void *)((const u32 *)key->a + i * e.poly) : (const void *)(key->a + i * e.poly), acc0, acc1, T.a1,
                                                                                   T.a2, key->perm_e, key->k32 ? 1 : 0)));
        if (cudaPeekAtLastError() != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "keyswitch");
    }
    dev_free(T, alpha);
    dev_free(T, tmp);
    return rc;
}

// ---- fused four-step key-switch pipeline -----------------------------------------------------------
// ks_pass1 column tile: 16 for both word sizes.  64-bit words: 128-byte row segments for the digit loads and the
// per-element twiddles, 256 threads per CTA (measured on cfg4: 8 columns 163.5 ms per 768 ciphertexts, 16 columns
// 159.2, 4 columns 186.6); 32-bit words: the u64 digit loads were L1-bound with 64-byte segments.
template <typename WD>
constexpr int ks_c1() {
    return 16;
}
#ifndef CKKS_KS1_E
#define CKKS_KS1_E 4
#endif
template <typename WD, int A>
static int launch_ks1_w(int lazy, bool reduce, bool diag, dim3 grid, cudaStream_t s, const KsArgs &a) {
    constexpr int E = (sizeof(WD) == 8 && A > CKKS_KS1_E) ? CKKS_KS1_E : 4, C = ks_c1<WD>();
    grid.x /= C;
    const size_t smem = (size_t)(1 << A) * (C + 1) * sizeof(WD);
    const int block = C << (A - E);
#define KS1(LZ, RD, DG)                                                                                               \
    do {                                                                                                              \
        if (FIX_MATCH(A, a.N) && a.a1 == A) KL("ks_pass1", (ks_pass1_kernel<WD, A, E, C, LZ, RD, DG, FIX_OF(A)><<<grid, block, smem, s>>>(a))); \
        else KL("ks_pass1", (ks_pass1_kernel<WD, A, E, C, LZ, RD, DG><<<grid, block, smem, s>>>(a)));                  \
    } while (0)
    if (lazy == 2 && sizeof(WD) == 8) {
        constexpr int LZ2 = sizeof(WD) == 8 ? 2 : 1;
        if (reduce) { if (diag) KS1(LZ2, true, true); else KS1(LZ2, true, false); }
        else { if (diag) KS1(LZ2, false, true); else KS1(LZ2, false, false); }
    } else if (lazy) {
        if (reduce) { if (diag) KS1(1, true, true); else KS1(1, true, false); }
        else { if (diag) KS1(1, false, true); else KS1(1, false, false); }
    } else {
        if (diag) KS1(0, true, true); else KS1(0, true, false);
    }
#undef KS1
    return CKKS_OK;
}
template <int A>
static int launch_ks1_a(bool w32, int lazy, bool reduce, bool diag, dim3 grid, cudaStream_t s, const KsArgs &a) {
    if (w32) return W32_DISPATCH(launch_ks1_w<u32, A>(lazy, reduce, diag, grid, s, a));
    return launch_ks1_w<u64, A>(lazy, reduce, diag, grid, s, a);
}
// Rank-3 tensor map (cols, rows, slabs) with a [rows][C] box for the TMA path of ks_pass2.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        cudaGetLastError();
    }
    return fn;
}
static bool make_tile_map(unsigned char *out128, const void *base, size_t elem, size_t cols, size_t rows, size_t slabs, unsigned box_cols) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || slabs == 0) return false;
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap image");
    CUtensorMap m;
    cuuint64_t dims[3] = {cols, rows, slabs};
    cuuint64_t strides[2] = {cols * elem, cols * rows * elem};
    cuuint32_t box[3] = {box_cols, (cuuint32_t)rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMapDataType dt = elem == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
    CUresult r = fn(&m, dt, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    memcpy(out128, &m, 128);
    return true;
}

template <typename WD, int A>
static int launch_ks2_w(int lazy, bool mul, bool tma, dim3 grid, cudaStream_t s, const KsArgs &a, const KsMaps &maps) {
    constexpr int E = KS_E2, C = KS_C2;
    const size_t smem = ks2_smem_bytes<WD, A, C>();
    const int block = C << (A - E);
#define KS2(LZ, MU, TM)                                                                                                 \
    do {                                                                                                                \
        if (smem > 48 * 1024)                                                                                           \
            CU(cudaFuncSetAttribute(ks_pass2_kernel<WD, A, E, C, LZ, MU, MU, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        KL(TM ? "ks_pass2_tma" : "ks_pass2", (ks_pass2_kernel<WD, A, E, C, LZ, MU, MU, TM><<<grid, block, smem, s>>>(a, maps)));  \
    } while (0)
#define M(LZ)                                                           \
    do {                                                                \
        constexpr int LZZ = sizeof(WD) == 8 ? LZ : (LZ ? 1 : 0);        \
        if (tma) { if (mul) KS2(LZZ, true, true); else KS2(LZZ, false, true); } \
        else { if (mul) KS2(LZZ, true, false); else KS2(LZZ, false, false); }   \
    } while (0)
    LZ_SWITCH(WD, lazy, M);
#undef M
#undef KS2
    return CKKS_OK;
}
template <int A>
static int launch_ks2_a(bool w32, int lazy, bool mul, bool tma, dim3 grid, cudaStream_t s, const KsArgs &a, const KsMaps &maps) {
    if (w32) return W32_DISPATCH(launch_ks2_w<u32, A>(lazy, mul, tma, grid, s, a, maps));
    return launch_ks2_w<u64, A>(lazy, mul, tma, grid, s, a, maps);
}
template <typename WD, int A>
static int launch_inv1_rescale_w(int lazy, dim3 grid, cudaStream_t s, const PassArgs &a, const u64 *last, const void *ql) {
    constexpr int E = 4, C = pass_c<WD>();
    grid.x = a.ncols / C;
    const size_t smem = (size_t)(1 << A) * (C + 1) * sizeof(WD);
    const int block = C << (A - E);
#define M(LZ)                                                                                                                                         \
    do {                                                                                                                                              \
        if (FIX_MATCH(A, a.N))                                                                                                                        \
            KL("ntt_inv_pass1_rescale", (inv_pass1_rescale_kernel<WD, A, E, C, (sizeof(WD) == 8 ? LZ : (LZ ? 1 : 0)), FIX_OF(A)><<<grid, block, smem, s>>>(a, last, ql))); \
        else                                                                                                                                          \
            KL("ntt_inv_pass1_rescale", (inv_pass1_rescale_kernel<WD, A, E, C, (sizeof(WD) == 8 ? LZ : (LZ ? 1 : 0))><<<grid, block, smem, s>>>(a, last, ql)));    \
    } while (0)
    LZ_SWITCH(WD, lazy, M);
#undef M
    return CKKS_OK;
}
template <int A>
static int launch_inv1_rescale_a(bool w32, int lazy, dim3 grid, cudaStream_t s, const PassArgs &a, const u64 *last, const void *ql) {
    if (w32) return W32_DISPATCH(launch_inv1_rescale_w<u32, A>(lazy, grid, s, a, last, ql));
    return launch_inv1_rescale_w<u64, A>(lazy, grid, s, a, last, ql);
}
#ifdef CKKS_ONLY_CFG4
#define DISPATCH_A(Aval, CALL)                   \
    switch (Aval) {                              \
        case 8: { constexpr int AA = 8; CALL; } break; \
        default: return CKKS_UNSUPPORTED;        \
    }
#else
#define DISPATCH_A(Aval, CALL)                   \
    switch (Aval) {                              \
        case 4: { constexpr int AA = 4; CALL; } break; \
        case 5: { constexpr int AA = 5; CALL; } break; \
        case 6: { constexpr int AA = 6; CALL; } break; \
        case 7: { constexpr int AA = 7; CALL; } break; \
        case 8: { constexpr int AA = 8; CALL; } break; \
        default: return CKKS_UNSUPPORTED;        \
    }
#endif

static int reduce_every_for(const Tables &T, size_t L) {
    u64 qmax = 0;
    for (size_t i = 0; i < L; ++i) qmax = T.moduli[i] > qmax ? T.moduli[i] : qmax;
    // products are x * k with k < q and x < q (x < 2q in lazy8 mode, where ks_pass2 reduces only once)
    hm::u128 sq = (hm::u128)(qmax - 1) * (qmax - 1);
    hm::u128 lim = ~(hm::u128)0;
    if (T.w32) lim = ~(u64)0;  // 32-bit words accumulate in one 64-bit register
    if (T.lazy == 2) lim /= 2;
    hm::u128 t = lim / sq;  // t products fit the accumulator
    if (t > 1) t -= 1;      // one slot for the carried-in residue
    return t > 1000000 ? 1000000 : (int)t;
}

// Key-switch of `cs` polynomials: digits (coefficient domain) [+ dig_ntt] -> transposed inverse-pass-2
// outputs out0t/out1t (finish with P_INV1).  mul: add d0/d1 and use the NTT-domain limb for i == j.
// `sh` describes where the target limbs held here sit in the whole basis: everything (0, 1, L digits) on the
// batch-sharded path, one GPU's share (rank, world, all digits) in limb-sharded mode.
struct KsShard {
    size_t Ld;        // digits (limbs of the whole basis)
    int joff, jstep;  // basis index of local limb j = joff + jstep * j
    size_t dig_ct_stride, dig_limb_stride;  // layout of `digits` in words
    bool reduce;      // digits need `% q_j` before entering the lazy transform
};
static int ks_fused_ex(const Tables &T, size_t L, const KsShard &sh, size_t cs, const u64 *digits, const u64 *dig_ntt,
                       const ckks_ksk *key, const u64 *add0, const u64 *add1, u64 *scratch, u64 *out0t, u64 *out1t, bool mul,
                       size_t j0 = 0, size_t nj = ~(size_t)0) {
    if (nj == ~(size_t)0) nj = L - j0;  // target limbs j0 .. j0+nj-1 (all of them by default)
    if (nj == 0) return CKKS_OK;
    if (key->perm_e != KS_E2 || key->k32 != T.w32) {
        g_err = "gadget key is not in the fused key-switch layout";
        return CKKS_BAD_HANDLE;
    }
    KsArgs a;
    a.digits = digits;
    a.dig_ntt = dig_ntt;
    a.scratch = scratch;
    a.key_b = key->b;
    a.key_a = key->a;
    a.add0 = add0;
    a.add1 = add1;
    a.out0 = out0t;
    a.out1 = out1t;
    a.lc = T.d_lc;
    a.P1 = T.d_P1;
    a.W2 = T.d_W2;
    a.W2i = T.d_W2i;
    a.TTt = T.d_TTt;
    a.TTi = T.d_TTi;
    a.w2_stride = T.w2_stride;
    a.L = (int)L;
    a.Ld = (int)sh.Ld;
    a.joff = sh.joff;
    a.jstep = sh.jstep;
    a.j0 = (int)j0;
    a.dig_ct_stride = sh.dig_ct_stride;
    a.dig_limb_stride = sh.dig_limb_stride;
    a.a1 = T.a1;
    a.a2 = T.a2;
    a.reduce_every = reduce_every_for(T, L);
    a.N = T.n;
    const size_t Ld = sh.Ld;
    const unsigned n1 = 1u << T.a1, n2 = 1u << T.a2;
    cudaStream_t s = S(T);
    dim3 g1(n2, (unsigned)(nj * Ld), (unsigned)cs);  // x: columns; the launcher divides by its column tile
    DISPATCH_A(T.a1, TRY(launch_ks1_a<AA>(T.w32, T.lazy, sh.reduce, mul, g1, s, a)));
    dim3 g2((unsigned)cs, n1 / KS_C2, (unsigned)nj);
    // TMA descriptors: (rho, j2 / gamma, slab) tensors with a [n2][16] box
    KsMaps maps;
    memset(&maps, 0, sizeof(maps));
    bool tma = g_use_tma && n1 >= (unsigned)KS_C2;
    tma = tma && make_tile_map(maps.scratch, scratch, T.w32 ? 4 : 8, n1, n2, cs * L * Ld, KS_C2);
    tma = tma && make_tile_map(maps.key_b, key->b, T.w32 ? 4 : 8, n1, n2, Ld * L, KS_C2);
    tma = tma && make_tile_map(maps.key_a, key->a, T.w32 ? 4 : 8, n1, n2, Ld * L, KS_C2);
    DISPATCH_A(T.a2, TRY(launch_ks2_a<AA>(T.w32, T.lazy, mul, tma, g2, s, a, maps)));
    return CKKS_OK;
}
static int ks_fused(const Tables &T, size_t L, size_t cs, const u64 *digits, const u64 *dig_ntt, const ckks_ksk *key,
                    const u64 *add0, const u64 *add1, u64 *scratch, u64 *out0t, u64 *out1t, bool mul) {
    KsShard sh{L, 0, 1, L * T.n, T.n, T.digit_reduce};
    return ks_fused_ex(T, L, sh, cs, digits, dig_ntt, key, add0, add1, scratch, out0t, out1t, mul);
}

static size_t g_ks_scratch_mib = 8192;  // key-switch scratch per chunk of ciphertexts (tuning knob)
extern "C" int ckks_set_ks_scratch_mib(int mib) {
    if (mib < 1) return CKKS_BAD_ARGUMENT;
    g_ks_scratch_mib = (size_t)mib;
    return CKKS_OK;
}
extern "C" size_t ckks_ks_chunk(const ckks_ctx *ctx, size_t batch);
static size_t ks_chunk(const Tables &T, size_t L, size_t batch) {
    size_t per = L * L * T.n * sizeof(u64);
    size_t c = (g_ks_scratch_mib << 20) / per;
    if (c < 1) c = 1;
    if (c > 32768) c = 32768;  // the ciphertext index is a grid dimension (y/z limit 65535)
    // 32-bit word path: ks_pass2 falls off a cliff beyond 512 ciphertexts per launch (cfg3: 9.97 ms per 3072 rotations
    // with 512 per launch, 14.25 ms with 1024; 12 CTAs per SM all streaming slabs a power-of-two 4 MiB apart)
    if (T.w32 && c > 512) c = 512;
    return c < batch ? c : batch;
}

extern "C" size_t ckks_ks_chunk(const ckks_ctx *ctx, size_t batch) { return ok_ctx(ctx) ? ks_chunk(*ctx->T, ctx->L, batch) : 0; }

#include "aux_ks.inl"

// mul_ciphertexts_gadget (+ rescale_ciphertext) on coefficient-domain device inputs, four-step path.
// o0/o1: [batch][L or L-1][N].
static int fused_mul_relin(const Tables &T, size_t L, size_t batch, const u64 *a0, const u64 *a1, const u64 *b0,
                           const u64 *b1, const ckks_ksk *rlk, bool rescale, u64 *o0, u64 *o1) {
    if (!batch) return CKKS_OK;
    if (rlk->aux_k && aux_wanted(T, L)) return fused_mul_relin_aux(T, L, batch, a0, a1, b0, b1, rlk, rescale, o0, o1);
    NvtxScope nvtx_call(rescale ? "ckks:mul_relin_rescale" : "ckks:mul_relin");
    const size_t n = T.n, cs_max = ks_chunk(T, L, batch);
    const size_t W = cs_max * L * n;
    u64 *A0 = nullptr, *A1 = nullptr, *B0 = nullptr, *B1 = nullptr, *TMP = nullptr, *SCR = nullptr, *LAST = nullptr;
    std::lock_guard<std::mutex> ws_lock(const_cast<Tables &>(T).ws_mu);
    int rc = ws_get(T, WS_A0, W * 8, &A0);
    if (rc == CKKS_OK) rc = ws_get(T, WS_A1, W * 8, &A1);
    if (rc == CKKS_OK) rc = ws_get(T, WS_B0, W * 8, &B0);
    if (rc == CKKS_OK) rc = ws_get(T, WS_B1, W * 8, &B1);
    if (rc == CKKS_OK) rc = ws_get(T, WS_TMP, W * 8, &TMP);
    if (rc == CKKS_OK) rc = ws_get(T, WS_SCR, cs_max * L * L * n * 8, &SCR);
    if (rc == CKKS_OK && rescale) rc = ws_get(T, WS_LAST, 2 * cs_max * n * 8, &LAST);
    const size_t outL = rescale ? L - 1 : L;
    for (size_t s0 = 0; s0 < batch && rc == CKKS_OK; s0 += cs_max) {
        const size_t cs = batch - s0 < cs_max ? batch - s0 : cs_max;
        const size_t off = s0 * L * n;
        Span sp = whole(cs, L);
        auto step = [&]() -> int {
            // forward transforms straight from the caller's (unmodified) inputs
            const u64 *in[4] = {a0 + off, a1 + off, b0 + off, b1 + off};
            u64 *nt[4] = {A0, A1, B0, B1};
            for (int t = 0; t < 4; ++t) {
                TRY(run_pass(T, P_FWD1, sp, in[t], TMP));
                TRY(run_pass(T, P_FWD2, sp, TMP, nt[t]));
            }
            EwArgs e = ew_args(T, L, cs);
            KL("tensor", (tensor_kernel<<<ew_grid(e.total), 256, 0, S(T)>>>(e, A0, A1, B0, B1, A0, A1, B0)));  // d0,d1,d2
            // d2 -> coefficient domain (engine.rs:493); B0 keeps NTT(d2), B1 receives the digits
            TRY(run_pass(T, P_INV2, sp, B0, TMP));
            TRY(run_pass(T, P_INV1, sp, TMP, B1));
            TRY(ks_fused(T, L, cs, B1, B0, rlk, A0, A1, SCR, TMP, B1, true));
            u64 *d0 = o0 + s0 * outL * n, *d1 = o1 + s0 * outL * n;
            if (!rescale) {
                TRY(run_pass(T, P_INV1, sp, TMP, d0));
                TRY(run_pass(T, P_INV1, sp, B1, d1));
                return CKKS_OK;
            }
            // last limb first, then the others with the rescale epilogue (poly.rs:214-225)
            Span last{cs, (int)L, (int)L - 1, 1, 1, (int)L - 1};
            TRY(run_pass(T, P_INV1, last, TMP, LAST));
            TRY(run_pass(T, P_INV1, last, B1, LAST + cs * n));
            PassArgs pa;
            memset(&pa, 0, sizeof(pa));
            pa.lc = T.d_lc;
            pa.tab = T.d_P1i;
            pa.tab_stride = (size_t)1 << T.a1;
            pa.elt = nullptr;
            pa.L = (int)L;
            pa.ncols = 1u << T.a2;
            pa.N = n;
            pa.limb0 = 0;
            pa.dstL = (int)L - 1;
            pa.dst_limb0 = 0;
            dim3 g((1u << T.a2) / 16, (unsigned)(L - 1), (unsigned)cs);
            const void *ql = T.w32 ? (const void *)((const tw32_t *)T.d_qlinv_w + (L - 1) * T.L) : (const void *)(T.d_qlinv + (L - 1) * T.L);
            pa.src = TMP;
            pa.dst = d0;
            DISPATCH_A(T.a1, TRY(launch_inv1_rescale_a<AA>(T.w32, T.lazy, g, S(T), pa, LAST, ql)));
            pa.src = B1;
            pa.dst = d1;
            DISPATCH_A(T.a1, TRY(launch_inv1_rescale_a<AA>(T.w32, T.lazy, g, S(T), pa, LAST + cs * n, ql)));
            return CKKS_OK;
        };
        rc = step();
    }
    return rc;  // the scratch stays cached in the context (ws_get)
}

// Key-switch half of rotate_ciphertext (engine.rs:429-452) on the rotated c1 (coefficient domain):
// ks0/ks1 receive sum_i alpha_i * key_b[i] / key_a[i] in the coefficient domain.
static int fused_keyswitch(const Tables &T, size_t L, size_t batch, const u64 *c1r, const ckks_ksk *key, u64 *ks0, u64 *ks1) {
    if (!batch) return CKKS_OK;
    const size_t n = T.n, cs_max = ks_chunk(T, L, batch);
    u64 *T0 = nullptr, *T1 = nullptr, *SCR = nullptr;
    std::lock_guard<std::mutex> ws_lock(const_cast<Tables &>(T).ws_mu);
    int rc = ws_get(T, WS_A0, cs_max * L * n * 8, &T0);
    if (rc == CKKS_OK) rc = ws_get(T, WS_A1, cs_max * L * n * 8, &T1);
    if (rc == CKKS_OK) rc = ws_get(T, WS_SCR, cs_max * L * L * n * 8, &SCR);
    for (size_t s0 = 0; s0 < batch && rc == CKKS_OK; s0 += cs_max) {
        const size_t cs = batch - s0 < cs_max ? batch - s0 : cs_max;
        const size_t off = s0 * L * n;
        Span sp = whole(cs, L);
        rc = ks_fused(T, L, cs, c1r + off, nullptr, key, nullptr, nullptr, SCR, T0, T1, false);
        if (rc == CKKS_OK) rc = run_pass(T, P_INV1, sp, T0, ks0 + off);
        if (rc == CKKS_OK) rc = run_pass(T, P_INV1, sp, T1, ks1 + off);
    }
    return rc;
}

// Last inverse pass with automorphism(rot_src) added on the fly (ntt_pass_kernel<..., ADDROT>).
template <typename WD, int A>
static int launch_inv1_addrot_w(int lazy, dim3 grid, cudaStream_t s, const PassArgs &a) {
    constexpr int E = 4, C = pass_c<WD>();
    grid.x = a.ncols / C;
    const size_t smem = (size_t)(1 << A) * (C + 1) * sizeof(WD);
    const int block = C << (A - E);
#define M(LZ) \
    do {                                                                                                                                                  \
        if (FIX_MATCH(A, a.N))                                                                                                                            \
            KL("ntt_inv_pass1_addrot", (ntt_pass_kernel<WD, XF_NEG_INV, A, E, C, (sizeof(WD) == 8 ? LZ : (LZ ? 1 : 0)), false, false, false, false, true, FIX_OF(A)><<<grid, block, smem, s>>>(a))); \
        else                                                                                                                                              \
            KL("ntt_inv_pass1_addrot", (ntt_pass_kernel<WD, XF_NEG_INV, A, E, C, (sizeof(WD) == 8 ? LZ : (LZ ? 1 : 0)), false, false, false, false, true><<<grid, block, smem, s>>>(a)));            \
    } while (0)
    LZ_SWITCH(WD, lazy, M);
#undef M
    return CKKS_OK;
}
template <int A>
static int launch_inv1_addrot_a(bool w32, int lazy, dim3 grid, cudaStream_t s, const PassArgs &a) {
    if (w32) return W32_DISPATCH(launch_inv1_addrot_w<u32, A>(lazy, grid, s, a));
    return launch_inv1_addrot_w<u64, A>(lazy, grid, s, a);
}

// rotate_ciphertext (engine.rs:412-463) on coefficient-domain device inputs, four-step path, net Galois exponent
// `e` (odd).  Only the rotated c1 is materialised (it is the digit polynomial); automorphism(c0) is gathered
// inside the last inverse pass of ks0, so c0 is read once and nothing else of the c0 path touches HBM.
static int fused_rotate(const Tables &T, size_t L, size_t batch, const u64 *c0, const u64 *c1, u64 e, const ckks_ksk *key, u64 *o0,
                        u64 *o1) {
    if (!batch) return CKKS_OK;
    if (key->aux_k && aux_wanted(T, L)) return fused_rotate_aux(T, L, batch, c0, c1, e, key, o0, o1);
    NvtxScope nvtx_call("ckks:rotate");
    const size_t n = T.n, cs_max = ks_chunk(T, L, batch);
    const u64 einv = inv_mod_pow2(e, 2 * n);
    u64 *D = nullptr, *T0 = nullptr, *T1 = nullptr, *SCR = nullptr;
    std::lock_guard<std::mutex> ws_lock(const_cast<Tables &>(T).ws_mu);
    int rc = ws_get(T, WS_B0, cs_max * L * n * 8, &D);
    if (rc == CKKS_OK) rc = ws_get(T, WS_A0, cs_max * L * n * 8, &T0);
    if (rc == CKKS_OK) rc = ws_get(T, WS_A1, cs_max * L * n * 8, &T1);
    if (rc == CKKS_OK) rc = ws_get(T, WS_SCR, cs_max * L * L * n * 8, &SCR);
    for (size_t s0 = 0; s0 < batch && rc == CKKS_OK; s0 += cs_max) {
        const size_t cs = batch - s0 < cs_max ? batch - s0 : cs_max;
        const size_t off = s0 * L * n;
        auto step = [&]() -> int {
            EwArgs ea = ew_args(T, L, cs);
            KL("automorphism", (automorphism_kernel<<<ew_grid(ea.total), 256, 0, S(T)>>>(ea, c1 + off, D, e, einv)));
            TRY(ks_fused(T, L, cs, D, nullptr, key, nullptr, nullptr, SCR, T0, T1, false));
            TRY(run_pass(T, P_INV1, whole(cs, L), T1, o1 + off));
            PassArgs a;
            memset(&a, 0, sizeof(a));
            a.lc = T.d_lc;
            a.L = (int)L;
            a.N = n;
            a.dstL = (int)L;
            a.src = T0;
            a.dst = o0 + off;
            a.tab = T.d_P1i;
            a.tab_stride = (size_t)1 << T.a1;
            a.ncols = 1u << T.a2;
            a.rot_src = c0 + off;
            a.rot_einv = einv;
            dim3 g(1, (unsigned)L, (unsigned)cs);
            DISPATCH_A(T.a1, TRY(launch_inv1_addrot_a<AA>(T.w32, T.lazy, g, S(T), a)));
            return CKKS_OK;
        };
        rc = step();
    }
    return rc;
}

static int check_ct(const ckks_poly *c0, const ckks_poly *c1) {
    if (!ok_poly(c0) || !ok_poly(c1)) return CKKS_BAD_HANDLE;
    if (!same_basis(c0->ctx, c1->ctx)) return CKKS_BASIS_MISMATCH;
    if (c0->batch != c1->batch) return CKKS_BATCH_MISMATCH;
    if (c0->ntt != c1->ntt) return CKKS_DOMAIN_MISMATCH;
    return CKKS_OK;
}
static void free2(ckks_poly *a, ckks_poly *b) {
    if (a) ckks_poly_free(a);
    if (b) ckks_poly_free(b);
}

extern "C" int ckks_ct_add(const ckks_poly *a0, const ckks_poly *a1, const ckks_poly *b0, const ckks_poly *b1, ckks_poly **c0,
                           ckks_poly **c1) {
    if (!c0 || !c1) return CKKS_BAD_ARGUMENT;
    *c0 = *c1 = nullptr;
    TRY(check_ct(a0, a1));
    TRY(check_ct(b0, b1));
    TRY(check_pair(a0, b0, false));
    TRY(check_pair(a1, b1, false));
    const Tables &T = *a0->ctx->T;
    CU(cudaSetDevice(T.device));
    ckks_poly *r0 = nullptr, *r1 = nullptr;
    int rc = poly_new(a0->ctx, a0->batch, a0->ntt, &r0);
    if (rc == CKKS_OK) rc = poly_new(a0->ctx, a0->batch, a0->ntt, &r1);
    if (rc == CKKS_OK) {
        EwArgs e = ew_args(T, a0->ctx->L, a0->batch);
        if (e.total) {
            KLV("ew_add3", (ew_add3_kernel<<<ew_grid(e.total), 256, 0, S(T)>>>(e, a0->d, b0->d, r0->d)));
            KLV("ew_add3", (ew_add3_kernel<<<ew_grid(e.total), 256, 0, S(T)>>>(e, a1->d, b1->d, r1->d)));
            if (cudaPeekAtLastError() != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "ew_add3");
        }
    }
    if (rc != CKKS_OK) {
        free2(r0, r1);
        return rc;
    }
    *c0 = r0;
    *c1 = r1;
    return CKKS_OK;
}

// mul_ciphertexts_gadget with the minimal exact schedule: 4L forward transforms of the inputs,
// NTT-domain tensor, L inverse (d2), L(L) digit transforms, NTT-domain accumulation, 2L inverse.
// Every step is exact in Z_q, so the coefficient-domain output words equal the reference's.
// On success *c0,*c1 are in the NTT domain if keep_ntt (internal use by the fused rescale) else coefficient.
static int ct_mul_relin_impl(const ckks_poly *a0, const ckks_poly *a1, const ckks_poly *b0, const ckks_poly *b1,
                             const ckks_ksk *rlk, ckks_poly **c0, ckks_poly **c1) {
    *c0 = *c1 = nullptr;
    TRY(check_ct(a0, a1));
    TRY(check_ct(b0, b1));
    TRY(check_pair(a0, b0, false));
    if (!ok_ksk(rlk)) return CKKS_BAD_HANDLE;
    if (!same_basis(a0->ctx, rlk->ctx)) return CKKS_BASIS_MISMATCH;
    if (a0->ntt) return CKKS_DOMAIN_MISMATCH;  // the engine works on coefficient-domain ciphertexts
    const Tables &T = *a0->ctx->T;
    const size_t L = a0->ctx->L, batch = a0->batch;
    CU(cudaSetDevice(T.device));
    if (T.path == 2 && !g_force_unfused) {
        ckks_poly *r0 = nullptr, *r1 = nullptr;
        int frc = poly_new(a0->ctx, batch, false, &r0);
        if (frc == CKKS_OK) frc = poly_new(a0->ctx, batch, false, &r1);
        if (frc == CKKS_OK) frc = fused_mul_relin(T, L, batch, a0->d, a1->d, b0->d, b1->d, rlk, false, r0->d, r1->d);
        if (frc != CKKS_OK) {
            free2(r0, r1);
            return frc;
        }
        *c0 = r0;
        *c1 = r1;
        return CKKS_OK;
    }
    ckks_poly *A0 = nullptr, *A1 = nullptr, *B0 = nullptr, *B1 = nullptr;
    int rc = ckks_poly_clone(const_cast<ckks_poly *>(a0), &A0);
    if (rc == CKKS_OK) rc = ckks_poly_clone(const_cast<ckks_poly *>(a1), &A1);
    if (rc == CKKS_OK) rc = ckks_poly_clone(const_cast<ckks_poly *>(b0), &B0);
    if (rc == CKKS_OK) rc = ckks_poly_clone(const_cast<ckks_poly *>(b1), &B1);
    if (rc == CKKS_OK) rc = ckks_poly_to_ntt_domain(A0);
    if (rc == CKKS_OK) rc = ckks_poly_to_ntt_domain(A1);
    if (rc == CKKS_OK) rc = ckks_poly_to_ntt_domain(B0);
    if (rc == CKKS_OK) rc = ckks_poly_to_ntt_domain(B1);
    EwArgs e = ew_args(T, L, batch);
    if (rc == CKKS_OK && e.total) {
        // d0 -> A0, d1 -> A1, d2 -> B0
        KLV("tensor", (tensor_kernel<<<ew_grid(e.total), 256, 0, S(T)>>>(e, A0->d, A1->d, B0->d, B1->d, A0->d, A1->d, B0->d)));
        if (cudaPeekAtLastError() != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "tensor");
    }
    if (rc == CKKS_OK) rc = ckks_poly_to_coeff_domain(B0);  // engine.rs:493
    if (rc == CKKS_OK) rc = keyswitch_accumulate(T, L, batch, B0->d, rlk, A0->d, A1->d);
    if (rc == CKKS_OK) rc = ckks_poly_to_coeff_domain(A0);
    if (rc == CKKS_OK) rc = ckks_poly_to_coeff_domain(A1);
    free2(B0, B1);
    if (rc != CKKS_OK) {
        free2(A0, A1);
        return rc;
    }
    *c0 = A0;
    *c1 = A1;
    return CKKS_OK;
}
extern "C" int ckks_ct_mul_relin(const ckks_poly *a0, const ckks_poly *a1, const ckks_poly *b0, const ckks_poly *b1,
                                 const ckks_ksk *rlk, ckks_poly **c0, ckks_poly **c1) {
    if (!c0 || !c1) return CKKS_BAD_ARGUMENT;
    return ct_mul_relin_impl(a0, a1, b0, b1, rlk, c0, c1);
}

static uint32_t bit_length(u64 q) { return 64 - (uint32_t)__builtin_clzll(q); }

extern "C" int ckks_ct_rescale(const ckks_poly *c0, const ckks_poly *c1, ckks_ctx *child, ckks_poly **o0, ckks_poly **o1,
                               uint32_t *bits) {
    if (!o0 || !o1) return CKKS_BAD_ARGUMENT;
    *o0 = *o1 = nullptr;
    TRY(check_ct(c0, c1));
    if (bits) *bits = bit_length(c0->ctx->T->moduli[c0->ctx->L - 1]);  // engine.rs:266-270
    ckks_poly *r0 = nullptr, *r1 = nullptr;
    int rc = ckks_poly_rescale_into(c0, child, &r0);
    if (rc == CKKS_OK) rc = ckks_poly_rescale_into(c1, child, &r1);
    if (rc != CKKS_OK) {
        free2(r0, r1);
        return rc;
    }
    *o0 = r0;
    *o1 = r1;
    return CKKS_OK;
}
extern "C" int ckks_ct_mul_relin_rescale(const ckks_poly *a0, const ckks_poly *a1, const ckks_poly *b0, const ckks_poly *b1,
                                         const ckks_ksk *rlk, ckks_ctx *child, ckks_poly **o0, ckks_poly **o1) {
    if (!o0 || !o1) return CKKS_BAD_ARGUMENT;
    *o0 = *o1 = nullptr;
    if (!ok_poly(a0) || !ok_ctx(child)) return CKKS_BAD_HANDLE;
    if (a0->ctx->L < 2) return CKKS_INVALID_MOD_DROP;
    if (child->T.get() != a0->ctx->T.get() || child->L + 1 != a0->ctx->L) return CKKS_BASIS_MISMATCH;
    if (a0->ctx->T->path == 2 && !g_force_unfused) {
        TRY(check_ct(a0, a1));
        TRY(check_ct(b0, b1));
        TRY(check_pair(a0, b0, false));
        if (!ok_ksk(rlk)) return CKKS_BAD_HANDLE;
        if (!same_basis(a0->ctx, rlk->ctx)) return CKKS_BASIS_MISMATCH;
        if (a0->ntt) return CKKS_DOMAIN_MISMATCH;
        const Tables &T = *a0->ctx->T;
        CU(cudaSetDevice(T.device));
        ckks_poly *r0 = nullptr, *r1 = nullptr;
        int frc = poly_new(child, a0->batch, false, &r0);
        if (frc == CKKS_OK) frc = poly_new(child, a0->batch, false, &r1);
        if (frc == CKKS_OK) frc = fused_mul_relin(T, a0->ctx->L, a0->batch, a0->d, a1->d, b0->d, b1->d, rlk, true, r0->d, r1->d);
        if (frc != CKKS_OK) {
            free2(r0, r1);
            return frc;
        }
        *o0 = r0;
        *o1 = r1;
        return CKKS_OK;
    }
    ckks_poly *m0 = nullptr, *m1 = nullptr;
    TRY(ct_mul_relin_impl(a0, a1, b0, b1, rlk, &m0, &m1));
    int rc = ckks_ct_rescale(m0, m1, child, o0, o1, nullptr);
    free2(m0, m1);
    return rc;
}

extern "C" int ckks_ct_rotate(const ckks_poly *c0, const ckks_poly *c1, const ckks_ksk *rotk, int32_t k, ckks_poly **o0,
                              ckks_poly **o1) {
    if (!o0 || !o1) return CKKS_BAD_ARGUMENT;
    *o0 = *o1 = nullptr;
    TRY(check_ct(c0, c1));
    if (!ok_ksk(rotk)) return CKKS_BAD_HANDLE;
    if (!same_basis(c0->ctx, rotk->ctx)) return CKKS_BASIS_MISMATCH;
    const Tables &T = *c0->ctx->T;
    const size_t L = c0->ctx->L, batch = c0->batch;
    CU(cudaSetDevice(T.device));
    if (T.path == 2 && !g_force_unfused && !c0->ntt) {
        // net exponent of rotate_slots (poly.rs:546-569): 5^|k|, times 2N-1 for k < 0 (two odd automorphisms compose)
        const u64 two_n = 2 * T.n;
        u64 e = rot_exponent(T.n, k);
        if (k < 0) e = (e * (two_n - 1)) % two_n;
        ckks_poly *f0 = nullptr, *f1 = nullptr;
        int frc = poly_new(c0->ctx, batch, false, &f0);
        if (frc == CKKS_OK) frc = poly_new(c0->ctx, batch, false, &f1);
        if (frc == CKKS_OK) frc = fused_rotate(T, L, batch, c0->d, c1->d, e, rotk, f0->d, f1->d);
        if (frc != CKKS_OK) {
            free2(f0, f1);
            return frc;
        }
        *o0 = f0;
        *o1 = f1;
        return CKKS_OK;
    }
    ckks_poly *r0 = nullptr, *r1 = nullptr, *k0 = nullptr, *k1 = nullptr;
    int rc = ckks_poly_rotate_slots(c0, k, &r0);  // engine.rs:417-419
    if (rc == CKKS_OK) rc = ckks_poly_rotate_slots(c1, k, &r1);
    if (rc == CKKS_OK) rc = ckks_poly_to_coeff_domain(r1);
    if (rc == CKKS_OK) rc = ckks_poly_alloc(c0->ctx, batch, &k0);
    if (rc == CKKS_OK) rc = ckks_poly_alloc(c0->ctx, batch, &k1);
    if (rc == CKKS_OK && T.path == 2 && !g_force_unfused) {
        rc = fused_keyswitch(T, L, batch, r1->d, rotk, k0->d, k1->d);
    } else if (rc == CKKS_OK) {
        k0->ntt = k1->ntt = true;  // zero is zero in either domain
        rc = keyswitch_accumulate(T, L, batch, r1->d, rotk, k0->d, k1->d);
    }
    if (rc == CKKS_OK) rc = ckks_poly_to_coeff_domain(k0);
    if (rc == CKKS_OK) rc = ckks_poly_to_coeff_domain(k1);
    if (rc == CKKS_OK) rc = ckks_poly_to_coeff_domain(r0);
    if (rc == CKKS_OK) rc = ckks_poly_add_assign(r0, k0);  // engine.rs:454-455
    free2(r1, k0);
    if (rc != CKKS_OK) {
        free2(r0, k1);
        return rc;
    }
    *o0 = r0;
    *o1 = k1;
    return CKKS_OK;
}

extern "C" int ckks_ct_encrypt(const ckks_poly *pk_b, const ckks_poly *pk_a, const ckks_poly *u, const ckks_poly *e0,
                               const ckks_poly *e1, const ckks_poly *m, ckks_poly **c0, ckks_poly **c1) {
    if (!c0 || !c1) return CKKS_BAD_ARGUMENT;
    *c0 = *c1 = nullptr;
    if (!ok_poly(pk_b) || !ok_poly(pk_a) || !ok_poly(u) || !ok_poly(e0) || !ok_poly(e1) || !ok_poly(m)) return CKKS_BAD_HANDLE;
    // engine.rs:96-104: c0 = pk.b.clone(); c0 *= u; c0 += e0; c0 += m;  c1 = pk.a.clone(); c1 *= u; c1 += e1.
    // Each `*=` of coefficient-domain operands is forward(both), pointwise, inverse (poly.rs:307-329); everything is
    // exact in Z_q, so u is transformed ONCE for both products (1 batched forward + 2 inverse transforms instead of
    // 2 + 2, plus the two transforms of the public key, which has batch 1) and the words are the reference's.
    TRY(check_pair(u, pk_b, true));
    TRY(check_pair(u, pk_a, true));
    if (u->ntt) return CKKS_DOMAIN_MISMATCH;
    ckks_poly *r0 = nullptr, *r1 = nullptr, *pb = nullptr, *pa = nullptr;
    int rc = ckks_poly_clone(const_cast<ckks_poly *>(u), &r1);
    if (rc == CKKS_OK) rc = ckks_poly_to_ntt_domain(r1);
    if (rc == CKKS_OK) rc = ckks_poly_clone(r1, &r0);
    if (rc == CKKS_OK) rc = ckks_poly_clone(const_cast<ckks_poly *>(pk_b), &pb);
    if (rc == CKKS_OK) rc = ckks_poly_to_ntt_domain(pb);
    if (rc == CKKS_OK) rc = ckks_poly_clone(const_cast<ckks_poly *>(pk_a), &pa);
    if (rc == CKKS_OK) rc = ckks_poly_to_ntt_domain(pa);
    if (rc == CKKS_OK) rc = ckks_poly_mul_assign(r0, pb);  // NTT domain: pointwise (poly.rs:297-306); pk broadcasts over the batch
    if (rc == CKKS_OK) rc = ckks_poly_mul_assign(r1, pa);
    if (rc == CKKS_OK) rc = ckks_poly_to_coeff_domain(r0);
    if (rc == CKKS_OK) rc = ckks_poly_to_coeff_domain(r1);
    if (rc == CKKS_OK) rc = ckks_poly_add_assign(r0, e0);
    if (rc == CKKS_OK) rc = ckks_poly_add_assign(r0, m);
    if (rc == CKKS_OK) rc = ckks_poly_add_assign(r1, e1);
    free2(pb, pa);
    if (rc != CKKS_OK) {
        free2(r0, r1);
        return rc;
    }
    *c0 = r0;
    *c1 = r1;
    return CKKS_OK;
}
extern "C" int ckks_ct_decrypt(const ckks_poly *c0, const ckks_poly *c1, const ckks_poly *s, ckks_poly **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    TRY(check_ct(c0, c1));
    if (!ok_poly(s)) return CKKS_BAD_HANDLE;
    ckks_poly *r = nullptr;  // engine.rs:121-124: c1 * s + c0
    int rc = ckks_poly_clone(const_cast<ckks_poly *>(c1), &r);
    if (rc == CKKS_OK) rc = ckks_poly_mul_assign(r, s);
    if (rc == CKKS_OK) rc = ckks_poly_add_assign(r, c0);
    if (rc != CKKS_OK) {
        if (r) ckks_poly_free(r);
        return rc;
    }
    *out = r;
    return CKKS_OK;
}

// -------------------------------------------------------------------------------------------------
// host-buffer entry points
// -------------------------------------------------------------------------------------------------
extern "C" int ckks_host_alloc(size_t bytes, void **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    CU(cudaMallocHost(out, bytes ? bytes : 1));
    return CKKS_OK;
}
extern "C" int ckks_host_free(void *p) {
    if (p) CU(cudaFreeHost(p));
    return CKKS_OK;
}

// Three-stage pipeline over chunks of the batch: H2D on one copy stream, the fused kernels on the
// context's stream, D2H on a second copy stream, double-buffered, ordered with events.  Host buffers
// should be page-locked (ckks_host_alloc) for the copies to overlap; pageable memory still works.
struct HostPipe {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t in_done[2] = {nullptr, nullptr}, comp_done[2] = {nullptr, nullptr}, out_done[2] = {nullptr, nullptr};
    u64 *in[2][4] = {{nullptr}}, *out[2][2] = {{nullptr}};
    int *flag = nullptr;  // set by the reducedness scan of the staged inputs (poly.rs:83-93)
    int init(const Tables &T, int n_in, size_t in_words, size_t out_words) {
        CU(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            CU(cudaEventCreateWithFlags(&in_done[b], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&comp_done[b], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&out_done[b], cudaEventDisableTiming));
            for (int t = 0; t < n_in; ++t) CU(cudaMalloc((void **)&in[b][t], in_words * 8));
            for (int t = 0; t < 2; ++t) CU(cudaMalloc((void **)&out[b][t], out_words * 8));
        }
        CU(cudaMalloc((void **)&flag, sizeof(int)));
        (void)T;
        return CKKS_OK;
    }
    void destroy() {
        for (int b = 0; b < 2; ++b) {
            for (int t = 0; t < 4; ++t)
                if (in[b][t]) cudaFree(in[b][t]);
            for (int t = 0; t < 2; ++t)
                if (out[b][t]) cudaFree(out[b][t]);
            if (in_done[b]) cudaEventDestroy(in_done[b]);
            if (comp_done[b]) cudaEventDestroy(comp_done[b]);
            if (out_done[b]) cudaEventDestroy(out_done[b]);
        }
        if (flag) cudaFree(flag);
        if (s_in) cudaStreamDestroy(s_in);
        if (s_out) cudaStreamDestroy(s_out);
    }
};

static size_t host_chunk(const Tables &T, size_t L, size_t batch) {
    // about g_host_chunk_mib MiB per component and chunk: small enough to pipeline, large enough to fill the GPU
    size_t per = L * T.n * sizeof(u64);
    size_t c = ((size_t)g_host_chunk_mib << 20) / per;
    if (c < 1) c = 1;
    return c < batch ? c : batch;
}

// kind 0: mul_ciphertexts_gadget + rescale_ciphertext (4 inputs, L-1 output limbs); kind 1: rotate (2 inputs).
static int host_pipeline(ckks_ctx *ctx, int kind, const ckks_ksk *key, int32_t rot, size_t batch, const u64 *const *hin,
                         u64 *const *hout) {
    const Tables &T = *ctx->T;
    CU(cudaSetDevice(T.device));
    if (!batch) return CKKS_OK;
    const size_t L = ctx->L, n = T.n;
    const int n_in = kind == 0 ? 4 : 2;
    const size_t outL = kind == 0 ? L - 1 : L;
    const size_t wi = L * n, wo = outL * n;
    const size_t chunk = host_chunk(T, L, batch);
    if (T.path != 2) {  // small-N path: plain sequential staging through device polynomials
        for (size_t s = 0; s < batch; s += chunk) {
            size_t nb = batch - s < chunk ? batch - s : chunk;
            ckks_poly *P[4] = {nullptr, nullptr, nullptr, nullptr}, *R0 = nullptr, *R1 = nullptr;
            int rc = CKKS_OK;
            for (int t = 0; t < n_in && rc == CKKS_OK; ++t)  // from_channels: reducedness scan included (poly.rs:83-93)
                rc = ckks_poly_from_channels(ctx, nb, (const uint64_t *)(hin[t] + s * wi), L, 0, &P[t]);
            ckks_ctx *child = nullptr;
            if (rc == CKKS_OK && kind == 0) rc = ckks_ctx_drop_last(ctx, 1, &child);
            if (rc == CKKS_OK) rc = kind == 0 ? ckks_ct_mul_relin_rescale(P[0], P[1], P[2], P[3], key, child, &R0, &R1) : ckks_ct_rotate(P[0], P[1], key, rot, &R0, &R1);
            if (rc == CKKS_OK && (cudaMemcpyAsync(hout[0] + s * wo, R0->d, nb * wo * 8, cudaMemcpyDeviceToHost, S(T)) != cudaSuccess ||
                                  cudaMemcpyAsync(hout[1] + s * wo, R1->d, nb * wo * 8, cudaMemcpyDeviceToHost, S(T)) != cudaSuccess))
                rc = cuda_fail(cudaGetLastError(), "d2h");
            if (cudaStreamSynchronize(S(T)) != cudaSuccess && rc == CKKS_OK) rc = cuda_fail(cudaGetLastError(), "sync");
            for (int t = 0; t < 4; ++t)
                if (P[t]) ckks_poly_free(P[t]);
            free2(R0, R1);
            if (child) ckks_ctx_destroy(child);
            TRY(rc);
        }
        return CKKS_OK;
    }
    Tables &TM = *ctx->T;
    std::lock_guard<std::mutex> pipe_lock(TM.pipe_mu);
    int rc = CKKS_OK;
    if (!TM.pipe || TM.pipe_nin < n_in || TM.pipe_in_words < chunk * wi || TM.pipe_out_words < chunk * wo) {
        destroy_host_pipe(TM.pipe);
        TM.pipe = new HostPipe();
        rc = TM.pipe->init(T, n_in, chunk * wi, chunk * wo);
        TM.pipe_nin = n_in;
        TM.pipe_in_words = chunk * wi;
        TM.pipe_out_words = chunk * wo;
    }
    HostPipe &hp = *TM.pipe;
    if (rc == CKKS_OK && cudaMemsetAsync(hp.flag, 0, sizeof(int), S(T)) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "memset");
    const u64 e1 = rot >= 0 ? rot_exponent(n, rot) : (rot_exponent(n, rot) * (2 * n - 1)) % (2 * n);
    size_t c = 0;
    for (size_t s = 0; s < batch && rc == CKKS_OK; s += chunk, ++c) {
        const size_t nb = batch - s < chunk ? batch - s : chunk;
        const int b = (int)(c & 1);
        auto step = [&]() -> int {
            if (c >= 2) CU(cudaStreamWaitEvent(hp.s_in, hp.comp_done[b], 0));  // compute of chunk c-2 has released in[b]
            for (int t = 0; t < n_in; ++t)
                CU(cudaMemcpyAsync(hp.in[b][t], hin[t] + s * wi, nb * wi * 8, cudaMemcpyHostToDevice, hp.s_in));
            CU(cudaEventRecord(hp.in_done[b], hp.s_in));
            CU(cudaStreamWaitEvent(S(T), hp.in_done[b], 0));
            if (c >= 2) CU(cudaStreamWaitEvent(S(T), hp.out_done[b], 0));  // D2H of chunk c-2 has drained out[b]
            // what from_channels checks on every polynomial the reference builds from raw words (poly.rs:83-93): the
            // lazy butterflies assume canonical inputs, so a word >= q must be an error, not silent garbage
            for (int t = 0; t < n_in; ++t) TRY(scan_reduced(T, L, nb, hp.in[b][t], hp.flag, S(T)));
            if (kind == 0) {
                TRY(fused_mul_relin(T, L, nb, hp.in[b][0], hp.in[b][1], hp.in[b][2], hp.in[b][3], key, true, hp.out[b][0], hp.out[b][1]));
            } else {
                // rotate_ciphertext (engine.rs:412-463): key-switch of the rotated c1, automorphism(c0) gathered in
                // the last inverse pass
                TRY(fused_rotate(T, L, nb, hp.in[b][0], hp.in[b][1], e1, key, hp.out[b][0], hp.out[b][1]));
            }
            CU(cudaEventRecord(hp.comp_done[b], S(T)));
            CU(cudaStreamWaitEvent(hp.s_out, hp.comp_done[b], 0));
            CU(cudaMemcpyAsync(hout[0] + s * wo, hp.out[b][0], nb * wo * 8, cudaMemcpyDeviceToHost, hp.s_out));
            CU(cudaMemcpyAsync(hout[1] + s * wo, hp.out[b][1], nb * wo * 8, cudaMemcpyDeviceToHost, hp.s_out));
            CU(cudaEventRecord(hp.out_done[b], hp.s_out));
            return CKKS_OK;
        };
        rc = step();
    }
    int hflag = 0;
    if (rc == CKKS_OK && cudaMemcpyAsync(&hflag, hp.flag, sizeof(int), cudaMemcpyDeviceToHost, S(T)) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "flag");
    cudaStreamSynchronize(hp.s_in);
    if (cudaStreamSynchronize(S(T)) != cudaSuccess && rc == CKKS_OK) rc = cuda_fail(cudaGetLastError(), "sync");
    if (cudaStreamSynchronize(hp.s_out) != cudaSuccess && rc == CKKS_OK) rc = cuda_fail(cudaGetLastError(), "sync");
    if (rc == CKKS_OK && hflag) return CKKS_NON_REDUCED_COEFFICIENT;  // outputs are unspecified in that case
    if (rc != CKKS_OK) {  // do not keep a pipeline whose events may be in an unknown state
        destroy_host_pipe(TM.pipe);
        TM.pipe = nullptr;
    }
    return rc;
}
static void destroy_host_pipe(HostPipe *p) {
    if (!p) return;
    p->destroy();
    delete p;
}

extern "C" int ckks_ct_mul_relin_rescale_host(ckks_ctx *ctx, ckks_ctx *child, const ckks_ksk *rlk, size_t batch,
                                              const uint64_t *a0, const uint64_t *a1, const uint64_t *b0, const uint64_t *b1,
                                              uint64_t *o0, uint64_t *o1) {
    if (!ok_ctx(ctx) || !ok_ctx(child) || !ok_ksk(rlk)) return CKKS_BAD_HANDLE;
    if (batch && (!a0 || !a1 || !b0 || !b1 || !o0 || !o1)) return CKKS_BAD_ARGUMENT;
    if (ctx->L < 2) return CKKS_INVALID_MOD_DROP;
    if (child->T.get() != ctx->T.get() || child->L + 1 != ctx->L) return CKKS_BASIS_MISMATCH;
    if (!same_basis(ctx, rlk->ctx)) return CKKS_BASIS_MISMATCH;
    const u64 *hin[4] = {(const u64 *)a0, (const u64 *)a1, (const u64 *)b0, (const u64 *)b1};
    u64 *hout[2] = {(u64 *)o0, (u64 *)o1};
    return host_pipeline(ctx, 0, rlk, 0, batch, hin, hout);
}

extern "C" int ckks_ct_rotate_host(ckks_ctx *ctx, const ckks_ksk *rotk, int32_t k, size_t batch, const uint64_t *c0,
                                   const uint64_t *c1, uint64_t *o0, uint64_t *o1) {
    if (!ok_ctx(ctx) || !ok_ksk(rotk)) return CKKS_BAD_HANDLE;
    if (batch && (!c0 || !c1 || !o0 || !o1)) return CKKS_BAD_ARGUMENT;
    if (!same_basis(ctx, rotk->ctx)) return CKKS_BASIS_MISMATCH;
    const u64 *hin[4] = {(const u64 *)c0, (const u64 *)c1, nullptr, nullptr};
    u64 *hout[2] = {(u64 *)o0, (u64 *)o1};
    return host_pipeline(ctx, 1, rotk, k, batch, hin, hout);
}

// -------------------------------------------------------------------------------------------------
// integer-pipe peak
// -------------------------------------------------------------------------------------------------
extern "C" double ckks_bench_modmul_peak(int device, int iters) {
    if (ckks_device_count() <= device) return 0.0;
    if (cudaSetDevice(device) != cudaSuccess) return 0.0;
    u64 *d;
    if (cudaMalloc((void **)&d, 64) != cudaSuccess) return 0.0;
    const u64 q = 2305843009211596801ull;  // a 61-bit NTT prime
    tw_t t = mk_tw(1234567890123456789ull % q, q);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 256;
    modmul_peak_kernel<<<blocks, threads>>>(d, 16, q, t);
    cudaEventRecord(e0);
    KLV("modmul_peak", (modmul_peak_kernel<<<blocks, threads>>>(d, iters, q, t)));
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (cudaGetLastError() != cudaSuccess || ms <= 0) return 0.0;
    return (double)blocks * threads * 8.0 * iters / (ms * 1e-3);
}

extern "C" double ckks_bench_mac32_peak(int device, int iters) { return ckks_bench_mac32_peak_ex(device, iters, 1); }
extern "C" double ckks_bench_mac32_peak_ex(int device, int iters, int vary) {
    if (ckks_device_count() <= device) return 0.0;
    if (cudaSetDevice(device) != cudaSuccess) return 0.0;
    u64 *d;
    if (cudaMalloc((void **)&d, 64) != cudaSuccess) return 0.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 256;
    if (vary) mac32_peak_kernel<true><<<blocks, threads>>>(d, 16, 0x3ffffff1u);
    else mac32_peak_kernel<false><<<blocks, threads>>>(d, 16, 0x3ffffff1u);
    cudaEventRecord(e0);
    if (vary) KLV("mac32_peak", (mac32_peak_kernel<true><<<blocks, threads>>>(d, iters, 0x3ffffff1u)));
    else KLV("mac32_peak", (mac32_peak_kernel<false><<<blocks, threads>>>(d, iters, 0x3ffffff1u)));
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (cudaGetLastError() != cudaSuccess || ms <= 0) return 0.0;
    return (double)blocks * threads * 8.0 * iters / (ms * 1e-3);
}

// -------------------------------------------------------------------------------------------------
// per-kernel timing, device-pointer interop
// -------------------------------------------------------------------------------------------------
extern "C" int ckks_prof_enable(int on) {
    g_prof = on != 0;
    return CKKS_OK;
}
extern "C" size_t ckks_prof_collect(char *buf, size_t cap) {
    std::map<std::string, std::pair<uint64_t, double>> agg;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto &r : g_prof_recs) {
        cudaEventSynchronize(r.e1);
        float ms = 0;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
            auto &a = agg[r.name];
            a.first++;
            a.second += ms;
        }
        g_prof_pool.push_back(r.e0);
        g_prof_pool.push_back(r.e1);
    }
    g_prof_recs.clear();
    cudaGetLastError();
    std::string s;
    char line[256];
    for (auto &kv : agg) {
        snprintf(line, sizeof line, "%s=%llu,%.6f\n", kv.first.c_str(), (unsigned long long)kv.second.first, kv.second.second);
        s += line;
    }
    if (buf && cap) {
        size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
        memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return s.size() + 1;
}
extern "C" int ckks_poly_from_device(ckks_ctx *ctx, size_t batch, const uint64_t *dev, int in_ntt, ckks_poly **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    if (!ok_ctx(ctx)) return CKKS_BAD_HANDLE;
    if (!dev && batch) return CKKS_BAD_ARGUMENT;
    const Tables &T = *ctx->T;
    CU(cudaSetDevice(T.device));
    ckks_poly *p;
    TRY(poly_new(ctx, batch, in_ntt != 0, &p));
    size_t words = poly_words(p);
    int rc = CKKS_OK;
    if (words) {
        int *flag = nullptr, hflag = 0;
        if (pool_malloc(T, (void **)&flag, sizeof(int)) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "malloc");
        if (rc == CKKS_OK) {
            cudaMemsetAsync(flag, 0, sizeof(int), S(T));
            EwArgs a = ew_args(T, ctx->L, batch);
            KLV("check_reduced", (check_reduced_kernel<<<ew_grid(a.total), 256, 0, S(T)>>>(a, (const u64 *)dev, flag)));
            if (in_ntt) rc = permute(T, (const u64 *)dev, p->d, words, true);
            else if (cudaMemcpyAsync(p->d, dev, words * 8, cudaMemcpyDeviceToDevice, S(T)) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "d2d");
            cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, S(T));
            if (cudaStreamSynchronize(S(T)) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "sync");
            if (rc == CKKS_OK && hflag) rc = CKKS_NON_REDUCED_COEFFICIENT;
        }
        dev_free(T, flag);
    }
    if (rc != CKKS_OK) {
        ckks_poly_free(p);
        return rc;
    }
    *out = p;
    return CKKS_OK;
}
extern "C" int ckks_poly_device_ptr(ckks_poly *p, uint64_t **out) {
    if (!ok_poly(p) || !out) return CKKS_BAD_HANDLE;
    *out = (uint64_t *)p->d;
    return CKKS_OK;
}

// -------------------------------------------------------------------------------------------------
// CkksEncoder on the device (ckks_encoder.rs:65-156, special_fft.rs:194-242): O(N log N), f64
// -------------------------------------------------------------------------------------------------
static int fft_run(const Tables &T, cplx *buf, size_t batch, double sign) {
    const size_t work = batch * (T.n / 2);
    for (int lh = 0; lh < T.logn; ++lh)
        KL("encoder_fft_stage", (fft_stage_kernel<<<(unsigned)((work + 255) / 256), 256, 0, S(T)>>>(buf, T.logn, lh, sign, batch)));
    return CKKS_OK;
}
static int pow5_table(const Tables &T, unsigned **out) {
    std::vector<unsigned> h(T.n / 2 ? T.n / 2 : 1);
    u64 v = 1;
    for (size_t i = 0; i < h.size(); ++i) {
        h[i] = (unsigned)v;
        v = (v * 5) % (2 * T.n);
    }
    CU(pool_malloc(T, (void **)out, h.size() * sizeof(unsigned)));
    CU(cudaMemcpyAsync(*out, h.data(), h.size() * sizeof(unsigned), cudaMemcpyHostToDevice, S(T)));
    CU(cudaStreamSynchronize(S(T)));  // `h` goes out of scope
    return CKKS_OK;
}
extern "C" int ckks_encode(ckks_ctx *ctx, uint32_t scale_bits, size_t batch, const double *values, size_t nvals, ckks_poly **out) {
    if (!out) return CKKS_BAD_ARGUMENT;
    *out = nullptr;
    if (!ok_ctx(ctx)) return CKKS_BAD_HANDLE;
    const Tables &T = *ctx->T;
    if (T.n < 2 || scale_bits == 0) return CKKS_BAD_ARGUMENT;  // ckks_encoder.rs:38-45
    if (nvals > T.n / 2) return CKKS_SHORT_INPUT;                // ckks_encoder.rs:70-75: more values than slots
    if (!values && batch && nvals) return CKKS_BAD_ARGUMENT;
    CU(cudaSetDevice(T.device));
    TRY(poly_new(ctx, batch, false, out));
    if (!batch) return CKKS_OK;
    cplx *dv = nullptr, *buf = nullptr;
    long long *dc = nullptr;
    unsigned *p5 = nullptr;
    auto body = [&]() -> int {
        TRY(pow5_table(T, &p5));
        CU(pool_malloc(T, (void **)&dv, (batch * nvals + 1) * sizeof(cplx)));
        CU(pool_malloc(T, (void **)&buf, batch * T.n * sizeof(cplx)));
        CU(pool_malloc(T, (void **)&dc, batch * T.n * sizeof(long long)));
        if (nvals) CU(cudaMemcpyAsync(dv, values, batch * nvals * sizeof(cplx), cudaMemcpyHostToDevice, S(T)));
        const size_t half = batch * (T.n / 2);
        KL("encoder_scatter", (enc_scatter_kernel<<<(unsigned)((half + 255) / 256), 256, 0, S(T)>>>(dv, nvals, ldexp(1.0, (int)scale_bits), p5, buf,
                                                                                                         T.logn, batch)));
        TRY(fft_run(T, buf, batch, +1.0));
        const size_t tot = batch * T.n;
        KL("encoder_finish", (enc_finish_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, S(T)>>>(buf, dc, T.logn, batch)));
        EwArgs a = ew_args(T, ctx->L, batch);
        KL("from_coeffs", (from_coeffs_kernel<<<ew_grid(a.total), 256, 0, S(T)>>>(a, dc, T.n, (*out)->d)));
        CU(cudaStreamSynchronize(S(T)));  // `values` may be pageable and reused by the caller
        return CKKS_OK;
    };
    int rc = body();
    dev_free(T, dv);
    dev_free(T, buf);
    dev_free(T, dc);
    dev_free(T, p5);
    if (rc != CKKS_OK) {
        ckks_poly_free(*out);
        *out = nullptr;
    }
    return rc;
}
extern "C" int ckks_decode(const ckks_poly *p, uint32_t scale_bits, size_t nslots, double *out) {
    if (!ok_poly(p)) return CKKS_BAD_HANDLE;
    const Tables &T = *p->ctx->T;
    if (T.n < 2 || nslots > T.n / 2) return CKKS_BAD_ARGUMENT;
    if (!p->batch || !nslots) return CKKS_OK;
    if (!out) return CKKS_BAD_ARGUMENT;
    CU(cudaSetDevice(T.device));
    // Q < 2^128: centred CRT exactly as the reference (basis.rs:158-180: u128 arithmetic on the host, `as i64`), then the
    // transform on the device.  Q >= 2^128 (where the reference's u128 product overflows and it cannot decode at all):
    // Garner's mixed-radix CRT on the device (crt_wide.cuh), coefficients handed to the transform as doubles.
    bool fits = true;
    {
        hm::u128 q = 1;
        for (size_t i = 0; i < p->ctx->L && fits; ++i) {
            if (q > (~(hm::u128)0) / T.moduli[i]) fits = false;
            else q *= T.moduli[i];
        }
    }
    std::vector<int64_t> co;
    if (fits) {
        co.resize(p->batch * T.n);
        TRY(ckks_poly_to_coeffs(p, co.data()));
    }
    cplx *buf = nullptr, *dout = nullptr;
    long long *dc = nullptr;
    double *dcf = nullptr;
    unsigned *p5 = nullptr;
    const size_t batch = p->batch;
    auto body = [&]() -> int {
        TRY(pow5_table(T, &p5));
        CU(pool_malloc(T, (void **)&buf, batch * T.n * sizeof(cplx)));
        CU(pool_malloc(T, (void **)&dout, batch * nslots * sizeof(cplx)));
        const size_t tot = batch * T.n;
        if (fits) {
            CU(pool_malloc(T, (void **)&dc, batch * T.n * sizeof(long long)));
            CU(cudaMemcpyAsync(dc, co.data(), batch * T.n * sizeof(long long), cudaMemcpyHostToDevice, S(T)));
            KL("decoder_twist", (dec_twist_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, S(T)>>>(dc, buf, T.logn, batch)));
        } else {
            CU(pool_malloc(T, (void **)&dcf, batch * T.n * sizeof(double)));
            TRY(crt_wide_dev(p, nullptr, dcf, nullptr));
            KL("decoder_twist", (dec_twist_f64_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, S(T)>>>(dcf, buf, T.logn, batch)));
        }
        TRY(fft_run(T, buf, batch, -1.0));
        const size_t ns = batch * nslots;
        KL("decoder_gather", (dec_gather_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, S(T)>>>(buf, p5, ldexp(1.0, -(int)scale_bits), dout, nslots,
                                                                                                     T.logn, batch)));
        CU(cudaMemcpyAsync(out, dout, ns * sizeof(cplx), cudaMemcpyDeviceToHost, S(T)));
        CU(cudaStreamSynchronize(S(T)));
        return CKKS_OK;
    };
    int rc = body();
    dev_free(T, buf);
    dev_free(T, dc);
    dev_free(T, dcf);
    dev_free(T, dout);
    dev_free(T, p5);
    return rc;
}

#include "batch_shard.inl"
#include "limb_shard.inl"
