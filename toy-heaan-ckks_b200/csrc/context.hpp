// context.hpp -- device-side state behind the opaque handles of include/ckks_b200.h.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <map>
#include <memory>
#include <unordered_map>
#include <vector>
#include <mutex>
#include <vector>

#include "arena.hpp"
#include "modarith.cuh"

// Everything derived from (N, moduli): the device image of RnsBasis<N> + NttTable<N>
// (basis.rs:6-17, 91-94).  Shared by a context and all contexts obtained from it with drop_last.
struct Tables {
    int device = 0;
    u64 n = 0;
    int logn = 0;
    int path = 0;  // 1 = small single-CTA NTT, 2 = four-step
    int a1 = 0, a2 = 0;  // four-step: N = 2^a1 * 2^a2 (small path: a1 = logn, a2 = 0)
    int lazy = 1;        // 0 strict; 1 Harvey lazy (q < 2^62, or < 2^30 with u32 words); 2 lazy8 (u64 words, q < 2^61)
    bool w32 = false;    // all q < 2^31 on the four-step path: 32-bit butterflies, tables and scratch
    bool digit_reduce = true;  // key-switch digits need `% q_j` before the lazy NTT
    size_t L = 0;
    std::vector<u64> moduli, psi;
    LimbConst *d_lc = nullptr;
    // small path
    tw_t *d_psi = nullptr, *d_psi_inv = nullptr, *d_ninv = nullptr;
    // four-step
    // (tw_t entries, or tw32_t when w32)
    void *d_P1 = nullptr, *d_P1i = nullptr, *d_W2 = nullptr, *d_W2i = nullptr, *d_TT = nullptr, *d_TTi = nullptr, *d_TTt = nullptr;
    void *d_qlinv_w = nullptr;  // qlinv in the transform word type (tw32_t when w32, else aliases d_qlinv)
    size_t w2_stride = 1;
    // rescale: qlinv[last][i] = q_last^-1 mod q_i (Shoup pair), [L][L]
    tw_t *d_qlinv = nullptr;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaMemPool_t pool = nullptr;  // private pool of the library's stream-ordered allocations
    // limb-sharded mode: host-visible word a barrier sets when it gives up waiting for a peer GPU; once non-zero the
    // blocking calls on this context (download, sync) report CKKS_NCCL_ERROR instead of handing out poisoned words
    const volatile unsigned *fail_word = nullptr;
    // cached staging buffers / copy streams of the host-buffer entry points
    // Arena on top of the pool (arena.hpp; pool_malloc / dev_free): requests of 1 MiB and more are carved out of
    // large segments taken from the pool once, so a loop of calls at changing levels never goes back to the driver.
    std::mutex cache_mu;
    Arena arena;
    // persistent scratch of the fused pipelines (ws_get): slot buffers grow to the largest request and are kept, so
    // a chain of calls at changing levels never goes back to the allocator; ws_mu serialises the calls that use them
    struct WsSlot {
        void *p = nullptr;
        size_t bytes = 0;
    } ws[8];
    std::mutex ws_mu;
    struct AuxKs *aux = nullptr;  // auxiliary-basis key-switch tables (aux_ks.inl), built on first use
    std::mutex aux_mu;
    std::mutex pipe_mu;  // the *_host entry points of one context tree take turns on the staging pipeline
    struct HostPipe *pipe = nullptr;
    int pipe_nin = 0;
    size_t pipe_in_words = 0, pipe_out_words = 0;
    ~Tables();
};

struct ckks_ctx {
    uint32_t magic;
    std::shared_ptr<Tables> T;
    size_t L;  // limbs visible through this context (prefix of T->moduli)
    std::atomic<int> refs;
};
struct ckks_poly {
    uint32_t magic;
    ckks_ctx *ctx;
    size_t batch;
    u64 *d;  // [batch][L][N]
    bool ntt;
};
struct ckks_ksk {
    uint32_t magic;
    ckks_ctx *ctx;
    u64 *a, *b;  // [digit][limb][N], NTT domain, device-internal order
    size_t digits;  // == ctx->L, except for a limb-sharded key slice (all digits x this GPU's limbs)
    int perm_e;     // >= 0: rows of every limb stored permuted for ks_pass2 (kernels.cuh perm_row), -1: natural
    bool k32;       // words are u32 (32-bit word path: half the key bytes in HBM, L2 and shared memory)
    // auxiliary-basis form (aux_ks.cuh): NTT_{p_k}(key[i][j] mod p_k), [k][j][i][N] u32; aux_k = 0: not present
    u32 *xa, *xb;
    int aux_k;
};

enum : uint32_t { MAGIC_CTX = 0x434b4358u, MAGIC_POLY = 0x434b504cu, MAGIC_KSK = 0x434b4b53u, MAGIC_LSHARD = 0x434b4c53u };
