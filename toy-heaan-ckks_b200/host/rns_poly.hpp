// rns_poly.hpp -- C++ host-side mirror of the reference's backend interface over the C ABI
// (include/ckks_b200.h).  Header-only, RAII, no CUDA or torch types: link with -lckks_b200.
//
// The reference is a Rust crate; with no Rust toolchain in the build image the compiled-language
// mirror is C++.  Names, argument meaning and error behaviour follow the reference:
//   ckks::RnsBasis      Arc<RnsBasis<N>>                    basis.rs:91-181
//   ckks::RnsPoly       RnsPoly<N> (a batch of them)        poly.rs:26-570  (+=, *=, unary -, clone)
//   ckks::RnsNttError   RnsNttError                         errors.rs:3-22
//   ckks::Ciphertext    Ciphertext<RnsPoly<N>, N>           types.rs:22-35
//   ckks::GadgetKey     RnsGadgetRelinKey / RotationKey     engine.rs:225-253
//   ckks::engine::*     CkksEngine<RnsPoly<N>, N>::*        engine.rs:84-151, 263-539
// Where Rust returns Result<_, RnsNttError> or panics, this mirror throws ckks::RnsNttError.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/ckks_b200.h"

namespace ckks {

struct RnsNttError : std::runtime_error {
    int code;
    explicit RnsNttError(int c) : std::runtime_error(std::string(ckks_status_str(c)) + (c == CKKS_CUDA_ERROR ? std::string(": ") + ckks_last_error() : "")), code(c) {}
};
inline void check(int rc) {
    if (rc != CKKS_OK) throw RnsNttError(rc);
}

inline std::vector<uint64_t> generate_primes(int bit_size, int count, uint64_t degree) {  // utils.rs:47-80
    std::vector<uint64_t> out(count > 0 ? count : 1);
    check(ckks_generate_primes(bit_size, count, degree, out.data()));
    out.resize(count);
    return out;
}

class RnsBasis {
    struct Del {
        void operator()(ckks_ctx *c) const { ckks_ctx_destroy(c); }
    };
    std::shared_ptr<ckks_ctx> h_;
    explicit RnsBasis(ckks_ctx *c) : h_(c, Del()) {}
    RnsBasis(ckks_ctx *c, std::shared_ptr<void> owner) : h_(std::move(owner), c) {}  // borrowed (kept alive by `owner`)
    friend class LimbShard;

  public:
    RnsBasis() = default;
    // RnsBasis::new(moduli) -> Result<Self, RnsNttError>   basis.rs:97-106
    static RnsBasis create(uint64_t degree, const std::vector<uint64_t> &moduli, int device = 0) {
        ckks_ctx *c = nullptr;
        check(ckks_ctx_create(degree, moduli.data(), moduli.size(), device, &c));
        return RnsBasis(c);
    }
    RnsBasis drop_last(size_t k) const {  // basis.rs:121-134
        ckks_ctx *c = nullptr;
        check(ckks_ctx_drop_last(h_.get(), k, &c));
        return RnsBasis(c);
    }
    std::vector<uint64_t> moduli() const {
        std::vector<uint64_t> m(channel_count());
        check(ckks_ctx_moduli(h_.get(), m.data()));
        return m;
    }
    size_t channel_count() const { return ckks_ctx_channel_count(h_.get()); }
    uint64_t degree() const { return ckks_ctx_degree(h_.get()); }
    uint32_t total_bits() const { return ckks_ctx_total_bits(h_.get()); }
    int64_t reconstruct_centered_coeff(const std::vector<uint64_t> &residues) const {  // basis.rs:158-180
        int64_t v = 0;
        check(ckks_ctx_reconstruct_centered_coeff(h_.get(), residues.data(), &v));
        return v;
    }
    void sync() const { check(ckks_ctx_sync(h_.get())); }
    ckks_ctx *raw() const { return h_.get(); }
};

class RnsPoly {
    ckks_poly *p_ = nullptr;
    RnsBasis basis_;
    RnsPoly(ckks_poly *p, RnsBasis b) : p_(p), basis_(std::move(b)) {}
    friend struct engine;
    friend class GadgetKey;
    friend class LimbShard;

  public:
    RnsPoly() = default;
    RnsPoly(RnsPoly &&o) noexcept : p_(o.p_), basis_(std::move(o.basis_)) { o.p_ = nullptr; }
    RnsPoly &operator=(RnsPoly &&o) noexcept {
        if (this != &o) {
            if (p_) ckks_poly_free(p_);
            p_ = o.p_;
            basis_ = std::move(o.basis_);
            o.p_ = nullptr;
        }
        return *this;
    }
    RnsPoly(const RnsPoly &o) : basis_(o.basis_) { check(ckks_poly_clone(o.p_, &p_)); }  // #[derive(Clone)]
    RnsPoly &operator=(const RnsPoly &o) {
        if (this != &o) *this = RnsPoly(o);
        return *this;
    }
    ~RnsPoly() {
        if (p_) ckks_poly_free(p_);
    }
    static RnsPoly zero(const RnsBasis &b, size_t batch = 1) {  // poly.rs:36-42
        ckks_poly *p = nullptr;
        check(ckks_poly_alloc(b.raw(), batch, &p));
        return RnsPoly(p, b);
    }
    // coeffs: [batch][coeffs_len]                                  poly.rs:49-66
    static RnsPoly from_coeffs(const std::vector<int64_t> &coeffs, size_t batch, const RnsBasis &b) {
        ckks_poly *p = nullptr;
        check(ckks_poly_from_coeffs(b.raw(), batch, coeffs.data(), batch ? coeffs.size() / batch : coeffs.size(), &p));
        return RnsPoly(p, b);
    }
    // channels: [batch][nchannels][N]                              poly.rs:72-99
    static RnsPoly from_channels(const std::vector<uint64_t> &channels, size_t batch, size_t nchannels, const RnsBasis &b, bool is_ntt_domain) {
        ckks_poly *p = nullptr;
        check(ckks_poly_from_channels(b.raw(), batch, channels.data(), nchannels, is_ntt_domain, &p));
        return RnsPoly(p, b);
    }
    std::vector<uint64_t> channels() const {  // poly.rs:119-121
        std::vector<uint64_t> out(batch() * basis_.channel_count() * basis_.degree());
        check(ckks_poly_download(p_, out.data()));
        return out;
    }
    const RnsBasis &basis() const { return basis_; }
    size_t batch() const { return ckks_poly_batch(p_); }
    bool is_ntt_domain() const { return ckks_poly_is_ntt_domain(p_) == 1; }
    void to_ntt_domain() { check(ckks_poly_to_ntt_domain(p_)); }
    void to_coeff_domain() { check(ckks_poly_to_coeff_domain(p_)); }
    RnsPoly &operator+=(const RnsPoly &rhs) {
        check(ckks_poly_add_assign(p_, rhs.p_));
        return *this;
    }
    RnsPoly &operator*=(const RnsPoly &rhs) {
        check(ckks_poly_mul_assign(p_, rhs.p_));
        return *this;
    }
    RnsPoly operator-() const {
        RnsPoly r(*this);
        check(ckks_poly_neg(r.p_));
        return r;
    }
    RnsPoly mod_drop_last(const RnsBasis &child) const {  // poly.rs:169-177
        ckks_poly *p = nullptr;
        check(ckks_poly_mod_drop_last(p_, child.raw(), &p));
        return RnsPoly(p, child);
    }
    RnsPoly rescale_into(const RnsBasis &child) const {  // poly.rs:187-228
        ckks_poly *p = nullptr;
        check(ckks_poly_rescale_into(p_, child.raw(), &p));
        return RnsPoly(p, child);
    }
    RnsPoly automorphism(uint64_t exponent) const {  // poly.rs:492-541
        ckks_poly *p = nullptr;
        check(ckks_poly_automorphism(p_, exponent, &p));
        return RnsPoly(p, basis_);
    }
    RnsPoly rotate_slots(int32_t k) const {  // poly.rs:546-569
        ckks_poly *p = nullptr;
        check(ckks_poly_rotate_slots(p_, k, &p));
        return RnsPoly(p, basis_);
    }
    std::vector<int64_t> to_coeffs() const {  // poly.rs:404-427
        std::vector<int64_t> out(batch() * basis_.degree());
        check(ckks_poly_to_coeffs(p_, out.data()));
        return out;
    }
    // The same for a basis of any size (Q >= 2^128 too): centred values as i64 (the reference's `as i64` truncation) and
    // as doubles; returns true if some |x| >= 2^63.  Below Q = 2^128 `wide_i64` == to_coeffs() bit for bit.
    bool to_coeffs_wide(std::vector<int64_t> *wide_i64, std::vector<double> *wide_f64) const {
        const size_t n = batch() * basis_.degree();
        if (wide_i64) wide_i64->assign(n, 0);
        if (wide_f64) wide_f64->assign(n, 0.0);
        int overflow = 0;
        check(ckks_poly_to_coeffs_wide(p_, wide_i64 ? wide_i64->data() : nullptr, wide_f64 ? wide_f64->data() : nullptr, &overflow));
        return overflow != 0;
    }
    RnsPoly rescale() const { return rescale_into(basis_.drop_last(1)); }  // poly.rs:246-249
    void mul_assign_naive(const RnsPoly &rhs) { check(ckks_poly_mul_assign_naive(p_, rhs.p_)); }  // poly.rs:339-367
    ckks_poly *raw() const { return p_; }
};

struct Ciphertext {  // types.rs:22-35
    RnsPoly c0, c1;
    uint32_t logp = 0, logq = 0;
};

class GadgetKey {  // engine.rs:225-253
    ckks_ksk *k_ = nullptr;
    friend struct engine;
    friend class LimbShard;

  public:
    int32_t rotation = 0;
    GadgetKey() = default;
    GadgetKey(const GadgetKey &) = delete;
    GadgetKey(GadgetKey &&o) noexcept : k_(o.k_), rotation(o.rotation) { o.k_ = nullptr; }
    ~GadgetKey() {
        if (k_) ckks_ksk_free(k_);
    }
    // a, b: [L][L][N] coefficient domain, as generated on the host (engine.rs:288-399)
    static GadgetKey upload(const RnsBasis &b, const std::vector<uint64_t> &a, const std::vector<uint64_t> &bb, int32_t rotation = 0) {
        GadgetKey k;
        check(ckks_ksk_upload(b.raw(), a.data(), bb.data(), &k.k_));
        k.rotation = rotation;
        return k;
    }
};

struct engine {
    static Ciphertext wrap(ckks_poly *p0, ckks_poly *p1, const RnsBasis &b, uint32_t logp, uint32_t logq) {
        Ciphertext c;
        c.c0 = RnsPoly(p0, b);
        c.c1 = RnsPoly(p1, b);
        c.logp = logp;
        c.logq = logq;
        return c;
    }
    static Ciphertext add_ciphertexts(const Ciphertext &a, const Ciphertext &b) {  // engine.rs:131-151
        if (a.logp != b.logp || a.logq != b.logq) throw RnsNttError(CKKS_LEVEL_MISMATCH);
        ckks_poly *p0, *p1;
        check(ckks_ct_add(a.c0.p_, a.c1.p_, b.c0.p_, b.c1.p_, &p0, &p1));
        return wrap(p0, p1, a.c0.basis(), a.logp, a.logq);
    }
    static Ciphertext mul_ciphertexts_gadget(const Ciphertext &a, const Ciphertext &b, const GadgetKey &rlk) {  // engine.rs:473-539
        if (a.logq != b.logq) throw RnsNttError(CKKS_LEVEL_MISMATCH);
        ckks_poly *p0, *p1;
        check(ckks_ct_mul_relin(a.c0.p_, a.c1.p_, b.c0.p_, b.c1.p_, rlk.k_, &p0, &p1));
        return wrap(p0, p1, a.c0.basis(), a.logp + b.logp, a.logq);
    }
    static Ciphertext rescale_ciphertext(const Ciphertext &ct) {  // engine.rs:263-282
        RnsBasis child = ct.c0.basis().drop_last(1);
        ckks_poly *p0, *p1;
        uint32_t bits = 0;
        check(ckks_ct_rescale(ct.c0.p_, ct.c1.p_, child.raw(), &p0, &p1, &bits));
        return wrap(p0, p1, child, ct.logp - bits, ct.logq - bits);
    }
    static Ciphertext rotate_ciphertext(const Ciphertext &ct, const GadgetKey &rotk) {  // engine.rs:412-463
        ckks_poly *p0, *p1;
        check(ckks_ct_rotate(ct.c0.p_, ct.c1.p_, rotk.k_, rotk.rotation, &p0, &p1));
        return wrap(p0, p1, ct.c0.basis(), ct.logp, ct.logq);
    }
    static RnsPoly decrypt(const Ciphertext &ct, const RnsPoly &s) {  // engine.rs:114-128
        ckks_poly *p;
        check(ckks_ct_decrypt(ct.c0.p_, ct.c1.p_, s.p_, &p));
        return RnsPoly(p, ct.c0.basis());
    }
};

// Optional limb-sharded mode (include/ckks_b200.h, ckks_lshard_*): this GPU's share -- limbs j = rank (mod world)
// -- of one batch and of the gadget keys; every rank of the group makes the same calls.
class LimbShard {
    struct Del {
        void operator()(ckks_lshard *s) const { ckks_lshard_destroy(s); }
    };
    std::shared_ptr<ckks_lshard> h_;
    std::shared_ptr<ckks_lshard> parent_;  // a child level shares its parent's exchange buffers
    int rank_ = 0, world_ = 1;
    explicit LimbShard(ckks_lshard *s, int rank, int world, std::shared_ptr<ckks_lshard> parent = nullptr)
        : h_(s, Del()), parent_(std::move(parent)), rank_(rank), world_(world) {}

  public:
    LimbShard() = default;
    static LimbShard create(uint64_t degree, const std::vector<uint64_t> &moduli, int rank, int world, int device = 0, size_t chunk = 0) {
        ckks_lshard *s = nullptr;
        check(ckks_lshard_create(degree, moduli.data(), moduli.size(), rank, world, device, chunk, &s));
        return LimbShard(s, rank, world);
    }
    LimbShard drop_last() const {
        ckks_lshard *c = nullptr;
        check(ckks_lshard_drop_last(h_.get(), &c));
        return LimbShard(c, rank_, world_, h_);
    }
    // RnsBasis over the limbs held here (local limb jl = basis limb rank + world * jl)
    RnsBasis local_basis() const { return RnsBasis(ckks_lshard_local_ctx(h_.get()), std::static_pointer_cast<void>(h_)); }
    size_t channel_count() const { return ckks_lshard_channel_count(h_.get()); }
    std::vector<unsigned char> ipc_handle() const {
        std::vector<unsigned char> b(ckks_lshard_ipc_size());
        check(ckks_lshard_ipc_export(h_.get(), b.data()));
        return b;
    }
    void connect(const std::vector<unsigned char> &handles_in_rank_order) { check(ckks_lshard_ipc_import(h_.get(), handles_in_rank_order.data())); }
    // a, b: [L][L_own][N], the rows of the key restricted to this GPU's limbs
    GadgetKey upload_key(const std::vector<uint64_t> &a, const std::vector<uint64_t> &b, int32_t rotation = 0) const {
        GadgetKey k;
        check(ckks_lshard_ksk_upload(h_.get(), a.data(), b.data(), &k.k_));
        k.rotation = rotation;
        return k;
    }
    // mul_ciphertexts_gadget (+ rescale_ciphertext into `child`'s level), engine.rs:473-539, 263-282
    Ciphertext mul_relin_rescale(const Ciphertext &a, const Ciphertext &b, const GadgetKey &rlk, const LimbShard *child = nullptr) const {
        if (a.logq != b.logq) throw RnsNttError(CKKS_LEVEL_MISMATCH);
        ckks_poly *p0, *p1;
        check(ckks_lshard_ct_mul_relin_rescale(h_.get(), a.c0.p_, a.c1.p_, b.c0.p_, b.c1.p_, rlk.k_, child ? child->h_.get() : nullptr, &p0, &p1));
        Ciphertext c;
        RnsBasis ob = child ? child->local_basis() : local_basis();
        c.c0 = RnsPoly(p0, ob);
        c.c1 = RnsPoly(p1, ob);
        c.logp = a.logp + b.logp;
        c.logq = a.logq;
        return c;  // the caller subtracts bit_length(q_last) from logp/logq when rescaling, as engine.rs:266-270 does
    }
    Ciphertext rotate_ciphertext(const Ciphertext &ct, const GadgetKey &rotk) const {  // engine.rs:412-463
        ckks_poly *p0, *p1;
        check(ckks_lshard_ct_rotate(h_.get(), ct.c0.p_, ct.c1.p_, rotk.k_, rotk.rotation, &p0, &p1));
        Ciphertext c;
        c.c0 = RnsPoly(p0, local_basis());
        c.c1 = RnsPoly(p1, local_basis());
        c.logp = ct.logp;
        c.logq = ct.logq;
        return c;
    }
    void check_peers() const { check(ckks_lshard_check(h_.get())); }  // sync + "no peer was lost"
};

// The batch-sharded multi-GPU group (ckks_comm_*): ONE process spreads a host batch of ciphertexts over the GPUs of a
// box; replaces the reference's serial loop over a Vec<Ciphertext> (examples/horner_chain.rs:211-278).
class BatchShard {
    struct Del {
        void operator()(ckks_comm *c) const { ckks_comm_destroy(c); }
    };
    std::shared_ptr<ckks_comm> h_;
    explicit BatchShard(ckks_comm *c) : h_(c, Del()) {}

public:
    class Key {
        friend class BatchShard;
        ckks_comm_ksk *k_ = nullptr;
        std::shared_ptr<ckks_comm> owner_;

    public:
        int32_t rotation = 0;
        Key() = default;
        Key(const Key &) = delete;
        Key(Key &&o) noexcept : k_(o.k_), owner_(std::move(o.owner_)), rotation(o.rotation) { o.k_ = nullptr; }
        ~Key() {
            if (k_) ckks_comm_ksk_free(k_);
        }
    };
    static BatchShard create(uint64_t degree, const std::vector<uint64_t> &moduli, const std::vector<int> &devices) {
        ckks_comm *c = nullptr;
        check(ckks_comm_init((int)devices.size(), devices.data(), degree, moduli.data(), moduli.size(), &c));
        return BatchShard(c);
    }
    BatchShard drop_last(size_t k) const {
        ckks_comm *c = nullptr;
        check(ckks_comm_drop_last(h_.get(), k, &c));
        return BatchShard(c);
    }
    int size() const { return ckks_comm_size(h_.get()); }
    // a, b: [L][L][N] as RnsGadgetRelinKey / RnsGadgetRotationKey hold them (engine.rs:225-253)
    Key upload_key(const std::vector<uint64_t> &a, const std::vector<uint64_t> &b, int32_t rotation = 0) const {
        Key k;
        check(ckks_comm_ksk_upload(h_.get(), a.data(), b.data(), &k.k_));
        k.owner_ = h_;
        k.rotation = rotation;
        return k;
    }
    // host [batch][L][N] in, [batch][L-1][N] out: mul_ciphertexts_gadget + rescale_ciphertext
    void mul_relin_rescale_host(const Key &rlk, size_t batch, const uint64_t *a0, const uint64_t *a1, const uint64_t *b0, const uint64_t *b1, uint64_t *o0,
                                uint64_t *o1) const {
        check(ckks_comm_ct_mul_relin_rescale_host(h_.get(), rlk.k_, batch, a0, a1, b0, b1, o0, o1));
    }
    void rotate_host(const Key &rotk, size_t batch, const uint64_t *c0, const uint64_t *c1, uint64_t *o0, uint64_t *o1) const {
        check(ckks_comm_ct_rotate_host(h_.get(), rotk.k_, rotk.rotation, batch, c0, c1, o0, o1));
    }
};

}  // namespace ckks
