/*
 * ckks_b200.h -- C ABI of libckks_b200.so, the B200 (sm_100a) RNS-NTT backend.
 *
 * This is the drop-in boundary for the reference's `RnsBasis<N>` / `RnsPoly<N>` backend
 * (src/rings/backends/rns_ntt/{basis,poly}.rs) and the RnsPoly-specific half of `CkksEngine`
 * (src/crypto/engine.rs:255-540).  Every entry point cites the reference item it replaces
 * (paths relative to the reference repository root).  A Rust `extern "C"` crate, the C++ mirror in
 * toy-heaan-ckks_b200/host/rns_poly.hpp and the Python ctypes mirror bind exactly these symbols.
 *
 * Conventions
 *   - plain pointers and sizes only; opaque handles; no exceptions/unwinding across the boundary;
 *   - every function returns 0 (CKKS_OK) or a ckks_status; codes 1..6 map 1:1 onto `RnsNttError`
 *     (src/rings/backends/rns_ntt/errors.rs:3-22);
 *   - host layout == reference layout: a polynomial is L limbs ("channels") of N u64 words,
 *     limb-major (`Vec<[u64; N]>`, poly.rs:26-30); a handle carries `batch` such polynomials,
 *     [batch][limb][N]; values are canonical representatives in [0, q_limb);
 *   - NTT-domain data crosses the boundary in the reference's natural order
 *     (slot k = p(psi^(2k+1)), poly.rs:136-148); the device-internal order is private;
 *   - all device work is enqueued on the context's stream; `*_download`, `ckks_ctx_sync` and the
 *     `*_host` calls block;
 *   - there is no CPU fallback: without a CUDA device every compute entry point returns
 *     CKKS_CUDA_ERROR.
 */
#ifndef CKKS_B200_H
#define CKKS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ckks_status {
    CKKS_OK = 0,
    CKKS_INVALID_DEGREE = 1,           /* RnsNttError::InvalidDegree            errors.rs:5  */
    CKKS_EMPTY_BASIS = 2,              /* RnsNttError::EmptyBasis               errors.rs:8  */
    CKKS_NON_NTT_FRIENDLY_MODULUS = 3, /* RnsNttError::NonNttFriendlyModulus    errors.rs:11 */
    CKKS_INVALID_MOD_DROP = 4,         /* RnsNttError::InvalidModDrop           errors.rs:14 */
    CKKS_CHANNEL_COUNT_MISMATCH = 5,   /* RnsNttError::ChannelCountMismatch     errors.rs:17 */
    CKKS_NON_REDUCED_COEFFICIENT = 6,  /* RnsNttError::NonReducedCoefficient    errors.rs:20 */
    CKKS_BASIS_MISMATCH = 20,          /* debug_assert!(Arc::ptr_eq) poly.rs:260-263,288-291 */
    CKKS_DOMAIN_MISMATCH = 21,         /* debug_assert_eq!(is_ntt)   poly.rs:264-267,292-295 */
    CKKS_BATCH_MISMATCH = 22,
    CKKS_LEVEL_MISMATCH = 23,          /* assert_eq!(logq/logp)      engine.rs:135-136,478   */
    CKKS_SHORT_INPUT = 24,             /* assert!(coeffs.len() >= N) poly.rs:50-54           */
    CKKS_BAD_HANDLE = 30,
    CKKS_BAD_ARGUMENT = 31,
    CKKS_UNSUPPORTED = 32,
    CKKS_CUDA_ERROR = 40,
    CKKS_NCCL_ERROR = 41
} ckks_status;

typedef struct ckks_ctx ckks_ctx;   /* Arc<RnsBasis<N>>  basis.rs:91-94  (+ device tables, stream) */
typedef struct ckks_poly ckks_poly; /* batch of RnsPoly<N>  poly.rs:26-30, device resident         */
typedef struct ckks_ksk ckks_ksk;   /* RnsGadgetRelinKey / RnsGadgetRotationKey engine.rs:225-253  */

/* Human-readable text for a status / the last CUDA error string seen by this thread. */
const char *ckks_status_str(int status);
const char *ckks_last_error(void);
/* Number of visible CUDA devices (0 when there is none: every compute call then fails loudly). */
int ckks_device_count(void);

/* ---- src/math/primes.rs, src/math/utils.rs (host number theory, identical prime chains) -------- */
int ckks_is_prime(uint64_t n);                                   /* primes.rs:67-93   */
int ckks_is_ntt_friendly_prime(uint64_t p, uint64_t n);          /* primes.rs:125-131 */
/* utils.rs:47-80 generate_primes(bit_size, count, degree); CKKS_BAD_ARGUMENT where it panics. */
int ckks_generate_primes(int bit_size, int count, uint64_t degree, uint64_t *out);

/* ---- RnsBasis<N>  (basis.rs) ------------------------------------------------------------------- */
/* RnsBasis::new(moduli) basis.rs:97-106 + NttTable::new basis.rs:21-84: validates power-of-two N
 * and q == 1 mod 2N prime; psi is chosen by the reference rule (find_primitive_root :217-237).
 * `device` is the CUDA ordinal. */
int ckks_ctx_create(uint64_t n, const uint64_t *moduli, size_t l, int device, ckks_ctx **out);
/* RnsBasis::drop_last(k) basis.rs:121-134 (shares the parent's device tables; the child keeps the
 * parent alive). */
int ckks_ctx_drop_last(ckks_ctx *ctx, size_t drop_count, ckks_ctx **out);
int ckks_ctx_destroy(ckks_ctx *ctx);
int ckks_ctx_sync(ckks_ctx *ctx);
/* Device scratch freed by the library stays cached in the context's PRIVATE memory pool (the device's default
 * pool, which other libraries of the process share, is never touched); this returns the cached memory. */
int ckks_ctx_trim(ckks_ctx *ctx);
/* Run this context's work on a caller-owned cudaStream_t (e.g. torch's current stream). */
int ckks_ctx_set_stream(ckks_ctx *ctx, void *cuda_stream);
uint64_t ckks_ctx_degree(const ckks_ctx *ctx);
size_t ckks_ctx_channel_count(const ckks_ctx *ctx);              /* basis.rs:117-119 */
int ckks_ctx_moduli(const ckks_ctx *ctx, uint64_t *out);         /* basis.rs:109-111 */
uint32_t ckks_ctx_total_bits(const ckks_ctx *ctx);               /* basis.rs:140-145 */
uint64_t ckks_ctx_psi(const ckks_ctx *ctx, size_t channel);      /* NttTable psi, basis.rs:33 */
/* RnsBasis::reconstruct_centered_coeff basis.rs:158-180 (host, u128 CRT, Q < 2^128). */
/* RnsBasis::ntt_table(channel) basis.rs:112-114 in the reference's NttTable layout (basis.rs:6-17): which = 0
 * forward_roots, 1 inverse_roots, 2 twist_factors, 3 untwist_factors (N words), 4 n_inv (1 word). */
int ckks_ctx_ntt_table(const ckks_ctx *ctx, size_t channel, int which, uint64_t *out);
int ckks_ctx_reconstruct_centered_coeff(const ckks_ctx *ctx, const uint64_t *residues, int64_t *out);
/* Selects the small-N single-CTA NTT (1) or the four-step NTT (2) for contexts created afterwards;
 * 0 = automatic.  Test hook: both paths must agree with the oracle. */
int ckks_set_ntt_path(int path);
/* Test hook: 1 = run the key-switch from its unfused building blocks (digit broadcast, NTT, MAC);
 * 0 (default) = fused ks_pass1 / ks_pass2 kernels on the four-step path.  Same results. */
int ckks_set_unfused(int on);
/* Test hook: contexts created afterwards with all moduli < 2^31 use 32-bit butterflies, tables and
 * scratch on the four-step path (1, default) or the generic 64-bit code (0).  Same results. */
int ckks_set_word32(int on);
/* Test hook: 1 (default) = moduli below 2^61 use the approximate-quotient butterflies (values in [0, 8q));
 * 0 = Harvey butterflies ([0, 4q)).  Same results. */
int ckks_set_lazy8(int on);
/* Test hook: 1 (default) = ks_pass2 stages its tiles with TMA (cp.async.bulk.tensor + mbarrier);
 * 0 = LDGSTS (cp.async).  Same results. */
int ckks_set_tma(int on);
/* Test hook: 1 (default) = 2^12 <= N <= 2^14 transforms run as one kernel with the limb resident in shared
 * memory; 0 = two passes through global memory.  Same results. */
int ckks_set_fused_ntt(int on);
/* Tuning knob of the fused key-switch: MiB of scratch per chunk of ciphertexts (default 8192: 28 ciphertexts at
 * N=2^16, L=24); ckks_ks_chunk reports the ciphertexts per chunk a batch of `batch` is cut into at ctx's level. */
int ckks_set_ks_scratch_mib(int mib);
size_t ckks_ks_chunk(const ckks_ctx *ctx, size_t batch);
/* Gadget product of mul_ciphertexts_gadget (engine.rs:501-541) through auxiliary 30-bit NTT primes (exact integer
 * convolution + Garner reconstruction, csrc/aux_ks.cuh; bit-identical results): 0 = never, 1 (default) = on the 64-bit
 * four-step path from 11 limbs up, 2 = whenever the path allows it (test hook).  Read when a key is uploaded (the
 * auxiliary form of the key is derived there) and when a product is computed. */
int ckks_set_ks_aux(int mode);
/* Tuning knob of the host-buffer entry points: MiB per component and pipeline chunk (default 64). */
int ckks_set_host_chunk_mib(int mib);

/* ---- RnsPoly<N>  (poly.rs) --------------------------------------------------------------------- */
/* RnsPoly::zero(basis) poly.rs:36-42, for `batch` polynomials (coefficient domain). */
int ckks_poly_alloc(ckks_ctx *ctx, size_t batch, ckks_poly **out);
/* RnsPoly::from_coeffs poly.rs:49-66: coeffs is [batch][coeffs_len] i64, coeffs_len >= N
 * (CKKS_SHORT_INPUT otherwise); rem_euclid per limb is computed on the device. */
int ckks_poly_from_coeffs(ckks_ctx *ctx, size_t batch, const int64_t *coeffs, size_t coeffs_len,
                          ckks_poly **out);
/* RnsPoly::from_channels poly.rs:72-99: host [batch][channels][N]; CKKS_CHANNEL_COUNT_MISMATCH if
 * channels != L, CKKS_NON_REDUCED_COEFFICIENT if any word >= its modulus (checked on the device). */
int ckks_poly_from_channels(ckks_ctx *ctx, size_t batch, const uint64_t *channels, size_t nchannels,
                            int in_ntt_domain, ckks_poly **out);
/* RnsPoly::channels() poly.rs:119-121: copies [batch][L][N] to the host in the reference layout
 * (natural NTT order if the polynomial is in the NTT domain). */
int ckks_poly_download(ckks_poly *p, uint64_t *out);
int ckks_poly_clone(ckks_poly *p, ckks_poly **out);              /* #[derive(Clone)] poly.rs:25 */
int ckks_poly_free(ckks_poly *p);
size_t ckks_poly_batch(const ckks_poly *p);
size_t ckks_poly_channel_count(const ckks_poly *p);
int ckks_poly_is_ntt_domain(const ckks_poly *p);                 /* poly.rs:127-129 */
int ckks_poly_to_ntt_domain(ckks_poly *p);                       /* poly.rs:136-148 */
int ckks_poly_to_coeff_domain(ckks_poly *p);                     /* poly.rs:154-166 */
int ckks_poly_add_assign(ckks_poly *a, const ckks_poly *rhs);    /* AddAssign poly.rs:254-275 */
int ckks_poly_sub_assign(ckks_poly *a, const ckks_poly *rhs);    /* a += -rhs (Neg + AddAssign) */
int ckks_poly_neg(ckks_poly *a);                                 /* Neg poly.rs:370-385 */
/* MulAssign poly.rs:277-331: both NTT domain -> pointwise; both coefficient domain -> negacyclic
 * product returned in the coefficient domain. */
int ckks_poly_mul_assign(ckks_poly *a, const ckks_poly *rhs);
/* mul_assign_naive poly.rs:339-367: O(N^2) schoolbook product in Z_q[X]/(X^N+1), coefficient domain only
 * (CKKS_DOMAIN_MISMATCH otherwise); the reference keeps it as the yardstick for the NTT path. */
int ckks_poly_mul_assign_naive(ckks_poly *a, const ckks_poly *rhs);
/* RnsPoly::mod_drop_last(k) poly.rs:169-177; `child` must be ckks_ctx_drop_last(ctx, k). */
int ckks_poly_mod_drop_last(const ckks_poly *p, ckks_ctx *child, ckks_poly **out);
/* RnsPoly::rescale_into(new_basis) poly.rs:187-228 (floor division by the last prime; result in
 * the coefficient domain; CKKS_INVALID_MOD_DROP if L < 2). */
int ckks_poly_rescale_into(const ckks_poly *p, ckks_ctx *child, ckks_poly **out);
/* PolyAutomorphism::automorphism poly.rs:492-541 (X -> X^exponent, any exponent incl. the even
 * and `% 2N == 0` quirks) and rotate_slots poly.rs:546-569.  Result in the coefficient domain
 * (except the exponent % 2N == 0 clone, which keeps the domain flag). */
int ckks_poly_automorphism(const ckks_poly *p, uint64_t exponent, ckks_poly **out);
int ckks_poly_rotate_slots(const ckks_poly *p, int32_t k, ckks_poly **out);
/* PolyRing::to_coeffs poly.rs:404-427: centred CRT into [batch][N] i64 (Q < 2^128). */
int ckks_poly_to_coeffs(const ckks_poly *p, int64_t *out);
/* The same for a basis of ANY size: RnsBasis::reconstruct_centered_coeff (basis.rs:158-180) forms Q
 * in a u128 and stops working at Q >= 2^128; this runs Garner's mixed-radix CRT on the device (no big integers).
 * out_i64 [batch][N]: the centred value truncated to 64 bits like the reference's `as i64` -- for Q < 2^128 the
 * reference's result bit for bit, the true value whenever |x| < 2^63 (*overflow = 1 if some |x| >= 2^63);
 * out_f64 [batch][N]: the centred value rounded to double.  Either output may be NULL; L <= 64. */
int ckks_poly_to_coeffs_wide(const ckks_poly *p, int64_t *out_i64, double *out_f64, int *overflow);

/* ---- gadget keys  (engine.rs:225-253, generated by engine.rs:288-399 on the host) -------------- */
/* a, b: [digit i < L][limb j < L][N] coefficient domain, as `rlk.a[i].channels()`.  The key is
 * transformed once and stays resident in HBM.  Words must be canonical, as in every RnsPoly the reference
 * builds (from_channels poly.rs:83-93): CKKS_NON_REDUCED_COEFFICIENT otherwise (scanned on the device). */
int ckks_ksk_upload(ckks_ctx *ctx, const uint64_t *a, const uint64_t *b, ckks_ksk **out);
/* Same, from device polynomials of batch L (digit-major), e.g. produced by ckks_gen_gadget_key. */
int ckks_ksk_from_polys(const ckks_poly *a, const ckks_poly *b, ckks_ksk **out);
int ckks_ksk_free(ckks_ksk *k);
/* engine.rs:304-332 / :364-392 with host-sampled a_i, e_i (batch L each, coefficient domain):
 * b_i = -(a_i*s) + e_i + [target in limb i].  `target` is s^2 (relin) or rotate_slots(s, k). */
int ckks_gen_gadget_key_b(const ckks_poly *s, const ckks_poly *target, const ckks_poly *a,
                          const ckks_poly *e, ckks_poly **out_b);

/* ---- CkksEngine<RnsPoly, N> ciphertext ops (engine.rs), batched, coefficient domain ------------ */
/* add_ciphertexts engine.rs:131-151 */
int ckks_ct_add(const ckks_poly *a0, const ckks_poly *a1, const ckks_poly *b0, const ckks_poly *b1,
                ckks_poly **c0, ckks_poly **c1);
/* mul_ciphertexts_gadget engine.rs:473-539 */
int ckks_ct_mul_relin(const ckks_poly *a0, const ckks_poly *a1, const ckks_poly *b0,
                      const ckks_poly *b1, const ckks_ksk *rlk, ckks_poly **c0, ckks_poly **c1);
/* rescale_ciphertext engine.rs:263-282; *bits_dropped = bit_length(q_last). */
int ckks_ct_rescale(const ckks_poly *c0, const ckks_poly *c1, ckks_ctx *child, ckks_poly **o0,
                    ckks_poly **o1, uint32_t *bits_dropped);
/* mul_ciphertexts_gadget followed by rescale_ciphertext, fused on the device. */
int ckks_ct_mul_relin_rescale(const ckks_poly *a0, const ckks_poly *a1, const ckks_poly *b0,
                              const ckks_poly *b1, const ckks_ksk *rlk, ckks_ctx *child,
                              ckks_poly **o0, ckks_poly **o1);
/* rotate_ciphertext engine.rs:412-463 with rotation `k` (the key's `rotation` field). */
int ckks_ct_rotate(const ckks_poly *c0, const ckks_poly *c1, const ckks_ksk *rotk, int32_t k,
                   ckks_poly **o0, ckks_poly **o1);
/* encrypt engine.rs:84-112 with host-sampled u, e0, e1 already uploaded: c0 = pk_b*u + e0 + m,
 * c1 = pk_a*u + e1.  pk_* have batch 1 or `batch`. */
int ckks_ct_encrypt(const ckks_poly *pk_b, const ckks_poly *pk_a, const ckks_poly *u,
                    const ckks_poly *e0, const ckks_poly *e1, const ckks_poly *m, ckks_poly **c0,
                    ckks_poly **c1);
/* decrypt engine.rs:114-128: c1*s + c0 (s has batch 1 or `batch`). */
int ckks_ct_decrypt(const ckks_poly *c0, const ckks_poly *c1, const ckks_poly *s, ckks_poly **out);

/* ---- CkksEncoder<N> (ckks_encoder.rs:65-156) on the device: O(N log N) instead of the reference's O(N^2) --- */
/* encode_complex: values [batch][nvals] complex (re, im interleaved), nvals <= N/2 (CKKS_SHORT_INPUT otherwise,
 * the reference asserts); scaled by 2^scale_bits, conjugate-symmetric slots, inverse canonical embedding, rounding
 * (f64::round), from_coeffs.  f64 summation order differs from the reference: tolerance-checked, not bit-exact. */
int ckks_encode(ckks_ctx *ctx, uint32_t scale_bits, size_t batch, const double *values, size_t nvals, ckks_poly **out);
/* decode_complex: centred CRT, canonical embedding, first `nslots` slots divided by 2^scale_bits; out
 * [batch][nslots] complex.  Q < 2^128: the CRT is the reference's u128 arithmetic (basis.rs:158-180); Q >= 2^128,
 * where the reference cannot decode: the mixed-radix CRT of ckks_poly_to_coeffs_wide, so a ciphertext at any level
 * (e.g. cfg4, 24 x 61 bits) decodes without mod_drop_last. */
int ckks_decode(const ckks_poly *p, uint32_t scale_bits, size_t nslots, double *out);

/* ---- host-buffer entry points (what a reference-side caller with `Vec<[u64;N]>` data uses) ------ */
/* mul_ciphertexts_gadget + rescale_ciphertext on `batch` ciphertext pairs held in HOST memory in
 * the reference layout ([batch][L][N] per component; outputs [batch][L-1][N]).  Copies are chunked
 * through pinned staging buffers and overlapped with compute.  Every staged chunk goes through the reducedness
 * scan from_channels applies (poly.rs:83-93): a word >= its modulus makes the call return
 * CKKS_NON_REDUCED_COEFFICIENT (the outputs are then unspecified); the same holds for ckks_ct_rotate_host. */
int ckks_ct_mul_relin_rescale_host(ckks_ctx *ctx, ckks_ctx *child, const ckks_ksk *rlk, size_t batch,
                                   const uint64_t *a0, const uint64_t *a1, const uint64_t *b0,
                                   const uint64_t *b1, uint64_t *o0, uint64_t *o1);
int ckks_ct_rotate_host(ckks_ctx *ctx, const ckks_ksk *rotk, int32_t k, size_t batch,
                        const uint64_t *c0, const uint64_t *c1, uint64_t *o0, uint64_t *o1);
/* Pinned host allocation helpers for the callers of the *_host entry points. */
int ckks_host_alloc(size_t bytes, void **out);
int ckks_host_free(void *p);

/* ---- batch-sharded multi-GPU group (SURVEY.md 8b "multi-GPU", 8e default mode) -----------------------------
 * Replaces the reference's serial loop over a `Vec<Ciphertext>` (examples/horner_chain.rs:211-278 calling
 * engine.rs:473-539 / :263-282 per ciphertext): ONE process spreads a host batch over the GPUs of a box.
 * Ciphertexts are independent, so there is no inter-GPU traffic; the gadget key is replicated once per device.
 * Every *_host call cuts the batch into contiguous shares (sizes differ by at most one) and runs the
 * single-GPU host pipeline on each device from its own host thread.  Output words equal the single-GPU ones. */
typedef struct ckks_comm ckks_comm;
typedef struct ckks_comm_ksk ckks_comm_ksk;
/* devices: `ndev` CUDA ordinals (NULL = 0..ndev-1; an ordinal may repeat).  Validates like RnsBasis::new. */
int ckks_comm_init(int ndev, const int *devices, uint64_t n, const uint64_t *moduli, size_t l, ckks_comm **out);
int ckks_comm_destroy(ckks_comm *c);
int ckks_comm_drop_last(ckks_comm *c, size_t drop_count, ckks_comm **out);  /* basis.rs:121-134 on every device */
int ckks_comm_size(const ckks_comm *c);
ckks_ctx *ckks_comm_ctx(ckks_comm *c, int i);  /* borrowed: the context of device slot i */
/* One host key (layout of ckks_ksk_upload) -> a transformed replica per device, uploaded concurrently. */
int ckks_comm_ksk_upload(ckks_comm *c, const uint64_t *a, const uint64_t *b, ckks_comm_ksk **out);
int ckks_comm_ksk_free(ckks_comm_ksk *k);
/* Same contracts as ckks_ct_mul_relin_rescale_host / ckks_ct_rotate_host (reducedness scan included). */
int ckks_comm_ct_mul_relin_rescale_host(ckks_comm *c, const ckks_comm_ksk *rlk, size_t batch, const uint64_t *a0,
                                        const uint64_t *a1, const uint64_t *b0, const uint64_t *b1, uint64_t *o0,
                                        uint64_t *o1);
int ckks_comm_ct_rotate_host(ckks_comm *c, const ckks_comm_ksk *rotk, int32_t k, size_t batch, const uint64_t *c0,
                             const uint64_t *c1, uint64_t *o0, uint64_t *o1);

/* ---- optional limb-sharded mode (north_star "limb-sharded mode at N=2^16 / L>=24") ------
 * Replaces nothing in the reference (it is single-threaded and single-device); it is the multi-GPU form
 * of mul_ciphertexts_gadget engine.rs:473-539 + rescale_ciphertext engine.rs:263-282 for ONE batch whose
 * limbs are spread over the GPUs of a box: limb j of the basis lives on GPU (j mod world); local limb jl of
 * rank r is basis limb r + world*jl.  Each GPU keeps 1/world of every ciphertext and of the gadget key.
 * The all-gather of the digits and the broadcast of the dropped limb are plain stores into peer HBM
 * issued by the kernels that produce those words, ordered by a flag barrier in peer memory; every
 * GPU of the group makes the same calls on its own share.  Output words equal the reference's. */
typedef struct ckks_lshard ckks_lshard;
/* moduli: the WHOLE basis (validated like RnsBasis::new basis.rs:97-106).  chunk: ciphertexts per pass
 * (0 = as many as 4 GiB of key-switch scratch allow).  CKKS_UNSUPPORTED if l < world or N < 256. */
int ckks_lshard_create(uint64_t n, const uint64_t *moduli, size_t l, int rank, int world, int device,
                       size_t chunk, ckks_lshard **out);
int ckks_lshard_destroy(ckks_lshard *s);
/* The level below (RnsBasis::drop_last basis.rs:122-134 for the whole basis): the owner of the dropped
 * limb loses one local limb; exchange buffers are shared with the parent. */
int ckks_lshard_drop_last(ckks_lshard *s, ckks_lshard **child);
/* Context over the limbs held here: build / read this GPU's share of a polynomial with the ckks_poly_* calls. */
ckks_ctx *ckks_lshard_local_ctx(ckks_lshard *s);
size_t ckks_lshard_channel_count(const ckks_lshard *s); /* limbs of the whole basis at this level */
size_t ckks_lshard_chunk(const ckks_lshard *s);
/* One process per GPU: export this rank's buffer handle (ckks_lshard_ipc_size() bytes), move the
 * handles of all ranks with any host-side all-gather, import them in rank order. */
size_t ckks_lshard_ipc_size(void);
int ckks_lshard_ipc_export(ckks_lshard *s, void *blob);
int ckks_lshard_ipc_import(ckks_lshard *s, const void *blobs);
/* All ranks in one process (several GPUs with peer access, or several ranks on one GPU). */
int ckks_lshard_connect_local(ckks_lshard **shards, int world);
/* Key slice: a, b host [digit i < L][own limb jl][N] = rows of RnsGadgetRelinKey (engine.rs:225-253)
 * restricted to this GPU's limbs, coefficient domain; transformed once. */
int ckks_lshard_ksk_upload(ckks_lshard *s, const uint64_t *a, const uint64_t *b, ckks_ksk **out);
/* mul_ciphertexts_gadget on this GPU's limbs of a batch; with child != NULL followed by
 * rescale_ciphertext into the child's level.  Inputs / outputs: polynomials of the local contexts,
 * coefficient domain.  Enqueues everything, exchanges included, without synchronising the host. */
int ckks_lshard_ct_mul_relin_rescale(ckks_lshard *s, const ckks_poly *a0, const ckks_poly *a1,
                                     const ckks_poly *b0, const ckks_poly *b1, const ckks_ksk *rlk,
                                     ckks_lshard *child, ckks_poly **o0, ckks_poly **o1);
/* The same in three phases (0: limb-local up to the digits, 1: key-switch, 2: rescale epilogue) over the
 * ciphertexts [s0, s0+cs), cs <= chunk, for a caller that runs the two exchanges itself on the exported
 * buffers (peer_stores = 0: e.g. NCCL all-gather / broadcast through torch.distributed) or that wants to
 * place the barriers (peer_stores = 1).  o0/o1: preallocated on the output level's local context. */
int ckks_lshard_mul_phase(ckks_lshard *s, int phase, size_t s0, size_t cs, const ckks_poly *a0,
                          const ckks_poly *a1, const ckks_poly *b0, const ckks_poly *b1,
                          const ckks_ksk *rlk, ckks_lshard *child, ckks_poly *o0, ckks_poly *o1,
                          int peer_stores);
/* rotate_ciphertext engine.rs:412-463 on this GPU's limbs (rotate_slots is limb-local; the rotated c1 is the
 * digit polynomial, pushed to every GPU); rotk: a key slice from ckks_lshard_ksk_upload. */
int ckks_lshard_ct_rotate(ckks_lshard *s, const ckks_poly *c0, const ckks_poly *c1, const ckks_ksk *rotk,
                          int32_t k, ckks_poly **o0, ckks_poly **o1);
/* The gadget key-switch of a coefficient-domain polynomial (engine.rs:429-452) in two phases (0: digits to
 * every GPU's gather buffer, 1: ks0/ks1 = sum_i alpha_i * key_b[i] / key_a[i] on the own limbs). */
int ckks_lshard_ks_phase(ckks_lshard *s, int phase, size_t s0, size_t cs, const ckks_poly *digits,
                         const ckks_ksk *key, ckks_poly *ks0, ckks_poly *ks1, int peer_stores);
int ckks_lshard_barrier(ckks_lshard *s);
/* Barrier by stream events for a group living in one process and driven phase by phase in lockstep. */
int ckks_lshard_barrier_local(ckks_lshard **shards, int world);
/* gather: [L of the top level][chunk][N] (digit limb i of chunk ciphertext c at (i*chunk + c)*N words);
 * last: [2][chunk][N] (finished last limb of c0 and c1). */
int ckks_lshard_buffers(ckks_lshard *s, uint64_t **gather, size_t *gather_words, uint64_t **last,
                        size_t *last_words);
/* Synchronise and report a barrier that gave up waiting for a peer (CKKS_NCCL_ERROR).
 * Failure model (fail closed): a barrier that times out writes an error word the host sees at once.  The kernels
 * behind it in the same call have consumed buffers the lost peer never filled, so a guard kernel at the end of every
 * entry point overwrites the outputs with all-ones words (never canonical: rejected by every reducedness scan); the
 * first blocking call on the local context (ckks_poly_download, ckks_ctx_sync, ckks_lshard_check) and EVERY later
 * ckks_lshard_* call return CKKS_NCCL_ERROR.  The error is sticky: the epochs of this rank are out of step with its
 * peers for good.  To recover, destroy every shard of the group on every rank and build the group again
 * (ckks_lshard_create + ipc export/import or connect_local): that re-creates buffers, flag words and epochs. */
int ckks_lshard_check(ckks_lshard *s);
int ckks_lshard_set_timeout_ms(ckks_lshard *s, uint64_t ms);
/* How the digits travel in the one-call entry points: 0 (default) = stores into peer HBM from the producing
 * kernel; 1 = that kernel stores locally and copy engines push the limbs to the peers (no SM is held while
 * NVLink is busy, but measured SLOWER on B200: 5126 vs 5427 ct-mult/s at 4 GPUs, DESIGN 5b -- kept as a knob). */
int ckks_lshard_set_exchange(ckks_lshard *s, int mode);

/* ---- instrumentation ---------------------------------------------------------------------------- */
/* Kernel launches issued by this library since process start (bench.py's gpu_launches). */
uint64_t ckks_launch_count(void);
/* Name/launch-count table of this library's kernels: writes up to `cap` bytes of
 * "name=count\n" lines, returns the length needed. */
size_t ckks_launch_table(char *buf, size_t cap);
/* Per-kernel timing: while enabled every launch is bracketed by CUDA events on the context's stream;
 * collect() synchronises and writes "name=launches,total_ms\n" lines (returns the length needed). */
int ckks_prof_enable(int on);
/* Allocator statistics since the last reset, for requests of 1 MiB and more: served by the library's block cache,
 * sent to the driver's pool, host microseconds spent there in total and in the slowest call. */
int ckks_alloc_stats(uint64_t *cache_hits, uint64_t *pool_allocs, uint64_t *pool_us, uint64_t *pool_max_us, int reset);
/* NVTX ranges (one per kernel launch, named like the launch table, plus "ckks:<entry point>" around the fused
 * pipelines) for Nsight timelines: off by default; 1 = on (or CKKS_NVTX=1 in the environment). */
int ckks_set_nvtx(int on);
size_t ckks_prof_collect(char *buf, size_t cap);
/* from_channels with the source already in device memory ([batch][L][N], reference layout), and the
 * raw device pointer of a polynomial (coefficient domain: reference layout; NTT domain: internal
 * order) for zero-copy interop with the caller's own CUDA code. */
int ckks_poly_from_device(ckks_ctx *ctx, size_t batch, const uint64_t *dev_channels, int in_ntt_domain,
                          ckks_poly **out);
int ckks_poly_device_ptr(ckks_poly *p, uint64_t **out);
/* Integer-pipe microbenchmark: dependent-free 64-bit Shoup modmuls; returns modmul/s (0 on error). */
double ckks_bench_modmul_peak(int device, int iters);
/* The same for the multiply-accumulate of the auxiliary-basis key-switch: 32 x 32 -> 64-bit products added into 64-bit
 * accumulators (IMAD.WIDE.U32, both factors in general registers), per second on the whole device. */
double ckks_bench_mac32_peak(int device, int iters);
/* vary = 0: warp-uniform multiplier (fed from a uniform register); 1: a different multiplier per thread, in a general
 * register, as in aux_mac_kernel (what ckks_bench_mac32_peak measures). */
double ckks_bench_mac32_peak_ex(int device, int iters, int vary);
/* Host-side copy ceiling of the *_host entry points: n_src H2D copies of [hsrc, +src_bytes) and n_dst D2H copies
 * into [hdst, +dst_bytes), both directions concurrently, in the pipeline's chunk size, no kernels; seconds per
 * iteration (0 on error). */
double ckks_bench_host_copy(int device, const void *hsrc, size_t src_bytes, int n_src, void *hdst, size_t dst_bytes,
                            int n_dst, int iters);

#ifdef __cplusplus
}
#endif
#endif /* CKKS_B200_H */
