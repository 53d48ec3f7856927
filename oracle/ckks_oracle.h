/*
 * ckks_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A C++17 restatement of the RNS-NTT hot path of oiwn/toy-heaan-ckks, following the
 * reference's algorithm AND schedule operation for operation (bit-reverse + radix-2 DIT with
 * natural-power root tables, psi pre/post twist, mul_mod = 128-bit product % q, coefficient
 * domain ciphertexts, every `*=` doing 2 forward + 1 inverse NTT per limb, L-digit gadget
 * re-transforming the key on every use).  Each function cites the reference file:line it
 * follows (paths relative to the reference repo root).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libckks_b200.so) never links or calls it.
 *
 * PARITY PIN: the Rust reference cannot be compiled in the build container (no cargo/rustc), so
 * this oracle is pinned against every known-answer test and algebraic identity the reference's
 * own unit tests hold for this path (tests/test_oracle_kats.py lists them with file:line) and
 * against the numpy encoder script shipped in the reference (scripts/reference_ckks_encode.py,
 * fixtures under tests/golden/).  RNG-level parity (ChaCha20 / rand_distr sampling) is
 * "parity unpinned": no reference test pins a sampled value, so sampled polynomials are supplied
 * by the caller as arrays and both the oracle and the CUDA path consume the same arrays.
 *
 * Layout convention (== reference `Vec<[u64; N]>`, poly.rs:26-30): a polynomial is L limbs
 * ("channels") of N u64 words, limb-major: ch[limb * N + coeff].
 */
#ifndef CKKS_ORACLE_H
#define CKKS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Error codes: 1..6 map 1:1 onto RnsNttError (src/rings/backends/rns_ntt/errors.rs:3-22). */
enum {
    ORC_OK = 0,
    ORC_INVALID_DEGREE = 1,
    ORC_EMPTY_BASIS = 2,
    ORC_NON_NTT_FRIENDLY_MODULUS = 3,
    ORC_INVALID_MOD_DROP = 4,
    ORC_CHANNEL_COUNT_MISMATCH = 5,
    ORC_NON_REDUCED_COEFFICIENT = 6,
    ORC_PANIC = 100 /* the reference would panic / assert here */
};

typedef struct orc_basis orc_basis;

/* ---- src/math/primes.rs, src/math/utils.rs ------------------------------------------------ */
int      orc_is_prime(uint64_t n);                                  /* primes.rs:67-93   */
int      orc_is_ntt_friendly_prime(uint64_t p, uint64_t n);         /* primes.rs:125-131 */
uint64_t orc_get_first_prime_up(uint32_t logq, uint64_t n);         /* primes.rs:171-187 */
uint64_t orc_get_first_prime_down(uint64_t bound, uint64_t n);      /* primes.rs:198-219; 0 = None */
int      orc_generate_primes(int bit_size, int count, uint64_t degree, uint64_t *out); /* utils.rs:47-80 */

/* ---- src/rings/backends/rns_ntt/basis.rs --------------------------------------------------- */
int      orc_basis_new(uint64_t n, const uint64_t *moduli, size_t l, orc_basis **out); /* :97-106, :21-84 */
void     orc_basis_free(orc_basis *b);
int      orc_basis_drop_last(const orc_basis *b, size_t drop_count, orc_basis **out);  /* :121-134 */
uint64_t orc_basis_degree(const orc_basis *b);
size_t   orc_basis_channel_count(const orc_basis *b);                                  /* :117-119 */
void     orc_basis_moduli(const orc_basis *b, uint64_t *out);                          /* :109-111 */
uint32_t orc_basis_total_bits(const orc_basis *b);                                     /* :140-145 */
uint64_t orc_basis_psi(const orc_basis *b, size_t channel);  /* psi chosen by find_primitive_root :217-237 */
/* which: 0 forward_roots, 1 inverse_roots, 2 twist_factors, 3 untwist_factors; out[N]. which=4: out[0]=n_inv */
void     orc_basis_table(const orc_basis *b, size_t channel, int which, uint64_t *out);
int64_t  orc_reconstruct_centered_coeff(const orc_basis *b, const uint64_t *residues); /* :158-180 */

/* ---- src/rings/backends/rns_ntt/poly.rs ---------------------------------------------------- */
void orc_from_coeffs(const orc_basis *b, const int64_t *coeffs, uint64_t *out);        /* :49-66   */
int  orc_from_channels_check(const orc_basis *b, const uint64_t *ch, size_t nch);      /* :72-99   */
void orc_to_ntt_domain(const orc_basis *b, uint64_t *ch);                              /* :136-148 */
void orc_to_coeff_domain(const orc_basis *b, uint64_t *ch);                            /* :154-166 */
void orc_add_assign(const orc_basis *b, uint64_t *a, const uint64_t *rhs);             /* :254-275 */
void orc_neg(const orc_basis *b, uint64_t *a);                                         /* :370-385 */
void orc_mul_assign(const orc_basis *b, uint64_t *a, const uint64_t *rhs, int in_ntt); /* :277-331 */
void orc_mul_assign_naive(const orc_basis *b, uint64_t *a, const uint64_t *rhs);       /* :339-367 */
/* out has (L-1)*N words, always coefficient domain. */
int  orc_rescale(const orc_basis *b, const uint64_t *ch, int in_ntt, uint64_t *out);   /* :187-228 */
/* returns the domain flag of the result (quirk: exponent % 2N == 0 clones, keeping the flag). */
int  orc_automorphism(const orc_basis *b, const uint64_t *ch, int in_ntt, uint64_t exponent, uint64_t *out); /* :492-541 */
int  orc_rotate_slots(const orc_basis *b, const uint64_t *ch, int in_ntt, int32_t k, uint64_t *out);         /* :546-569 */
void orc_to_coeffs(const orc_basis *b, const uint64_t *ch, int in_ntt, int64_t *out);  /* :404-427 */

/* ---- src/crypto/engine.rs (all polynomials coefficient domain, as the engine produces them) - */
/* pk.b*u + e0 + m ; pk.a*u + e1  (engine.rs:84-112); u/e0/e1 are the host-sampled polynomials.  */
void orc_encrypt(const orc_basis *b, const uint64_t *pk_b, const uint64_t *pk_a, const uint64_t *u,
                 const uint64_t *e0, const uint64_t *e1, const uint64_t *m, uint64_t *c0, uint64_t *c1);
void orc_decrypt(const orc_basis *b, const uint64_t *c0, const uint64_t *c1, const uint64_t *s,
                 uint64_t *out);                                                       /* :114-128 */
void orc_add_ciphertexts(const orc_basis *b, const uint64_t *a0, const uint64_t *a1, const uint64_t *b0,
                         const uint64_t *b1, uint64_t *c0, uint64_t *c1);              /* :131-151 */
/* key layout: key_a / key_b = [digit i][limb][N], coefficient domain (engine.rs:225-253). */
void orc_mul_ciphertexts_gadget(const orc_basis *b, const uint64_t *a0, const uint64_t *a1,
                                const uint64_t *b0, const uint64_t *b1, const uint64_t *rlk_a,
                                const uint64_t *rlk_b, uint64_t *c0, uint64_t *c1);    /* :473-539 */
/* out polys have (L-1) limbs. *bits_dropped = bit_length(q_last) (engine.rs:266-270). */
int  orc_rescale_ciphertext(const orc_basis *b, const uint64_t *c0, const uint64_t *c1, uint64_t *o0,
                            uint64_t *o1, uint32_t *bits_dropped);                     /* :263-282 */
void orc_rotate_ciphertext(const orc_basis *b, const uint64_t *c0, const uint64_t *c1,
                           const uint64_t *rotk_a, const uint64_t *rotk_b, int32_t rotation,
                           uint64_t *o0, uint64_t *o1);                                /* :412-463 */
/* b = -(a*s) + e  (src/keys/public_key.rs:111-131) */
void orc_gen_public_key(const orc_basis *b, const uint64_t *s, const uint64_t *a, const uint64_t *e,
                        uint64_t *out_b);
/* a, e: [L][L][N] host-sampled; out_b: [L][L][N]  (engine.rs:288-335 / :348-399) */
void orc_gen_gadget_relin_key(const orc_basis *b, const uint64_t *s, const uint64_t *a,
                              const uint64_t *e, uint64_t *out_b);
void orc_gen_gadget_rotation_key(const orc_basis *b, const uint64_t *s, int32_t rotation,
                                 const uint64_t *a, const uint64_t *e, uint64_t *out_b);

/* ---- src/encoding (f64 Vandermonde; decode-tolerance checks only) --------------------------- */
/* values: nvals complex (re,im interleaved), nvals <= n/2.  ckks_encoder.rs:65-122, special_fft.rs:194-220 */
void orc_encode(uint64_t n, uint32_t scale_bits, const double *values, size_t nvals, int64_t *out_coeffs);
/* out: slots complex (re,im interleaved).  ckks_encoder.rs:129-156, special_fft.rs:224-242 */
void orc_decode(uint64_t n, uint32_t scale_bits, const int64_t *coeffs, size_t slots, double *out);

/* ---- CPU baseline drivers (bench.py cpu_baseline / --impl reference) ------------------------ */
/* `count` independent mul_ciphertexts_gadget + rescale_ciphertext (reference schedule), spread over
 * `threads` host threads; inputs [count][L][N] each, outputs [count][L-1][N]. Returns seconds. */
double orc_bench_mul_rescale(const orc_basis *b, size_t count, int threads, const uint64_t *a0,
                             const uint64_t *a1, const uint64_t *b0, const uint64_t *b1,
                             const uint64_t *rlk_a, const uint64_t *rlk_b, uint64_t *o0, uint64_t *o1);
double orc_bench_mul_gadget(const orc_basis *b, size_t count, int threads, const uint64_t *a0,
                            const uint64_t *a1, const uint64_t *b0, const uint64_t *b1, const uint64_t *rlk_a,
                            const uint64_t *rlk_b, uint64_t *o0, uint64_t *o1);
double orc_bench_rotate(const orc_basis *b, size_t count, int threads, const uint64_t *c0,
                        const uint64_t *c1, const uint64_t *rotk_a, const uint64_t *rotk_b,
                        int32_t rotation, uint64_t *o0, uint64_t *o1);
/* `count` polynomials ([count][L][N]) each sent to_ntt_domain (dir=0) or to_coeff_domain (dir=1). */
double orc_bench_ntt(const orc_basis *b, size_t count, int threads, int dir, uint64_t *polys);

#ifdef __cplusplus
}
#endif
#endif
