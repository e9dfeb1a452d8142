/*
 * ckks_oracle.h -- CPU restatement (TEST INFRASTRUCTURE ONLY) of the RNS-CKKS arithmetic that
 * Hansard-T/FHE-Linformer reaches through OpenFHE (reference call sites:
 * /root/reference/src/FHEController.cpp:37-49,353,373-404,409-436).
 *
 * PARITY UNPINNED: OpenFHE (the third-party library that holds the arithmetic; version not pinned by
 * the reference's CMakeLists.txt:13) is absent from /root/reference and from this image, and the
 * reference ships no golden vectors, keys or ciphertexts.  This oracle restates the published
 * full-RNS CKKS algorithms with the OpenFHE conventions listed in SURVEY.md Appendix A; it is pinned
 * only by algebraic identities and an independent Python big-int model (tests/test_oracle_*.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product (fhe_linformer_b200/csrc) never links or calls it.
 */
#ifndef CKKS_ORACLE_H
#define CKKS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

/* context: prime chain (L Q-limbs: first_bits + (L-1) x scale_bits, FLEXIBLEAUTO rule; K P-limbs of
 * aux_bits), roots, twiddles, CRT constants.  dnum = number of key-switch digits. */
orc_ctx* orc_create(int logN, int L, int dnum, int first_bits, int scale_bits, int aux_bits);
void orc_destroy(orc_ctx* c);
/* info[0]=logN [1]=L [2]=K [3]=alpha [4]=dnum */
void orc_info(const orc_ctx* c, int* info);
void orc_moduli(const orc_ctx* c, uint64_t* out /* L+K */);
void orc_roots(const orc_ctx* c, uint64_t* out /* L+K, psi (2N-th root) */);
void orc_scale_factors(const orc_ctx* c, double* sf /* L, sf[level] */);

/* limb-wise primitives; midx[i] = modulus index (0..L-1 = Q, L..L+K-1 = P) of limb i */
void orc_ntt(const orc_ctx* c, uint64_t* a, const int* midx, int nl);
void orc_intt(const orc_ctx* c, uint64_t* a, const int* midx, int nl);
void orc_add(const orc_ctx* c, uint64_t* out, const uint64_t* a, const uint64_t* b, const int* midx, int nl);
void orc_sub(const orc_ctx* c, uint64_t* out, const uint64_t* a, const uint64_t* b, const int* midx, int nl);
void orc_mul(const orc_ctx* c, uint64_t* out, const uint64_t* a, const uint64_t* b, const int* midx, int nl);
/* out = a * (scalar mod q_i), scalar given as signed 128-bit (lo, hi) */
void orc_mul_scalar(const orc_ctx* c, uint64_t* out, const uint64_t* a, uint64_t s_lo, int64_t s_hi, const int* midx, int nl);
/* Galois automorphism X -> X^g on evaluation-format (bit-reversed) limbs */
void orc_automorph_eval(const orc_ctx* c, uint64_t* out, const uint64_t* in, int nl, uint32_t g);
/* same on coefficient-format limbs (with sign flips) */
void orc_automorph_coeff(const orc_ctx* c, uint64_t* out, const uint64_t* in, const int* midx, int nl, uint32_t g);
uint32_t orc_galois_for_rotation(const orc_ctx* c, int k);   /* 5^k mod 2N, k<0 -> inverse; */
uint32_t orc_galois_conj(const orc_ctx* c);                  /* 2N-1 */

/* rescale: in = l Q-limbs (eval) -> out = l-1 limbs */
void orc_rescale(const orc_ctx* c, uint64_t* out, const uint64_t* in, int l);
/* hybrid key switch pieces (l active Q limbs) */
void orc_modup(const orc_ctx* c, uint64_t* out /* (l+K) limbs eval; order Q0..l-1,P0..K-1 */,
               const uint64_t* in_eval /* l limbs eval */, int l, int digit);
void orc_moddown(const orc_ctx* c, uint64_t* out /* l */, const uint64_t* in /* l+K eval */, int l);
/* evk layout: [dnum][2 (b,a)][L+K][N] full chain */
void orc_keyswitch(const orc_ctx* c, uint64_t* out0, uint64_t* out1, const uint64_t* in /* l limbs */,
                   const uint64_t* evk, int l);
/* ct layout [2][l][N] */
void orc_rotate(const orc_ctx* c, uint64_t* out, const uint64_t* ct, int l, uint32_t g, const uint64_t* evk);
void orc_mul_relin(const orc_ctx* c, uint64_t* out, const uint64_t* a, const uint64_t* b, int l, const uint64_t* evk);
void orc_mul_plain(const orc_ctx* c, uint64_t* out, const uint64_t* ct, const uint64_t* pt, int l);

/* sampling (splitmix64 streams; spec in DESIGN.md "Randomness") */
void orc_sample_uniform(const orc_ctx* c, uint64_t seed, uint64_t* out, const int* midx, int nl);
void orc_sample_ternary(const orc_ctx* c, uint64_t seed, int8_t* out /* N */);
void orc_sample_sparse_ternary(const orc_ctx* c, uint64_t seed, int h, int8_t* out /* N */);
void orc_sample_gauss(const orc_ctx* c, uint64_t seed, int8_t* out /* N */);
/* small signed coefficients -> eval-format limbs */
void orc_small_to_eval(const orc_ctx* c, uint64_t* out, const int8_t* in, const int* midx, int nl);

/* keys.  sk_eval: (L+K) limbs eval.  pk: [2][L][N].  evk: [dnum][2][L+K][N]. */
void orc_gen_sk(const orc_ctx* c, uint64_t seed, int h /*0 = uniform ternary*/, uint64_t* sk_eval);
void orc_gen_pk(const orc_ctx* c, uint64_t seed, const uint64_t* sk_eval, uint64_t* pk);
void orc_gen_evk(const orc_ctx* c, uint64_t seed, const uint64_t* sk_old_eval, const uint64_t* sk_new_eval, uint64_t* evk);
void orc_gen_relin_key(const orc_ctx* c, uint64_t seed, const uint64_t* sk_eval, uint64_t* evk);
void orc_gen_galois_key(const orc_ctx* c, uint64_t seed, const uint64_t* sk_eval, uint32_t g, uint64_t* evk);

/* CKKS encoding: slots complex values (re, im interleaved) -> l limbs eval, at given scale */
void orc_encode(const orc_ctx* c, uint64_t* out, const double* vals_reim, int slots, double scale, int l);
/* int64 coefficient vector of the encoding (before RNS/NTT), for parity of the FFT+round stage */
void orc_encode_coeffs(const orc_ctx* c, int64_t* out /* N */, const double* vals_reim, int slots, double scale);
/* decode: poly with l limbs (eval) -> slots complex values */
void orc_decode(const orc_ctx* c, double* vals_reim, const uint64_t* poly, int slots, double scale, int l);
/* encrypt pt (l limbs eval) under pk with seeded randomness -> ct [2][l][N] */
void orc_encrypt(const orc_ctx* c, uint64_t seed, uint64_t* ct, const uint64_t* pt, const uint64_t* pk, int l);
/* decrypt: ct [ncomp][l][N] (ncomp 2 or 3) -> pt l limbs eval */
void orc_decrypt(const orc_ctx* c, uint64_t* pt, const uint64_t* ct, const uint64_t* sk_eval, int l, int ncomp);

int orc_num_threads(void);
void orc_set_threads(int n);
#ifdef __cplusplus
}
#endif
#endif
