/*
 * fl_ckks.h -- C-ABI of the B200-native CKKS evaluation engine that stands behind the reference's
 * FHEController (drop-in boundary of SURVEY.md section 8(b)).
 *
 * Every entry point replaces one OpenFHE call the reference makes through
 * `CryptoContext<DCRTPoly> context` (/root/reference/src/FHEController.h:23); the call site is cited
 * next to each declaration (F.cpp = /root/reference/src/FHEController.cpp, M = src/main.cpp).
 * The reference-side binding (a re-backed FHEController.cpp) is shown in INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a non-zero
 * status otherwise, with the message available from fl_last_error() (thread-local).  There is NO CPU
 * fallback: fl_ctx_create fails when no CUDA device is present.
 *
 * Data layout: a polynomial with l limbs is uint64_t[l][N], limb-major, residues in [0, q_i),
 * EVALUATION format in OpenFHE's bit-reversed order.  A ciphertext is
 * uint64_t[2][l][N] (c0 then c1).  An evaluation key is uint64_t[dnum][2][L+K][N] (b then a per digit).
 */
#ifndef FL_CKKS_H
#define FL_CKKS_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct fl_ctx fl_ctx;

/* CCParams<CryptoContextCKKSRNS> as set at F.cpp:4-35 */
typedef struct fl_params {
    int logN;        /* SetRingDim(1<<15)            F.cpp:13 (1<<16 variant F.cpp:12) */
    int L;           /* SetMultiplicativeDepth(27)+1 F.cpp:35 */
    int dnum;        /* SetNumLargeDigits(4)         F.cpp:11 */
    int first_bits;  /* SetFirstModSize(55)          F.cpp:25 */
    int scale_bits;  /* SetScalingModSize(52)        F.cpp:23 */
    int aux_bits;    /* OpenFHE default 60-bit P primes */
    int sparse_h;    /* SPARSE_TERNARY weight        F.cpp:8  (0 = uniform ternary) */
} fl_params;

const char* fl_last_error(void);

/* ---- context: GenCryptoContext F.cpp:37 ---- */
int fl_ctx_create(const fl_params* p, int device, fl_ctx** out);
void fl_ctx_destroy(fl_ctx* c);
/* info[0..4] = logN, L, K, alpha, dnum; info[5..7] = allocator statistics: pool allocations, cache trims, cached MiB (8 ints) */
int fl_ctx_info(fl_ctx* c, int* info);
int fl_ctx_moduli(fl_ctx* c, uint64_t* out /* L+K */);
int fl_ctx_roots(fl_ctx* c, uint64_t* out /* L+K */);
int fl_ctx_scale_factors(fl_ctx* c, double* out /* L */);
uint32_t fl_galois_for_rotation(fl_ctx* c, int k);   /* FindAutomorphismIndex2nComplex inside EvalRotate F.cpp:435 */
uint32_t fl_galois_conj(fl_ctx* c);
void* fl_ctx_stream(fl_ctx* c);                       /* cudaStream_t the engine launches on */
int fl_sync(fl_ctx* c);
/* Cap of the context's device block cache (freed operands are kept for reuse up to this many bytes; default 96 GB of the 180 GB).
 * A context parameter: lower it when several contexts share one GPU (one controller per host thread). */
int fl_ctx_set_cache_bytes(fl_ctx* c, uint64_t bytes);

/* ---- device buffers (HBM-resident operands) ---- */
int fl_dev_alloc(fl_ctx* c, size_t words, uint64_t** out);
int fl_dev_free(fl_ctx* c, uint64_t* p);
int fl_dev_upload(fl_ctx* c, uint64_t* dst, const uint64_t* src_host, size_t words);
int fl_dev_download(fl_ctx* c, uint64_t* dst_host, const uint64_t* src, size_t words);

/* ---- raw primitives on device buffers (asynchronous on the engine stream) ----
 * midx[i] = modulus index of limb i (0..L-1 Q limbs, L..L+K-1 P limbs) */
int fl_raw_ntt(fl_ctx* c, uint64_t* d, const int* midx, int nl);    /* DCRTPoly::SwitchFormat -> EVALUATION */
int fl_raw_intt(fl_ctx* c, uint64_t* d, const int* midx, int nl);   /* DCRTPoly::SwitchFormat -> COEFFICIENT */
int fl_raw_ntt_batch(fl_ctx* c, uint64_t* d, const int* midx, int nl, int batch, int inverse);   /* d: [batch][nl][N], one launch pair */
int fl_raw_add(fl_ctx* c, uint64_t* out, const uint64_t* a, const uint64_t* b, const int* midx, int nl);   /* EvalAdd F.cpp:410 */
int fl_raw_sub(fl_ctx* c, uint64_t* out, const uint64_t* a, const uint64_t* b, const int* midx, int nl);
int fl_raw_mul(fl_ctx* c, uint64_t* out, const uint64_t* a, const uint64_t* b, const int* midx, int nl);
int fl_raw_automorph(fl_ctx* c, uint64_t* out, const uint64_t* in, int nl, uint32_t g);                    /* AutomorphismTransform */
int fl_raw_rescale(fl_ctx* c, uint64_t* out, const uint64_t* in, int l, int polys);   /* ModReduceInternal (FLEXIBLEAUTO, F.cpp:18) */
int fl_raw_modup(fl_ctx* c, uint64_t* out_ext, const uint64_t* c_eval, int l, int digit);   /* ApproxModUp */
int fl_raw_moddown(fl_ctx* c, uint64_t* out, const uint64_t* in_ext, int l);                /* ApproxModDown */
int fl_raw_keyswitch(fl_ctx* c, uint64_t* out2 /* [2][l][N] */, const uint64_t* poly, const uint64_t* evk, int l);   /* KeySwitch (HYBRID) */
int fl_raw_rotate(fl_ctx* c, uint64_t* out, const uint64_t* ct, int l, uint32_t g, const uint64_t* evk);   /* EvalRotate F.cpp:435,833,843 */
/* the same for `batch` ciphertexts stored back to back ([batch][2][l][N]) and one key: every stage is one launch */
int fl_raw_rotate_batch(fl_ctx* c, uint64_t* out, const uint64_t* ct, int l, uint32_t g, const uint64_t* evk, int batch);
int fl_raw_mul_relin(fl_ctx* c, uint64_t* out, const uint64_t* a, const uint64_t* b, int l, const uint64_t* evk);   /* EvalMult(ct,ct) F.cpp:431 */
int fl_raw_mul_plain(fl_ctx* c, uint64_t* out, const uint64_t* ct, const uint64_t* pt, int l);   /* EvalMult(ct,pt) F.cpp:427 */

/* ---- limb-sharded key switch (optional multi-GPU mode, fhe_linformer_b200/sharded.py): the stages of the hybrid key switch
 * restricted to a contiguous limb range [first, first + count) of the extended basis Q_l u P.  Work buffers are full size:
 * dco [l][N], up [beta][l+K][N], acc [2][l+K][N], tq [2][l][N], out [2][l][N].  Between fl_raw_ks_pcoef and
 * fl_raw_ks_moddown the ranks exchange the P limbs of acc (coefficient form), after fl_raw_ks_moddown the limbs of out. ---- */
int fl_raw_ks_digits(fl_ctx* c, uint64_t* dco, const uint64_t* poly, int l);                                   /* scaled INTT of the digits */
int fl_raw_ks_digits_part(fl_ctx* c, uint64_t* dco, const uint64_t* poly, int l, int first, int count);        /* ... of limbs in range only */
int fl_raw_ks_modup(fl_ctx* c, uint64_t* up, const uint64_t* dco, int l, int first, int count);                /* ModUp of targets in range */
int fl_raw_ks_inner(fl_ctx* c, uint64_t* acc, const uint64_t* up, const uint64_t* poly, const uint64_t* evk, int l, int first, int count);
int fl_raw_ks_pcoef(fl_ctx* c, uint64_t* acc, int l, int first, int count);   /* first, count: range inside the K special limbs */
int fl_raw_ks_moddown(fl_ctx* c, uint64_t* out, uint64_t* tq, const uint64_t* acc, int l, int first, int count, const uint64_t* add0,
                      const uint64_t* add1, uint32_t g);

/* ---- the same operations with HOST buffers: H2D copy, kernels, D2H copy (what a host-resident caller pays) ---- */
int fl_host_ntt(fl_ctx* c, uint64_t* poly_host, int l, int inverse);
int fl_host_rotate(fl_ctx* c, uint64_t* out_host, const uint64_t* ct_host, int l, uint32_t g, const uint64_t* evk_dev);
int fl_host_rotate_batch(fl_ctx* c, uint64_t* out_host, const uint64_t* ct_host, int l, uint32_t g, const uint64_t* evk_dev, int batch);
/* The same call without the final wait: uploads, key switches and downloads of successive calls overlap on three streams (two
 * device staging slots used alternately); out_host is valid after fl_sync().  Buffers should be pinned (cudaHostAlloc /
 * cudaHostRegister), and neither buffer of a call may be reused before the second following call or fl_sync(). */
int fl_host_rotate_batch_async(fl_ctx* c, uint64_t* out_host, const uint64_t* ct_host, int l, uint32_t g, const uint64_t* evk_dev, int batch);
int fl_host_mul_relin(fl_ctx* c, uint64_t* out_host, const uint64_t* a_host, const uint64_t* b_host, int l, const uint64_t* evk_dev);

/* ================= scheme level: what FHEController's methods call on `context` ================= */
typedef struct fl_elem fl_elem;   /* ciphertext (2 polynomials) or plaintext (1), with level / noiseScaleDeg / scale / slots */
typedef fl_elem fl_ct;            /* Ctxt = Ciphertext<DCRTPoly>   FHEController.h:20 */
typedef fl_elem fl_pt;            /* Ptxt = Plaintext              FHEController.h:19 */

/* keys: KeyGen F.cpp:47, EvalMultKeyGen F.cpp:49, EvalRotateKeyGen F.cpp:248.
 * fl_keygen(c, 0) draws the secret key, the public polynomials, the key errors and all later encryption randomness from
 * ChaCha20 streams under independent keys taken from the operating system (getrandom), as OpenFHE's CSPRNG does; nothing
 * is derived from a seed and nothing seed-like is ever written to a file.  A non-zero seed is fl_keygen_seeded. */
int fl_keygen(fl_ctx* c, uint64_t seed_or_zero);
/* TEST ONLY: every key stream derives from `seed` (SplitMix64), so the CPU restatement reproduces the keys limb for limb.
 * Keys made this way are not secret: a public polynomial reveals the seed. */
int fl_keygen_seeded(fl_ctx* c, uint64_t seed);
int fl_gen_mult_key(fl_ctx* c);
int fl_gen_rot_keys(fl_ctx* c, const int* indices, int n);
int fl_gen_conj_key(fl_ctx* c);
int fl_keys_clear(fl_ctx* c, int kind);   /* 0: ClearEvalAutomorphismKeys F.cpp:336, 1: ClearEvalMultKeys F.cpp:342 */
int fl_num_rot_keys(fl_ctx* c);
double fl_rot_key_bytes(fl_ctx* c);      /* device memory held by the automorphism keys (dnum x 2 x (L + K) x N words each) */
/* raw key material <-> host (Serial::SerializeToFile / DeserializeFromFile of keys, F.cpp:59-89,192-220,251,291) */
int fl_export_sk(fl_ctx* c, uint64_t* out /* (L+K) N */);
int fl_export_pk(fl_ctx* c, uint64_t* out /* 2 L N */);
int fl_export_evk(fl_ctx* c, uint32_t galois /* 0 = mult key */, uint64_t* out);
int fl_import_keys(fl_ctx* c, const uint64_t* sk_or_null, const uint64_t* pk_or_null);
int fl_import_evk(fl_ctx* c, uint32_t galois, const uint64_t* evk);
int fl_keys_save(fl_ctx* c, const char* path);    /* one file: sk, pk, mult key, all automorphism keys */
/* what: 1 secret-key.txt | 2 public-key.txt | 4 mult-keys.txt | 8 rot_<name> (the four files of F.cpp:59-89,251) */
int fl_keys_save_sel(fl_ctx* c, const char* path, int what);
int fl_keys_load(fl_ctx* c, const char* path);   /* merges whatever records the file holds */

/* MakeCKKSPackedPlaintext(vec, 1, level, nullptr, slots) F.cpp:353; im may be NULL */
int fl_encode(fl_ctx* c, const double* re, const double* im, int n, int level, int slots, fl_pt** out);
/* `count` real vectors (n values each, row-major in re) -> ONE batched plaintext in one upload and five launches; element i is
   bit-identical to fl_encode(re + i n, NULL, n, ...).  The read_*_input loops of a forward (F.cpp:628-649 per file) go through it. */
int fl_encode_many(fl_ctx* c, const double* re, int count, int n, int level, int slots, fl_pt** out);
/* MakeCKKSPackedPlaintext + Encrypt of `count` real vectors in one pass (F.cpp:628-649 per input file): ONE batched ciphertext;
   message and error e0 share a forward transform, three transforms per ciphertext instead of four */
int fl_encrypt_values_many(fl_ctx* c, const double* re, int count, int n, int level, int slots, fl_ct** out);
int fl_encrypt(fl_ctx* c, const fl_pt* p, fl_ct** out);                         /* Encrypt F.cpp:380 */
/* the same for n plaintexts of one level in a dozen launches: *out is ONE batched operand (fl_batch_slice gives ciphertext i).
   n = 1 with a batched plaintext (fl_encode_many) encrypts every element of it. */
int fl_encrypt_many(fl_ctx* c, const fl_pt* const* pts, int n, fl_ct** out);
/* TEST ONLY: fixed encryption randomness (v, e0, e1 from SplitMix64(seed)) for the limb-exact parity tests */
int fl_encrypt_seeded(fl_ctx* c, const fl_pt* p, uint64_t seed, fl_ct** out);
int fl_decrypt(fl_ctx* c, const fl_ct* a, double* re, double* im, int slots);   /* Decrypt + GetRealPackedValue F.cpp:389,402 */
int fl_decode(fl_ctx* c, const fl_pt* p, double* re, double* im, int slots);

int fl_add(fl_ctx* c, const fl_elem* a, const fl_elem* b, fl_ct** out);         /* EvalAdd F.cpp:410,414 */
int fl_sub(fl_ctx* c, const fl_elem* a, const fl_elem* b, fl_ct** out);
int fl_add_many(fl_ctx* c, fl_ct* const* v, int n, fl_ct** out);                /* EvalAddMany F.cpp:418,1067 */
int fl_add_const(fl_ctx* c, const fl_ct* a, double k, fl_ct** out);
int fl_mul(fl_ctx* c, const fl_elem* a, const fl_elem* b, fl_ct** out);         /* EvalMult F.cpp:427,431 */
int fl_mul_const(fl_ctx* c, const fl_ct* a, double k, fl_ct** out);
int fl_mul_many(fl_ctx* c, fl_ct* const* v, int n, fl_ct** out);                /* EvalMultMany F.cpp:1297 */
/* EvalLinearWSum over a batched operand: out[o] = sum_t w[o * n_in + t] * in[t] (n_in = fl_elem_batch(in)); the result is a
 * batch of n_out ciphertexts.  Used for the encrypted Linformer E / F projection that the reference leaves to the client
 * (src/python/dimReduce.py:153-160). */
int fl_linear_wsum(fl_ctx* c, const fl_ct* in, const double* w, int n_out, fl_ct** out);
int fl_rotate(fl_ctx* c, const fl_ct* a, int k, fl_ct** out);                   /* EvalRotate F.cpp:435,833,843 */
int fl_has_rot_key(fl_ctx* c, int k);                                           /* is the EvalRotateKeyGen key for index k resident? */
/* out = sum_{t < 2^steps} rot(a, t * stride): FHEController::rotsum / rotsum_padded / repeat, F.cpp:829-867 */
/* rotation amounts fl_rotsum(steps, stride) uses when all of them have keys: the doubling steps stride * 2^i plus the extra
 * multiples of its hoisted groups (up to four doubling steps = 15 rotations share one ModUp / ModDown); returns their count.
 * Missing extra keys only make the ladder fall back to smaller groups. */
int fl_rotsum_rotations(int steps, int stride, int* out, int cap);
int fl_rotsum(fl_ctx* c, const fl_ct* a, int steps, int stride, fl_ct** out);
int fl_conjugate(fl_ctx* c, const fl_ct* a, fl_ct** out);
int fl_rescale(fl_ctx* c, const fl_ct* a, fl_ct** out);
int fl_eval_poly(fl_ctx* c, const fl_ct* a, const double* coeffs, int n, fl_ct** out);                       /* EvalPoly F.cpp:1291 */
/* EvalChebyshevFunction(f, ct, a, b, degree) F.cpp:486,1319-1335: coefficients from fl_chebyshev_coefficients */
int fl_eval_chebyshev(fl_ctx* c, const fl_ct* x, const double* coeffs, int n, double a, double b, fl_ct** out);
int fl_chebyshev_coefficients(double (*f)(double, void*), void* user, double a, double b, int degree, double* out /* degree+1 */);

/* BSGS diagonal ciphertext x plaintext matrix product (OpenFHE EvalLinearTransform, the kernel of CoeffsToSlots / SlotsToCoeffs
 * inside EvalBootstrap F.cpp:445; BASELINE.json north star: "the BSGS diagonal ciphertext x plaintext matmul behind the Linformer
 * E/F projections and the Q/K/V/FFN linears").  The matrix is given by its generalised diagonals:
 *   (M v)[p] = sum_k diag_k[p] * v[(p + shifts[k]) mod slots],   re / im: [ndiag][slots] (im may be NULL).
 * fl_lt_create plans baby / giant steps (max_baby = 0: ~sqrt of the diagonal span, <= 16 each) and, for level >= 0, encodes the
 * pre-rotated diagonals over Q_l u P at that level; fl_lt_rotations lists the rotation keys fl_lt_apply needs (returns their
 * count).  fl_lt_apply evaluates the product in one call with double hoisting (one ModUp for all baby steps, one ModDown per giant
 * step, one for all giant rotations); a batched operand is transformed as one.  The result has one more scaling degree, like
 * EvalMult(ct, pt).  fl_lt_apply_plain is the same plan as separate EvalRotate / EvalMult / EvalAdd calls (test checker). */
typedef struct fl_lt fl_lt;
int fl_lt_create(fl_ctx* c, const int* shifts, int ndiag, const double* re, const double* im, int slots, int level, int max_baby, fl_lt** out);
int fl_lt_rotations(fl_ctx* c, const fl_lt* t, int* out, int cap);
int fl_lt_shape(const fl_lt* t, int* n1, int* n2, int* stride, int* ndiag);
int fl_lt_apply(fl_ctx* c, fl_lt* t, const fl_ct* a, fl_ct** out);
int fl_lt_apply_plain(fl_ctx* c, fl_lt* t, const fl_ct* a, fl_ct** out);
void fl_lt_free(fl_lt* t);

/* EvalBootstrapSetup F.cpp:238,280, EvalBootstrapKeyGen F.cpp:239, EvalBootstrap F.cpp:445 */
int fl_bootstrap_setup(fl_ctx* c, int budget_cts, int budget_stc, int slots);
int fl_bootstrap_keygen(fl_ctx* c, int slots);
int fl_bootstrap(fl_ctx* c, const fl_ct* a, fl_ct** out);
int fl_bootstrap_iter(fl_ctx* c, const fl_ct* a, int iterations, int precision, fl_ct** out);   /* EvalBootstrap(c, 2, precision) F.cpp:461 */

/* Ciphertext / Plaintext accessors: GetLevel (M:231..), GetSlots F.cpp:448, Clone M:223 */
int fl_elem_level(const fl_elem* a);
int fl_elem_limbs(const fl_elem* a);
int fl_elem_deg(const fl_elem* a);
int fl_elem_slots(const fl_elem* a);
int fl_elem_ncomp(const fl_elem* a);
double fl_elem_scale(const fl_elem* a);
int fl_elem_clone(fl_ctx* c, const fl_elem* a, fl_elem** out);
/* Batched operands: n ciphertexts of identical level / degree / scale stored back to back.  fl_add / fl_mul (ct x pt) /
 * fl_rotate / fl_rotsum / fl_rescale accept a batched ciphertext and launch once per stage for the whole batch: the
 * independent iterations of the reference's `for (i < rows.size())` loops (F.cpp:872-1120) share kernel launches. */
int fl_batch_pack(fl_ctx* c, fl_elem* const* v, int n, fl_elem** out);
int fl_batch_slice(fl_ctx* c, const fl_elem* a, int i, fl_elem** out);   /* zero-copy view of element i */
int fl_batch_range(fl_ctx* c, const fl_elem* a, int first, int count, fl_elem** out);   /* zero-copy view of elements first .. first + count - 1 */
int fl_elem_batch(const fl_elem* a);
void fl_elem_free(fl_elem* a);
int fl_elem_export(fl_ctx* c, const fl_elem* a, uint64_t* host /* ncomp * limbs * N */);
int fl_elem_import(fl_ctx* c, const uint64_t* host, int ncomp, int limbs, int deg, double scale, int slots, fl_elem** out);
int fl_elem_save(fl_ctx* c, const fl_elem* a, const char* path);    /* Serial::SerializeToFile(ct) F.cpp:1361 */
int fl_elem_load(fl_ctx* c, const char* path, fl_elem** out);       /* Serial::DeserializeFromFile F.cpp:1386 */

/* per-entry-point GPU time (CUDA events on the engine stream around each call) and host time: "name calls gpu_ms host_ms" lines */
int fl_prof_enable(fl_ctx* c, int on);
int fl_prof_dump(fl_ctx* c, char* buf, size_t cap);

/* op ledger: algorithmic bytes per SURVEY.md section 8(d) */
int fl_ledger_enable(fl_ctx* c, int on);
int fl_ledger_reset(fl_ctx* c);
int fl_ledger_dump(fl_ctx* c, char* buf, size_t cap);   /* "op@limbs count bytes\n" lines */

#ifdef __cplusplus
}
#endif
#endif
