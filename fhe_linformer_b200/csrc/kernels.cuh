// kernels.cuh -- launchers for the non-NTT kernels of the CKKS hot path (K2-K7 of SURVEY.md 2.1).
#pragma once
#include "device_ctx.h"

namespace flk {

enum class EwOp : int { Add = 0, Sub = 1, Mul = 2 };

// out = a (op) b, limb-wise.  b_poly_stride = 0 broadcasts one polynomial b over `polys` polynomials of a
// (ct x pt).  Buffers are [batch][polys][l][N]; limb i of every polynomial uses modulus sel.m[i].
void launch_ew(const DevTables& t, EwOp op, u64* out, const u64* a, const u64* b, const LimbSel& sel, int polys, int batch,
               size_t a_batch_stride, size_t b_batch_stride, size_t b_poly_stride, cudaStream_t s);
// out = a * b + c, limb-wise (one polynomial per batch element; out and a share a_batch_stride; b is broadcast when its stride is 0)
void launch_ew_muladd(const DevTables& t, u64* out, const u64* a, const u64* b, const u64* c, const LimbSel& sel, int batch, size_t a_batch_stride,
                      size_t b_batch_stride, size_t c_batch_stride, cudaStream_t s);

struct ScalarSet {   // per-limb multiplier with Shoup companion
    u64 c[64];
    u64 c_sh[64];
};
// out = a * c[limb]  (polys polynomials of sel.n limbs)
void launch_mul_scalar(const DevTables& t, u64* out, const u64* a, const ScalarSet& sc, const LimbSel& sel, int polys, cudaStream_t s);
// out = a + c[limb] broadcast to every slot of the evaluation representation
void launch_add_scalar(const DevTables& t, u64* out, const u64* a, const ScalarSet& sc, const LimbSel& sel, int batch, size_t batch_stride, cudaStream_t s);

// out[limb][j] = in[limb][map[j]]
void launch_automorph(u64* out, const u64* in, const uint32_t* map, int N, int limbs, cudaStream_t s);

// tensor product of two 2-component ciphertexts: d0 = a0 b0, d1 = a0 b1 + a1 b0, d2 = a1 b1
void launch_tensor(const DevTables& t, u64* d0, u64* d1, u64* d2, const u64* a, const u64* b, int l, int batch, size_t d_bs, size_t a_bs, size_t b_bs,
                   cudaStream_t s);

// ---- hybrid key switch pieces ----
struct KsLevel {          // device constants for key switching at l active limbs
    const u64* post;      // [T] by modulus: N^-1 * (Q_d/q_m)^-1 mod q_m   (INTT post-scale)
    const u64* post_sh;
    const u64* hm;        // [beta][alpha][l+K]: (Q_d/q_{d*alpha+i}) mod q_ext(t)
    const u64* hm30;      // the same constants times 2^30 (mod q_ext(t)): the high 30-bit half of a source residue multiplies these
    int l, beta, alpha;
};
struct MdConst {          // ModDown constants (level independent)
    const u64* post;      // [T] by modulus (P limbs): N^-1 * (P/p_k)^-1 mod p_k
    const u64* post_sh;
    const u64* phm;       // [K][L]: (P/p_k) mod q_i
    const u64* phm30;     // [K][L]: (P/p_k) 2^30 mod q_i
    const u64* pinv;      // [L] P^-1 mod q_i
    const u64* pinv_sh;
};

// A contiguous range of limbs a launch is restricted to (count < 0: all).  One GPU always takes everything; the limb-sharded
// key switch (sharded.py) gives every rank its own ranges of the extended basis.
struct LimbRange {
    int first, count;
};
constexpr LimbRange kAllLimbs{0, -1};

// All key-switch kernels take a batch of ciphertexts (same limb count, same key); *_bs are batch strides in words.
// up[b][d][t][N] (coefficient form) for all digits d and all extended limbs t outside digit d.
// dcoef = INTT(c) already scaled by KsLevel::post, [b][l][N].
void launch_modup_conv(const DevTables& t, const KsLevel& ks, u64* up, const u64* dcoef, int batch, size_t up_bs, size_t dco_bs, cudaStream_t s,
                       LimbRange targets = kAllLimbs);
// acc[b][{0,1}][t] = sum_d U_d[t] * evk_{b,a}[d][mod(t)];  U_d[t] = c_eval[b][t] inside digit d else up[b][d][t]
void launch_inner_product(const DevTables& t, const KsLevel& ks, u64* acc, const u64* up, const u64* c_eval, const u64* evk, int batch,
                          size_t acc_bs, size_t up_bs, size_t c_bs, cudaStream_t s, LimbRange targets = kAllLimbs);
// Several (operand, key, output) jobs of identical shape in ONE launch: the baby steps of a BSGS transform (one operand, up to 15
// keys: the operand's limb is fetched from HBM once) and its giant steps (one operand per key).
constexpr int kIpJobsMax = 16;
struct IpJobs {
    u64* acc[kIpJobsMax];
    const u64* up[kIpJobsMax];
    const u64* c[kIpJobsMax];
    const u64* evk[kIpJobsMax];
    int n;
};
void launch_inner_product_jobs(const DevTables& t, const KsLevel& ks, const IpJobs& jobs, int batch, size_t acc_bs, size_t up_bs, size_t c_bs,
                               cudaStream_t s, LimbRange limbs = kAllLimbs);
// hoisted multi-rotation: acc[b][{0,1}][t][j] = sum_k (sum_d U_d[b][t] evk_k[d][mod(t)])[map_k[j]], nk <= kHoistMax keys with their gather maps
void launch_inner_product_multi(const DevTables& t, const KsLevel& ks, u64* acc, const u64* up, const u64* c_eval, const u64* const* evks,
                                const uint32_t* const* maps, int nk, int batch, size_t acc_bs, size_t up_bs, size_t c_bs, cudaStream_t s);
// s0[b][i][j] = (self ? c0[b][i][j] : 0) + sum_k c0[b][i][map_k[j]]
void launch_gather_sum(const DevTables& t, u64* s0, const u64* c0, const uint32_t* const* maps, int nk, int l, int batch, size_t s0_bs, size_t c_bs,
                       bool self, cudaStream_t s);
// tq[b][p][i][N] (coefficient form), i < l, from the scaled INTT of the P part of `polys` accumulators.
// pcoef = [b][polys][K][N] with poly stride pstride and batch stride p_bs
void launch_moddown_conv(const DevTables& t, const MdConst& md, u64* tq, const u64* pcoef, size_t pstride, int l, int polys, int batch,
                         size_t tq_bs, size_t p_bs, cudaStream_t s, LimbRange targets = kAllLimbs);
// out[b][p][i][j] = ((acc[b][p][i] - tq[b][p][i]) * P^-1 + add_p[b][i])[map ? map[j] : j] + (plus ? plus[b][p][i][j] : 0)
constexpr int kHoistMax = 15;   // rotations sharing one ModUp / ModDown in a hoisted rotate-and-sum (four doubling steps)

// BSGS linear transform (engine.cu: Engine::linear_transform)
constexpr int kBsgsMax = 16;
struct BsgsArgs {
    const u64* accb[kBsgsMax];        // unpermuted key inner products of baby rotation i >= 1, [B][2][l+K][N]; null = baby not used
    const uint32_t* map[kBsgsMax];    // automorphism gather map of baby rotation i (unused for i = 0, the identity)
    uint32_t mask[kBsgsMax];          // per giant step: bit i set when diagonal (j, i) is present
    int n1, n2;
};
struct GatherArgs {
    const u64* src[kBsgsMax];
    const uint32_t* map[kBsgsMax];
    int n;
};
// W[(j B + b)] = sum_i pt[j][i] * u_i[b] in the extended basis; pts [n2][n1][l+K][N], pc = P * ct [B][2][l][N]
void launch_bsgs_inner(const DevTables& t, u64* W, const u64* pts, const u64* pc, const BsgsArgs& a, int l, int B, size_t acc_bs, size_t pc_bs,
                       cudaStream_t s);
// out[b][r] = sum_k src_k[b][r] gathered through map_k; rows of rows_per_poly limbs per polynomial
void launch_gather_multi(const DevTables& t, u64* out, const GatherArgs& g, int l, int rows_per_poly, int rows, int batch, size_t out_bs, size_t src_bs,
                         cudaStream_t s);

struct FinishArgs {
    u64* out; size_t out_bs;
    const u64* acc; size_t acc_ps, acc_bs;
    const u64* tq; size_t tq_bs;
    const u64* add0; size_t add0_bs;
    const u64* add1; size_t add1_bs;
    const u64* plus; size_t plus_bs;
};
void launch_moddown_finish(const DevTables& t, const MdConst& md, const FinishArgs& a, const uint32_t* map, int l, int polys, int batch,
                           cudaStream_t s, LimbRange limbs = kAllLimbs);
// The same epilogue folded into the LAST pass of the forward transform of the ModDown conversion (ntt.cu): a thread finishes its 8
// transformed words in registers -- (acc - word) P^-1 + addend -- and stores them through the INVERSE automorphism map (an
// aligned 64-byte block goes to an aligned 64-byte block: bit-reversed order keeps blocks of 2^b together), so the transformed
// conversion is never written to or re-read from HBM and the separate finish launch disappears.  a.tq is the work buffer of the
// transform (its first pass still runs in place there).
struct NttFinish {
    FinishArgs a;
    const uint32_t* imap;     // map of g^-1 (nullptr: no permutation)
    const u64* pinv;          // [L] P^-1 mod q_i and Shoup companions
    const u64* pinv_sh;
    int l, polys;
    // rescale: the operand of the transform is ONE coefficient-form polynomial per batch element modulo limb switch_mod
    // ([batch][N] at switch_src), switched to each limb's modulus by the first pass on load (tq is then written, not read, by that pass)
    const u64* switch_src = nullptr;
    int switch_mod = -1;
};
void launch_ntt_finish(const DevTables& t, u64* tq, int batch, size_t tq_bs, const NttFinish& f, cudaStream_t s);

// out[o] = sum_t k[o][t] in[t] over operands of `rows` limb rows each (2 l for a ciphertext, 2 l B for a batched one):
// in [n_in][rows][N], out [n_out][rows][N], k [n_out][n_in][l][2] = {residue, Shoup}
// in_ptrs (device array of n_in pointers), when given, replaces the contiguous `in`
void launch_lincomb(const DevTables& t, u64* out, const u64* in, const u64* k, int l, int rows, int n_in, int n_out, cudaStream_t s,
                    const u64* const* in_ptrs = nullptr);

// ---- rescale ----
struct RsConst {
    const u64* qlinv;     // [L][L]: q_r^-1 mod q_i
    const u64* qlinv_sh;
};
// (the centred switch of the dropped limb to every other modulus happens in the first pass of launch_ntt_finish: NttFinish::switch_src)
// (out[p][i] = (in[p][i] - NTT(tq[p][i])) * q_{l-1}^-1 is the ModDown finish with P = q_{l-1}: launch_ntt_finish)

// ModRaise / modulus switch: out[p][i][j] = centred(x[p][j] mod q_src) mod q_{sel.m[i]}   (x in coefficient form)
void launch_mod_switch(const DevTables& t, u64* out, const u64* x, int src_mod, const LimbSel& sel, int polys, cudaStream_t s);
// multiplication by the monomial X^(N/2) (every slot times i) in evaluation format: first half of each limb times zeta,
// second half times -zeta, zeta = psi^(N/2); sc holds zeta per limb
void launch_mul_i(const DevTables& t, u64* out, const u64* a, const ScalarSet& sc, const LimbSel& sel, int polys, cudaStream_t s);

// coefficients (signed, as int64 or int128 lo/hi) -> residues, [l][N] for moduli sel
void launch_reduce_i64(const DevTables& t, u64* out, const int64_t* coef, const LimbSel& sel, cudaStream_t s);
void launch_reduce_i128(const DevTables& t, u64* out, const int64_t* coef_lohi, const LimbSel& sel, cudaStream_t s);
// small signed int8 coefficients -> residues
void launch_reduce_i8(const DevTables& t, u64* out, const int8_t* coef, const LimbSel& sel, cudaStream_t s);

// ---- encode.cu: device-side encoding and samplers ----
// dst[sel.pos][N] <- residues of a ternary (kind 0) or discrete-Gaussian (kind 1) polynomial drawn from SplitMix64(seed)
void launch_sample_limbs(const DevTables& t, u64* dst, u64 seed, int kind, const LimbSel& sel, cudaStream_t s);
// dst limb i <- uniform residues mod q_{sel.m[i]} from SplitMix64(seeds[i])
void launch_uniform_limbs(const DevTables& t, u64* dst, const u64* seeds, const LimbSel& sel, cudaStream_t s);
// the production samplers: ChaCha20 key stream (chacha.cuh), block j of stream `nonce` per coefficient; uniform limb i uses stream nonce + i
struct ChaChaKey;
// (batch polynomials batch_stride words apart: polynomial z is stream nonce + z)
void launch_sample_limbs_csprng(const DevTables& t, u64* dst, const ChaChaKey& key, u64 nonce, int kind, const LimbSel& sel, cudaStream_t s, int batch = 1,
                                size_t batch_stride = 0, bool accumulate = false);
void launch_uniform_limbs_csprng(const DevTables& t, u64* dst, const ChaChaKey& key, u64 nonce, const LimbSel& sel, cudaStream_t s);
// special inverse FFT of (re, im)[slots] in place, then coefficient form of round(scale * values) in l limbs (not yet NTT'd);
// kext > 0 appends the residues modulo the first kext special limbs (plaintexts in the extended basis Q_l u P)
void launch_encode(const DevTables& t, u64* dst, double* re, double* im, int slots, double scale, int l, const uint32_t* rot, const double* cre,
                   const double* cim, cudaStream_t s, int kext = 0, int batch = 1);

}  // namespace flk
