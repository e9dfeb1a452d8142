// encode.cu -- device side of CKKS encoding and of the samplers (kernel family K8/K9 of SURVEY.md 2.1).
//
// Replaces the host work inside MakeCKKSPackedPlaintext / Encrypt / KeyGen that the reference pays on every call
// (/root/reference/src/FHEController.cpp:348-385; 193 + 28 encodings and 195 encryptions per forward at S = 129):
//   * special inverse FFT of the canonical embedding in double precision.  The operation order and the rounding of every
//     floating-point step are those of the host restatement (no FMA contraction: explicit __dmul_rn / __dadd_rn), so the
//     encoded limbs stay bit-identical to the oracle's;
//   * scale, round to nearest-even, exact conversion to a 128-bit integer and reduction into every RNS limb;
//   * ternary / discrete-Gaussian (CDT) / uniform sampling from counter-based SplitMix64 streams: output j of a stream is
//     mix(seed + (j + 1) * gamma), so every coefficient is generated independently and equals the sequential host stream.
#include "chacha.cuh"
#include "kernels.cuh"
#include "modarith.cuh"

namespace flk {
namespace {
using namespace dev;

constexpr int kThreads = 256;
constexpr u64 kGamma = 0x9E3779B97F4A7C15ull;

__device__ __forceinline__ u64 splitmix_at(u64 seed, u64 k) {   // k-th output (k >= 1) of SplitMix64 seeded with `seed`
    u64 z = seed + k * kGamma;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__constant__ u64 c_gauss_cdt[30] = FLK_GAUSS_CDT_INIT;   // part of the device image: nothing to upload, nothing to order

__device__ __forceinline__ int ternary_of(u64 r) { const u64 v = r % 3; return v == 2 ? -1 : (int)v; }
__device__ __forceinline__ int gauss_of(u64 u, u64 sign_bit) {
    int k = 0;
    while (k < 29 && u >= c_gauss_cdt[k]) ++k;
    return sign_bit ? -k : k;
}

// kind 0: uniform ternary {0, 1, -1};  kind 1: discrete Gaussian by cumulative-distribution table (two outputs per coefficient)
__global__ void __launch_bounds__(kThreads) sample_limbs_kernel(u64* __restrict__ dst, u64 seed, int kind, DevTables T, LimbSel sel) {
    const int j = blockIdx.x * kThreads + threadIdx.x, limb = blockIdx.y;
    if (j >= T.N) return;
    const int v = kind == 0 ? ternary_of(splitmix_at(seed, (u64)j + 1))
                            : gauss_of(splitmix_at(seed, 2 * (u64)j + 1), splitmix_at(seed, 2 * (u64)j + 2) & 1);
    const u64 q = T.q[sel.m[limb]];
    dst[(size_t)sel.pos[limb] * T.N + j] = v >= 0 ? (u64)v : q - (u64)(-v);
}
// the same two samplers from ChaCha20 key stream: coefficient j takes block j of stream `nonce`; the thread draws the small integer
// once and writes its residue into every limb (one key-stream block per coefficient, not per limb)
// accumulate: dst <- dst + sample (mod q) instead of dst <- sample (an error polynomial added to a message in coefficient form)
__global__ void __launch_bounds__(kThreads) sample_limbs_csprng_kernel(u64* __restrict__ dst, ChaChaKey key, u64 nonce, int kind, DevTables T,
                                                                       LimbSel sel, size_t batch_stride, int accumulate) {
    const int j = blockIdx.x * kThreads + threadIdx.x;
    if (j >= T.N) return;
    dst += (size_t)blockIdx.z * batch_stride;      // polynomial z of a batch is stream nonce + z
    uint32_t blk[16];
    chacha20_block(key, (u64)j, nonce + (u64)blockIdx.z, blk);
    const int v = kind == 0 ? ternary_of(chacha_u64(blk, 0)) : gauss_of(chacha_u64(blk, 0), chacha_u64(blk, 1) & 1);
    for (int limb = 0; limb < sel.n; ++limb) {
        const u64 q = T.q[sel.m[limb]];
        const u64 r = v >= 0 ? (u64)v : q - (u64)(-v);
        u64* p = dst + (size_t)sel.pos[limb] * T.N + j;
        *p = accumulate ? addmod(*p, r, q) : r;
    }
}

struct SeedSet {
    u64 s[kMaxLimbSel];
};
__global__ void __launch_bounds__(kThreads) uniform_limbs_kernel(u64* __restrict__ dst, SeedSet seeds, DevTables T, LimbSel sel) {
    const int j = blockIdx.x * kThreads + threadIdx.x, limb = blockIdx.y;
    if (j >= T.N) return;
    dst[(size_t)sel.pos[limb] * T.N + j] = splitmix_at(seeds.s[limb], (u64)j + 1) % T.q[sel.m[limb]];
}

// uniform residues from 128 key-stream bits each (bias below 2^-64); limb i is stream nonce + i
__global__ void __launch_bounds__(kThreads) uniform_limbs_csprng_kernel(u64* __restrict__ dst, ChaChaKey key, u64 nonce, DevTables T, LimbSel sel) {
    const int j = blockIdx.x * kThreads + threadIdx.x, limb = blockIdx.y;
    if (j >= T.N) return;
    uint32_t blk[16];
    chacha20_block(key, (u64)j, nonce + (u64)limb, blk);
    const int m = sel.m[limb];
    dst[(size_t)sel.pos[limb] * T.N + j] = barrett128(U128{chacha_u64(blk, 0), chacha_u64(blk, 1)}, T.q[m], T.mu_lo[m], T.mu_hi[m]);
}

// ---- special inverse FFT: stage `len` pairs (i + j, i + j + len/2) with twiddle zeta^(-5^j) ----
__device__ __forceinline__ void inv_butterfly(double& ar, double& ai, double& br, double& bi, double kr, double ki) {
    const double ur = __dadd_rn(ar, br), ui = __dadd_rn(ai, bi);
    const double wr = __dsub_rn(ar, br), wi = __dsub_rn(ai, bi);
    ar = ur; ai = ui;
    br = __dsub_rn(__dmul_rn(wr, kr), __dmul_rn(wi, ki));
    bi = __dadd_rn(__dmul_rn(wr, ki), __dmul_rn(wi, kr));
}
__device__ __forceinline__ uint32_t inv_twiddle_index(const uint32_t* __restrict__ rot, int j, int len, uint32_t m) {
    const uint32_t lenq = (uint32_t)len << 2;
    return (lenq - (rot[j] % lenq)) * (m / lenq);
}

// one stage in global memory (used while len exceeds the shared-memory block)
__global__ void __launch_bounds__(kThreads) fft_inv_stage_kernel(double* __restrict__ re, double* __restrict__ im, int n, int len,
                                                                 const uint32_t* __restrict__ rot, const double* __restrict__ cre,
                                                                 const double* __restrict__ cim) {
    const int t = blockIdx.x * kThreads + threadIdx.x;
    if (t >= n / 2) return;
    re += (size_t)blockIdx.y * n; im += (size_t)blockIdx.y * n;      // one slot vector of the batch per grid row
    const int lenh = len >> 1, j = t % lenh, i = (t / lenh) * len;
    const uint32_t idx = inv_twiddle_index(rot, j, len, 4u * n);
    inv_butterfly(re[i + j], im[i + j], re[i + j + lenh], im[i + j + lenh], cre[idx], cim[idx]);
}

// all remaining stages (len <= BLOCK) of one BLOCK-sized slice in shared memory
constexpr int kFftBlock = 2048;
__global__ void __launch_bounds__(kFftBlock / 2) fft_inv_block_kernel(double* __restrict__ re, double* __restrict__ im, int n, int first_len,
                                                                      const uint32_t* __restrict__ rot, const double* __restrict__ cre,
                                                                      const double* __restrict__ cim) {
    __shared__ double sr[kFftBlock], si[kFftBlock];
    re += (size_t)blockIdx.y * n; im += (size_t)blockIdx.y * n;
    const int base = blockIdx.x * first_len, t = threadIdx.x;
    for (int k = t; k < first_len; k += blockDim.x) { sr[k] = re[base + k]; si[k] = im[base + k]; }
    __syncthreads();
    for (int len = first_len; len >= 2; len >>= 1) {
        const int lenh = len >> 1;
        for (int b = t; b < first_len / 2; b += blockDim.x) {
            const int j = b % lenh, i = (b / lenh) * len;
            const uint32_t idx = inv_twiddle_index(rot, j, len, 4u * n);
            inv_butterfly(sr[i + j], si[i + j], sr[i + j + lenh], si[i + j + lenh], cre[idx], cim[idx]);
        }
        __syncthreads();
    }
    for (int k = t; k < first_len; k += blockDim.x) { re[base + k] = sr[k]; im[base + k] = si[k]; }
}

// exact integer value of an already rounded double as sign + 128-bit magnitude
__device__ __forceinline__ void double_to_u128(double v, bool& neg, u64& lo, u64& hi) {
    const long long bits = __double_as_longlong(v);
    neg = bits < 0;
    const int e = (int)((bits >> 52) & 0x7ff);
    const u64 mant = ((u64)bits & 0xFFFFFFFFFFFFFull) | (e ? (1ull << 52) : 0);
    const int sh = e - 1075;   // value = mant * 2^sh
    if (e == 0 || sh <= -64) { lo = hi = 0; return; }
    if (sh < 0) { lo = mant >> (-sh); hi = 0; }
    else if (sh == 0) { lo = mant; hi = 0; }
    else if (sh < 64) { lo = mant << sh; hi = mant >> (64 - sh); }
    else { lo = 0; hi = sh < 128 ? mant << (sh - 64) : 0; }
}

// coefficient i * gap <- rint(re[bitrev(i)] / n * scale), coefficient N/2 + i * gap <- same for im; reduced into every limb
__global__ void __launch_bounds__(kThreads) encode_finish_kernel(u64* __restrict__ dst, const double* __restrict__ re, const double* __restrict__ im,
                                                                 int slots, int log_slots, int gap, double scale, DevTables T, int l, int kext) {
    const int t = blockIdx.x * kThreads + threadIdx.x;
    if (t >= 2 * slots) return;
    re += (size_t)blockIdx.y * slots; im += (size_t)blockIdx.y * slots; dst += (size_t)blockIdx.y * (l + kext) * T.N;
    const int i = t < slots ? t : t - slots;
    const int src = log_slots ? (int)(__brev((unsigned)i) >> (32 - log_slots)) : 0;
    const double x = (t < slots ? re : im)[src];
    const double v = rint(__dmul_rn(__ddiv_rn(x, (double)slots), scale));
    bool neg; u64 lo, hi;
    double_to_u128(v, neg, lo, hi);
    const size_t pos = (size_t)(t < slots ? 0 : T.N / 2) + (size_t)i * gap;
    for (int k = 0; k < l + kext; ++k) {            // kext > 0: also the first kext special limbs (extended-basis plaintexts)
        const int m = k < l ? k : T.L + (k - l);
        const u64 q = T.q[m];
        u64 r = barrett128(U128{lo, hi}, q, T.mu_lo[m], T.mu_hi[m]);
        if (neg && r) r = q - r;
        dst[(size_t)k * T.N + pos] = r;
    }
}

inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace

void launch_sample_limbs(const DevTables& t, u64* dst, u64 seed, int kind, const LimbSel& sel, cudaStream_t s) {
    sample_limbs_kernel<<<dim3(cdiv(t.N, kThreads), sel.n), kThreads, 0, s>>>(dst, seed, kind, t, sel);
    FLK_CUDA(cudaGetLastError());
}
void launch_uniform_limbs(const DevTables& t, u64* dst, const u64* seeds, const LimbSel& sel, cudaStream_t s) {
    SeedSet ss;
    for (int i = 0; i < sel.n; ++i) ss.s[i] = seeds[i];
    uniform_limbs_kernel<<<dim3(cdiv(t.N, kThreads), sel.n), kThreads, 0, s>>>(dst, ss, t, sel);
    FLK_CUDA(cudaGetLastError());
}
void launch_sample_limbs_csprng(const DevTables& t, u64* dst, const ChaChaKey& key, u64 nonce, int kind, const LimbSel& sel, cudaStream_t s, int batch,
                                size_t batch_stride, bool accumulate) {
    sample_limbs_csprng_kernel<<<dim3(cdiv(t.N, kThreads), 1, batch), kThreads, 0, s>>>(dst, key, nonce, kind, t, sel, batch_stride, accumulate ? 1 : 0);
    FLK_CUDA(cudaGetLastError());
}
void launch_uniform_limbs_csprng(const DevTables& t, u64* dst, const ChaChaKey& key, u64 nonce, const LimbSel& sel, cudaStream_t s) {
    uniform_limbs_csprng_kernel<<<dim3(cdiv(t.N, kThreads), sel.n), kThreads, 0, s>>>(dst, key, nonce, t, sel);
    FLK_CUDA(cudaGetLastError());
}
// batch > 1: re / im hold `batch` slot vectors back to back (each `slots` long), dst `batch` plaintexts back to back
void launch_encode(const DevTables& t, u64* dst, double* re, double* im, int slots, double scale, int l, const uint32_t* rot, const double* cre,
                   const double* cim, cudaStream_t s, int kext, int batch) {
    int log_slots = 0;
    while ((1 << log_slots) < slots) ++log_slots;
    int len = slots;
    for (; len > kFftBlock; len >>= 1) fft_inv_stage_kernel<<<dim3(cdiv(slots / 2, kThreads), batch), kThreads, 0, s>>>(re, im, slots, len, rot, cre, cim);
    if (len >= 2) fft_inv_block_kernel<<<dim3(slots / len, batch), std::max(32, std::min(len / 2, kFftBlock / 2)), 0, s>>>(re, im, slots, len, rot, cre, cim);
    const int gap = (t.N / 2) / slots;
    if (gap > 1) FLK_CUDA(cudaMemsetAsync(dst, 0, (size_t)batch * (l + kext) * t.N * 8, s));
    encode_finish_kernel<<<dim3(cdiv((size_t)2 * slots, kThreads), batch), kThreads, 0, s>>>(dst, re, im, slots, log_slots, gap, scale, t, l, kext);
    FLK_CUDA(cudaGetLastError());
}

}  // namespace flk
