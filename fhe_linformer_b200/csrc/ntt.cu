// ntt.cu -- limb-batched negacyclic NTT / INTT for sm_100a  (kernel family K1 of SURVEY.md 2.1).
//
// Replaces the per-limb ForwardTransformToBitReverse / InverseTransformFromBitReverse calls OpenFHE
// makes inside every EvalMult / EvalRotate / rescale the reference issues
// (/root/reference/src/FHEController.cpp:426-436); conventions per SURVEY.md Appendix A.4:
// forward = Cooley-Tukey, natural -> bit-reversed order; inverse = Gentleman-Sande back to natural.
//
// Two passes per transform, one HBM round trip each (the second is L2-resident for one ciphertext):
//   column pass : the 4 widest-stride stages, radix-16 entirely in registers, one column per thread,
//                 adjacent threads on adjacent columns (fully coalesced, no shared memory);
//   chunk pass  : the remaining logN-4 stages on contiguous chunks of 2^(logN-4) words, one CTA per
//                 chunk, radix-8 register blocks exchanged through XOR-swizzled shared memory.
// Butterflies are Harvey lazy (values < 4q forward, < 2q inverse) with Shoup twiddles.
#include "device_ctx.h"
#include "modarith.cuh"

namespace flk {
namespace {

using namespace dev;

// forward radix-2^LOG block over e[0..2^LOG): twiddle index of group g at sub-stage s is (J<<s)+g
template <int LOG>
__device__ __forceinline__ void ct_block(u64* e, const u64* __restrict__ tw, const u64* __restrict__ tws, u32 J, u64 q) {
#pragma unroll
    for (int s = 0; s < LOG; ++s) {
        const int half = (1 << LOG) >> (s + 1);
#pragma unroll
        for (int g = 0; g < (1 << s); ++g) {
            const u64 w = __ldg(tw + ((J << s) + g)), ws = __ldg(tws + ((J << s) + g));
#pragma unroll
            for (int j = 0; j < half; ++j) ct_bfly(e[g * 2 * half + j], e[g * 2 * half + half + j], w, ws, q);
        }
    }
}
template <int LOG>
__device__ __forceinline__ void gs_block(u64* e, const u64* __restrict__ tw, const u64* __restrict__ tws, u32 J, u64 q) {
#pragma unroll
    for (int s = LOG - 1; s >= 0; --s) {
        const int half = (1 << LOG) >> (s + 1);
#pragma unroll
        for (int g = 0; g < (1 << s); ++g) {
            const u64 w = __ldg(tw + ((J << s) + g)), ws = __ldg(tws + ((J << s) + g));
#pragma unroll
            for (int j = 0; j < half; ++j) gs_bfly(e[g * 2 * half + j], e[g * 2 * half + half + j], w, ws, q);
        }
    }
}

// ---------------- column pass (register radix-16) ----------------
constexpr int R1 = 1 << kRadix1Log;

template <bool FWD>
__global__ void __launch_bounds__(256) ntt_column_kernel(u64* __restrict__ data, DevTables T, LimbSel sel, size_t batch_stride,
                                                         const u64* __restrict__ post, const u64* __restrict__ post_sh) {
    const int limb = blockIdx.y, m = sel.m[limb];
    const int cols = T.N >> kRadix1Log;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    u64* a = data + (size_t)blockIdx.z * batch_stride + (size_t)sel.pos[limb] * T.N + c;
    const u64 q = T.q[m];
    u64 e[R1];
#pragma unroll
    for (int k = 0; k < R1; ++k) e[k] = a[(size_t)k * cols];
    if (FWD) {
        ct_block<kRadix1Log>(e, T.tw + (size_t)m * T.N, T.tw_sh + (size_t)m * T.N, 1, q);
#pragma unroll
        for (int k = 0; k < R1; ++k) a[(size_t)k * cols] = e[k];   // lazy, < 4q
    } else {
        gs_block<kRadix1Log>(e, T.itw + (size_t)m * T.N, T.itw_sh + (size_t)m * T.N, 1, q);
        const u64 w = post ? post[m] : T.ninv[m], ws = post ? post_sh[m] : T.ninv_sh[m];
#pragma unroll
        for (int k = 0; k < R1; ++k) a[(size_t)k * cols] = mul_shoup(e[k], w, ws, q);
    }
}

// ---------------- chunk pass (shared-memory radix-8 rounds) ----------------
// swizzled shared-memory position: conflict-free for the stride patterns of every round
__device__ __forceinline__ int spos(int idx) { return (idx ^ ((idx >> 3) & 7)) + ((idx >> 6) << 3); }

template <int S2>
struct Sched {
    static constexpr int C = 1 << S2;
    static constexpr int NT = C / 8 < 1 ? 1 : C / 8;
    static constexpr int NR = (S2 + 2) / 3;
    static constexpr int last_log = S2 - 3 * (NR - 1);
    __host__ __device__ static constexpr int log_of(int r) { return r < NR - 1 ? 3 : last_log; }
    __host__ __device__ static constexpr int ulog_of(int r) { return r < NR - 1 ? S2 - 3 * (r + 1) : 0; }
};

// one round: each thread owns 8 elements = G groups of E = 2^LOG; element k of group (hi,lo) is at hi*E*u + lo + k*u
template <int LOG, int ULOG, bool FWD, bool FINAL>
__device__ __forceinline__ void chunk_round(u64* sm, int tid, u32 chunk, int logN, int S2, const u64* tw, const u64* tws, u64 q) {
    constexpr int E = 1 << LOG, G = 8 / E, U = 1 << ULOG;
#pragma unroll
    for (int h = 0; h < G; ++h) {
        const int gid = tid * G + h;
        const int lo = gid & (U - 1), hi = gid >> ULOG;
        const int base = hi * (E * U) + lo;
        u64 e[E];
#pragma unroll
        for (int k = 0; k < E; ++k) e[k] = sm[spos(base + k * U)];
        const u32 J = (1u << (logN - LOG - ULOG)) + (chunk << (S2 - LOG - ULOG)) + hi;
        if (FWD) ct_block<LOG>(e, tw, tws, J, q); else gs_block<LOG>(e, tw, tws, J, q);
#pragma unroll
        for (int k = 0; k < E; ++k) {
            u64 v = e[k];
            if (FINAL) { v = csub(v, q << 1); v = csub(v, q); }
            sm[spos(base + k * U)] = v;
        }
    }
}

template <int S2, bool FWD, int R>
struct Rounds {
    using S = Sched<S2>;
    __device__ static __forceinline__ void run(u64* sm, int tid, u32 chunk, int logN, const u64* tw, const u64* tws, u64 q) {
        // forward visits rounds 0..NR-1 (wide strides first); inverse visits NR-1..0
        constexpr int r = FWD ? R : S::NR - 1 - R;
        chunk_round<S::log_of(r), S::ulog_of(r), FWD, FWD && (R == S::NR - 1)>(sm, tid, chunk, logN, S2, tw, tws, q);
        __syncthreads();
        if constexpr (R + 1 < S::NR) Rounds<S2, FWD, R + 1>::run(sm, tid, chunk, logN, tw, tws, q);
    }
};

template <int S2, bool FWD>
__global__ void __launch_bounds__(Sched<S2>::NT) ntt_chunk_kernel(u64* __restrict__ data, DevTables T, LimbSel sel, size_t batch_stride) {
    using S = Sched<S2>;
    constexpr int C = S::C, NT = S::NT;
    __shared__ u64 sm[C + C / 8 + 8];
    const int limb = blockIdx.y, m = sel.m[limb], tid = threadIdx.x;
    const u32 chunk = blockIdx.x;
    u64* a = data + (size_t)blockIdx.z * batch_stride + (size_t)sel.pos[limb] * T.N + (size_t)chunk * C;
    const u64 q = T.q[m];
    const u64* tw = (FWD ? T.tw : T.itw) + (size_t)m * T.N;
    const u64* tws = (FWD ? T.tw_sh : T.itw_sh) + (size_t)m * T.N;
#pragma unroll
    for (int k = 0; k < C / NT; ++k) sm[spos(tid + k * NT)] = a[tid + k * NT];
    __syncthreads();
    Rounds<S2, FWD, 0>::run(sm, tid, chunk, T.logN, tw, tws, q);
#pragma unroll
    for (int k = 0; k < C / NT; ++k) a[tid + k * NT] = sm[spos(tid + k * NT)];
}

template <bool FWD>
void launch_chunk(const DevTables& t, u64* data, const LimbSel& sel, int batch, size_t bs, cudaStream_t s) {
    const int S2 = t.logN - kRadix1Log;
    dim3 grid(1u << kRadix1Log, sel.n, batch);
    switch (S2) {
#define FLK_CASE(X) case X: ntt_chunk_kernel<X, FWD><<<grid, Sched<X>::NT, 0, s>>>(data, t, sel, bs); break;
        FLK_CASE(6) FLK_CASE(7) FLK_CASE(8) FLK_CASE(9) FLK_CASE(10) FLK_CASE(11) FLK_CASE(12)
#undef FLK_CASE
        default: throw std::invalid_argument("unsupported ring dimension (logN must be 10..16)");
    }
}

template <bool FWD>
void launch_column(const DevTables& t, u64* data, const LimbSel& sel, int batch, size_t bs, const u64* post, const u64* post_sh,
                   cudaStream_t s) {
    const int cols = t.N >> kRadix1Log, threads = cols < 256 ? cols : 256;
    dim3 grid((cols + threads - 1) / threads, sel.n, batch);
    ntt_column_kernel<FWD><<<grid, threads, 0, s>>>(data, t, sel, bs, post, post_sh);
}

}  // namespace

void launch_ntt(const DevTables& t, u64* data, const LimbSel& sel, int batch, size_t batch_stride, cudaStream_t s) {
    if (sel.n == 0 || batch == 0) return;
    launch_column<true>(t, data, sel, batch, batch_stride, nullptr, nullptr, s);
    launch_chunk<true>(t, data, sel, batch, batch_stride, s);
    FLK_CUDA(cudaGetLastError());
}

void launch_intt(const DevTables& t, u64* data, const LimbSel& sel, int batch, size_t batch_stride, const u64* post,
                 const u64* post_sh, cudaStream_t s) {
    if (sel.n == 0 || batch == 0) return;
    launch_chunk<false>(t, data, sel, batch, batch_stride, s);
    launch_column<false>(t, data, sel, batch, batch_stride, post, post_sh, s);
    FLK_CUDA(cudaGetLastError());
}

}  // namespace flk
