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
//                 chunk, radix-8 register blocks exchanged through XOR-swizzled shared memory, the
//                 next round's twiddles prefetched (16-byte {w, w'} loads) across the barrier; the
//                 unit-stride round reads / writes global memory directly with 16-byte accesses, and the
//                 barrier between two rounds only spans the threads that exchange values (CTA, then a
//                 named barrier per 64 threads, then a warp).
// Butterflies are Harvey lazy with Shoup twiddles whose quotient is estimated from three 32x32 partial products
// (dev::mulhi_lazy: low by at most 2), so a twiddle product lies in [0,4q).  Forward values grow by 4q per stage;
// for moduli below 2^56 (every Q limb) 65q < 2^64, so the forward transform carries NO conditional subtraction
// until one Barrett-style reduction at the very end; the 60-bit P limbs keep values below 8q.  The inverse keeps
// values below 4q with one conditional subtraction per butterfly.  Per butterfly: 9 IMAD-class + 9 IADD3-class SASS.
#include <cstdlib>

#include "device_ctx.h"
#include "modarith.cuh"

namespace flk {
namespace {

using namespace dev;

__device__ __forceinline__ bool is_wide(u64 q) { return (q >> 56) != 0; }

// a*w - qhat*q with qhat in [floor(a*ws/2^64) - 2, floor(a*ws/2^64)], as fused multiply-adds with nq = -q: result in [0,4q)
__device__ __forceinline__ u64 shoup_nq(u64 a, u64 w, u64 ws, u64 nq) { return shoup_lazy4(a, w, ws, nq); }

// forward: x,y < B  ->  < B + 4q  (narrow limbs: no reduction at all; wide limbs: B = 8q kept by one conditional subtraction)
template <bool WIDE>
__device__ __forceinline__ void ct_bf(u64& x, u64& y, ulonglong2 t, u64 nq, u64 q4) {
    u64 u = x;
    if (WIDE) u = csub(u, q4);
    const u64 v = shoup_nq(y, t.x, t.y, nq);
    x = u + v;
    y = u - v + q4;
}
// inverse: x,y in [0,4q) -> [0,4q)
__device__ __forceinline__ void gs_bf(u64& x, u64& y, ulonglong2 t, u64 nq, u64 q4) {
    const u64 s = csub(x + y, q4);
    const u64 d = x - y + q4;
    x = s;
    y = shoup_nq(d, t.x, t.y, nq);
}

// twiddles of one radix-2^LOG block: entry (1<<s)-1+g is table index (J<<s)+g
template <int LOG>
__device__ __forceinline__ void load_tw(ulonglong2* t, const ulonglong2* __restrict__ tab, u32 J) {
#pragma unroll
    for (int s = 0; s < LOG; ++s)
#pragma unroll
        for (int g = 0; g < (1 << s); ++g) t[(1 << s) - 1 + g] = __ldg(tab + ((J << s) + g));
}
template <int LOG, bool WIDE>
__device__ __forceinline__ void ct_block(u64* e, const ulonglong2* t, u64 nq, u64 q4) {
#pragma unroll
    for (int s = 0; s < LOG; ++s) {
        const int half = (1 << LOG) >> (s + 1);
#pragma unroll
        for (int g = 0; g < (1 << s); ++g)
#pragma unroll
            for (int j = 0; j < half; ++j) ct_bf<WIDE>(e[g * 2 * half + j], e[g * 2 * half + half + j], t[(1 << s) - 1 + g], nq, q4);
    }
}
template <int LOG>
__device__ __forceinline__ void gs_block(u64* e, const ulonglong2* t, u64 nq, u64 q4) {
#pragma unroll
    for (int s = LOG - 1; s >= 0; --s) {
        const int half = (1 << LOG) >> (s + 1);
#pragma unroll
        for (int g = 0; g < (1 << s); ++g)
#pragma unroll
            for (int j = 0; j < half; ++j) gs_bf(e[g * 2 * half + j], e[g * 2 * half + half + j], t[(1 << s) - 1 + g], nq, q4);
    }
}
// canonical residue of a lazily accumulated forward value
template <bool WIDE>
__device__ __forceinline__ u64 final_reduce(u64 v, u64 q, u64 q4, u64 nq, u64 qinv64) {
    if (WIDE) return csub(csub(csub(v, q4), q4 >> 1), q);   // v < 8q
    // v < 69q < 2^63: the quotient from the top words only, k = hi32((v >> 32) floor(2^64/q)) (floor(2^64/q) < 2^32 as q > 2^33),
    // is floor(v/q) or one less (dropping the low word of v loses < 2^-19, the floor of the reciprocal < 0.25)
    const u32 k = (u32)(((u64)(u32)(v >> 32) * (u32)qinv64) >> 32);
    return csub(v + (u64)k * nq, q);
}

// ---------------- column pass (register radix-16) ----------------
constexpr int R1 = 1 << kRadix1Log;

template <bool FWD>
__global__ void __launch_bounds__(256, 3) ntt_column_kernel(u64* __restrict__ data, DevTables T, LimbSel sel, size_t batch_stride,
                                                         const u64* __restrict__ post, const u64* __restrict__ post_sh) {
    // the 15 twiddles of these four stages are the same for every column of a limb: one copy per CTA in shared memory
    // (broadcast reads) instead of 60 registers per thread
    __shared__ ulonglong2 stw[R1];
    const int limb = blockIdx.y, m = sel.m[limb];
    const int cols = T.N >> kRadix1Log;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (threadIdx.x < R1 - 1) stw[threadIdx.x] = __ldg((FWD ? T.tw2 : T.itw2) + (size_t)m * T.N + 1 + threadIdx.x);   // table entries 1..15 in block order
    __syncthreads();
    if (c >= cols) return;
    u64* a = data + (size_t)blockIdx.z * batch_stride + (size_t)sel.pos[limb] * T.N + c;
    const u64 q = T.q[m], nq = 0 - q, q4 = q << 2;
    u64 e[R1];
#pragma unroll
    for (int k = 0; k < R1; ++k) e[k] = a[(size_t)k * cols];
    if (FWD) {
        if (is_wide(q)) ct_block<kRadix1Log, true>(e, stw, nq, q4); else ct_block<kRadix1Log, false>(e, stw, nq, q4);
#pragma unroll
        for (int k = 0; k < R1; ++k) a[(size_t)k * cols] = e[k];   // lazy: < 8q (wide) or < 21q (narrow)
    } else {
        gs_block<kRadix1Log>(e, stw, nq, q4);
        const u64 w = post ? post[m] : T.ninv[m], ws = post ? post_sh[m] : T.ninv_sh[m];
#pragma unroll
        for (int k = 0; k < R1; ++k) a[(size_t)k * cols] = mul_shoup(e[k], w, ws, q);
    }
}

// ---------------- chunk pass (shared-memory radix-8 rounds) ----------------
// Shared-memory position of logical index idx: spos(idx) = (idx ^ ((idx>>3)&7)) + ((idx>>6)<<3).
// Conflict-free for every round's stride pattern.  For idx = base | (k<<ULOG) with disjoint bit
// fields this splits into a per-thread part and compile-time constants:
struct SBase {
    int px, d0;   // px = base ^ ((base>>3)&7), d0 = (base>>6)<<3
    __device__ __forceinline__ explicit SBase(int base) : px(base ^ ((base >> 3) & 7)), d0((base >> 6) << 3) {}
    template <int KC>
    __device__ __forceinline__ int at() const { return (px ^ (KC ^ ((KC >> 3) & 7))) + d0 + ((KC >> 6) << 3); }
};

template <int S2>
struct Sched {
    static constexpr int C = 1 << S2;
    static constexpr int NT = C / 8;
    static constexpr int NR = (S2 + 2) / 3;
    static constexpr int last_log = S2 - 3 * (NR - 1);
    static constexpr int MINB = NT >= 512 ? 2 : (NT >= 256 ? 4 : 1);
    __host__ __device__ static constexpr int log_of(int r) { return r < NR - 1 ? 3 : last_log; }
    __host__ __device__ static constexpr int ulog_of(int r) { return r < NR - 1 ? S2 - 3 * (r + 1) : 0; }
};

// Geometry of round (LOG, ULOG): a thread owns G = 8>>LOG groups of E = 2^LOG elements; element k of group
// gid = tid*G + h sits at logical index hi*E*U + lo + k*U  (lo = gid mod U, hi = gid / U).
template <int LOG, int ULOG>
struct Geo {
    static constexpr int E = 1 << LOG, G = 8 / E, U = 1 << ULOG, TW = E - 1;
    __device__ static __forceinline__ int base(int tid, int h) {
        const int gid = tid * G + h;
        return (gid >> ULOG) * (E * U) + (gid & (U - 1));
    }
    __device__ static __forceinline__ u32 J(int tid, int h, u32 chunk, int logN, int S2) {
        const int gid = tid * G + h;
        return (1u << (logN - LOG - ULOG)) + (chunk << (S2 - LOG - ULOG)) + (u32)(gid >> ULOG);
    }
};

template <int LOG, int ULOG, int K>
__device__ __forceinline__ void lds_group(u64* e, const u64* sm, const SBase& b) {
    if constexpr (K < (1 << LOG)) {
        e[K] = sm[b.at<(K << ULOG)>()];
        lds_group<LOG, ULOG, K + 1>(e, sm, b);
    }
}
template <int LOG, int ULOG, int K>
__device__ __forceinline__ void sts_group(const u64* e, u64* sm, const SBase& b) {
    if constexpr (K < (1 << LOG)) {
        sm[b.at<(K << ULOG)>()] = e[K];
        sts_group<LOG, ULOG, K + 1>(e, sm, b);
    }
}

template <int S2, bool FWD, int R>
struct Rounds {
    using S = Sched<S2>;
    static constexpr int r = FWD ? R : S::NR - 1 - R;     // forward: wide strides first; inverse: narrow first
    static constexpr int LOG = S::log_of(r), ULOG = S::ulog_of(r);
    using G = Geo<LOG, ULOG>;
    static constexpr bool FIRST = R == 0, LAST = R == S::NR - 1;

    __device__ static __forceinline__ void load_twiddles(ulonglong2* t, const ulonglong2* tab, int tid, u32 chunk, int logN) {
#pragma unroll
        for (int h = 0; h < G::G; ++h) load_tw<LOG>(t + h * G::TW, tab, G::J(tid, h, chunk, logN, S2));
    }

    // tw: this round's twiddles (already in flight / loaded).  Data: forward round 0 reads global memory directly,
    // inverse last round writes global memory directly; everything else goes through swizzled shared memory.
    __device__ static __forceinline__ void run(u64* __restrict__ a, u64* sm, int tid, u32 chunk, int logN, const ulonglong2* tab,
                                               ulonglong2* tw, u64 q, u64 nq, u64 q4, u64 qinv64, bool wide) {
        u64 e[8];
        if constexpr (FWD && FIRST) {
#pragma unroll
            for (int k = 0; k < 8; ++k) e[k] = a[tid + k * S::NT];
        } else if constexpr (!FWD && FIRST) {
            // the inverse starts with the unit-stride round: a thread's 8 values are contiguous in global memory
            const ulonglong2* in = reinterpret_cast<const ulonglong2*>(a + tid * 8);
#pragma unroll
            for (int k = 0; k < 4; ++k) { const ulonglong2 v = in[k]; e[2 * k] = v.x; e[2 * k + 1] = v.y; }
        } else {
#pragma unroll
            for (int h = 0; h < G::G; ++h) lds_group<LOG, ULOG, 0>(e + h * G::E, sm, SBase(G::base(tid, h)));
        }
#pragma unroll
        for (int h = 0; h < G::G; ++h) {
            if (FWD) {
                if (wide) ct_block<LOG, true>(e + h * G::E, tw + h * G::TW, nq, q4);
                else ct_block<LOG, false>(e + h * G::E, tw + h * G::TW, nq, q4);
            } else {
                gs_block<LOG>(e + h * G::E, tw + h * G::TW, nq, q4);
            }
        }
        if constexpr (FWD && LAST) {
#pragma unroll
            for (int k = 0; k < 8; ++k) e[k] = wide ? final_reduce<true>(e[k], q, q4, nq, qinv64) : final_reduce<false>(e[k], q, q4, nq, qinv64);
        }
        if constexpr (!LAST) {
            // prefetch the next round's twiddles so their L2 latency overlaps the exchange and the barrier
            Rounds<S2, FWD, R + 1>::load_twiddles(tw, tab, tid, chunk, logN);
        }
        if constexpr (!FWD && LAST) {
#pragma unroll
            for (int k = 0; k < 8; ++k) a[tid + k * S::NT] = e[k];    // lazy < 4q, consumed by the column pass
        } else if constexpr (FWD && LAST) {
            // unit-stride round: the thread's 8 results are contiguous, four 16-byte stores
            ulonglong2* out = reinterpret_cast<ulonglong2*>(a + tid * 8);
#pragma unroll
            for (int k = 0; k < 4; ++k) out[k] = make_ulonglong2(e[2 * k], e[2 * k + 1]);
        } else {
#pragma unroll
            for (int h = 0; h < G::G; ++h) sts_group<LOG, ULOG, 0>(e + h * G::E, sm, SBase(G::base(tid, h)));
            // A round only exchanges values inside aligned groups of 2^max(ULOG, ULOG_next) threads (the thread that reads
            // index hi' U + lo' + k U/8 next round finds it written by threads (hi'/8) U + lo' + k U/8 of this round), so the
            // barrier shrinks with the stride: whole CTA, then a named barrier per group, then a warp.
            constexpr int NEXT_ULOG = Rounds<S2, FWD, R + 1>::ULOG;
            constexpr int SCOPE = 1 << (ULOG > NEXT_ULOG ? ULOG : NEXT_ULOG);
            if constexpr (SCOPE >= S::NT) __syncthreads();
            else if constexpr (SCOPE <= 32) __syncwarp();
            else asm volatile("bar.sync %0, %1;" ::"r"(1 + tid / SCOPE), "n"(SCOPE) : "memory");
        }
        if constexpr (!LAST) Rounds<S2, FWD, R + 1>::run(a, sm, tid, chunk, logN, tab, tw, q, nq, q4, qinv64, wide);
    }
};

template <int S2, bool FWD>
__global__ void __launch_bounds__(Sched<S2>::NT, Sched<S2>::MINB) ntt_chunk_kernel(u64* __restrict__ data, DevTables T, LimbSel sel,
                                                                                    size_t batch_stride) {
    using S = Sched<S2>;
    constexpr int C = S::C;
    __shared__ u64 sm[C + C / 8];
    const int limb = blockIdx.y, m = sel.m[limb], tid = threadIdx.x;
    const u32 chunk = blockIdx.x;
    u64* a = data + (size_t)blockIdx.z * batch_stride + (size_t)sel.pos[limb] * T.N + (size_t)chunk * C;
    const u64 q = T.q[m], nq = 0 - q, q4 = q << 2, qinv64 = T.mu_hi[m];
    const bool wide = is_wide(q);
    const ulonglong2* tab = (FWD ? T.tw2 : T.itw2) + (size_t)m * T.N;
    ulonglong2 tw[7];
    Rounds<S2, FWD, 0>::load_twiddles(tw, tab, tid, chunk, T.logN);
    Rounds<S2, FWD, 0>::run(a, sm, tid, chunk, T.logN, tab, tw, q, nq, q4, qinv64, wide);
}

template <int S2, bool FWD>
void launch_chunk_s(const DevTables& t, u64* data, const LimbSel& sel, dim3 grid, size_t bs, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(ntt_chunk_kernel<S2, FWD>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured = true;
    }
    ntt_chunk_kernel<S2, FWD><<<grid, Sched<S2>::NT, 0, s>>>(data, t, sel, bs);
}

template <bool FWD>
void launch_chunk(const DevTables& t, u64* data, const LimbSel& sel, int batch, size_t bs, cudaStream_t s) {
    const int S2 = t.logN - kRadix1Log;
    dim3 grid(1u << kRadix1Log, sel.n, batch);
    switch (S2) {
#define FLK_CASE(X) case X: launch_chunk_s<X, FWD>(t, data, sel, grid, bs, s); break;
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


// ======================= radix-16 two-kernel transform (logN >= 12) =======================
// N = 256 x R (R = 2^S2, S2 = logN - 8).  Every thread works on radix-16 register blocks (32 butterflies on 16 values), so a
// value is touched by one global load, one shared-memory exchange and one global store per kernel:
//   head kernel : the 8 widest-stride stages on a tile of 256 rows x 16 adjacent columns.  Round A takes rows 16k + r
//                 (k = 0..15) per thread, round B rows 16r + j after a 32 KiB shared-memory exchange; both rounds move
//                 whole 128-byte row segments per half-warp.  The 15 round-A twiddles are the same for every thread.
//   tail kernel : the remaining S2 stages inside each contiguous row, 4096 contiguous words per CTA.  Round A strides
//                 16..R/2 (elements 16k + i), round B the last four stages on 16 contiguous words per thread; the exchange
//                 buffer is padded by one word per 16 (address 17 b + j) so both rounds are bank-conflict free.
// The inverse runs the same rounds backwards with Gentleman-Sande blocks (tail first, then head with the n^-1 scaling).
constexpr int kTile = 16;

template <bool FWD>
__global__ void __launch_bounds__(256, 2) ntt_head_kernel(u64* __restrict__ data, DevTables T, LimbSel sel, size_t batch_stride,
                                                          const u64* __restrict__ post, const u64* __restrict__ post_sh) {
    __shared__ u64 tile[256 * kTile];
    const int limb = blockIdx.y, m = sel.m[limb];
    const size_t cols = (size_t)(T.N >> 8);
    const int c = threadIdx.x & 15, r = threadIdx.x >> 4;
    u64* a = data + (size_t)blockIdx.z * batch_stride + (size_t)sel.pos[limb] * T.N + (size_t)blockIdx.x * kTile + c;
    const u64 q = T.q[m], nq = 0 - q, q4 = q << 2;
    const ulonglong2* tab = (FWD ? T.tw2 : T.itw2) + (size_t)m * T.N;
    u64 e[16];
    ulonglong2 t[15];
    if (FWD) {
        load_tw<4>(t, tab, 1);
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = a[(size_t)(16 * k + r) * cols];
        if (is_wide(q)) ct_block<4, true>(e, t, nq, q4); else ct_block<4, false>(e, t, nq, q4);
#pragma unroll
        for (int k = 0; k < 16; ++k) tile[(16 * k + r) * kTile + c] = e[k];
        load_tw<4>(t, tab, 16 + r);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 16; ++j) e[j] = tile[(16 * r + j) * kTile + c];
        if (is_wide(q)) ct_block<4, true>(e, t, nq, q4); else ct_block<4, false>(e, t, nq, q4);
#pragma unroll
        for (int j = 0; j < 16; ++j) a[(size_t)(16 * r + j) * cols] = e[j];   // lazy: < 8q (wide) or < 33q (narrow)
    } else {
        load_tw<4>(t, tab, 16 + r);
#pragma unroll
        for (int j = 0; j < 16; ++j) e[j] = a[(size_t)(16 * r + j) * cols];   // < 4q from the tail kernel
        gs_block<4>(e, t, nq, q4);
#pragma unroll
        for (int j = 0; j < 16; ++j) tile[(16 * r + j) * kTile + c] = e[j];
        load_tw<4>(t, tab, 1);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = tile[(16 * k + r) * kTile + c];
        gs_block<4>(e, t, nq, q4);
        const u64 w = post ? post[m] : T.ninv[m], ws = post ? post_sh[m] : T.ninv_sh[m];
#pragma unroll
        for (int k = 0; k < 16; ++k) a[(size_t)(16 * k + r) * cols] = mul_shoup(e[k], w, ws, q);
    }
}

template <int S2, bool FWD>
__global__ void __launch_bounds__(256, 2) ntt_tail_kernel(u64* __restrict__ data, DevTables T, LimbSel sel, size_t batch_stride) {
    constexpr int R = 1 << S2, LOGA = S2 - 4, EA = 1 << LOGA, GA = 16 / EA, RP = R + R / 16;
    __shared__ u64 sm[4096 + 256];
    const int limb = blockIdx.y, m = sel.m[limb], tid = threadIdx.x;
    const u32 chunk = blockIdx.x;
    u64* a = data + (size_t)blockIdx.z * batch_stride + (size_t)sel.pos[limb] * T.N + (size_t)chunk * 4096;
    const u64 q = T.q[m], nq = 0 - q, q4 = q << 2, qinv64 = T.mu_hi[m];
    const bool wide = is_wide(q);
    const ulonglong2* tab = (FWD ? T.tw2 : T.itw2) + (size_t)m * T.N;
    const u32 jb = (1u << (T.logN - 4)) + chunk * 256u + (u32)tid;   // round B: block of 16 contiguous words
    u64 e[16];
    if (FWD) {
        // round A: GA groups of EA elements 16k + i of one row
        if constexpr (LOGA > 0) {
            ulonglong2 t[GA * (EA - 1)];
#pragma unroll
            for (int h = 0; h < GA; ++h) {
                const int g = h * 256 + tid, row = g >> 4, i = g & 15;
                load_tw<LOGA>(t + h * (EA - 1), tab, 256u + chunk * (4096 / R) + (u32)row);
#pragma unroll
                for (int k = 0; k < EA; ++k) e[h * EA + k] = a[row * R + 16 * k + i];
            }
#pragma unroll
            for (int h = 0; h < GA; ++h) {
                if (wide) ct_block<LOGA, true>(e + h * EA, t + h * (EA - 1), nq, q4); else ct_block<LOGA, false>(e + h * EA, t + h * (EA - 1), nq, q4);
            }
        } else {
#pragma unroll
            for (int h = 0; h < 16; ++h) e[h] = a[h * 256 + tid];
        }
#pragma unroll
        for (int h = 0; h < GA; ++h) {
            const int g = h * 256 + tid, row = g >> 4, i = g & 15;
#pragma unroll
            for (int k = 0; k < EA; ++k) sm[row * RP + 17 * k + i] = e[h * EA + k];
        }
        ulonglong2 tb[15];
        load_tw<4>(tb, tab, jb);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 16; ++j) e[j] = sm[17 * tid + j];
        if (wide) ct_block<4, true>(e, tb, nq, q4); else ct_block<4, false>(e, tb, nq, q4);
        ulonglong2* o = reinterpret_cast<ulonglong2*>(a + 16 * tid);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            ulonglong2 v;
            v.x = wide ? final_reduce<true>(e[2 * j], q, q4, nq, qinv64) : final_reduce<false>(e[2 * j], q, q4, nq, qinv64);
            v.y = wide ? final_reduce<true>(e[2 * j + 1], q, q4, nq, qinv64) : final_reduce<false>(e[2 * j + 1], q, q4, nq, qinv64);
            o[j] = v;
        }
    } else {
        ulonglong2 tb[15];
        load_tw<4>(tb, tab, jb);
        const ulonglong2* in = reinterpret_cast<const ulonglong2*>(a + 16 * tid);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const ulonglong2 v = in[j]; e[2 * j] = v.x; e[2 * j + 1] = v.y; }
        gs_block<4>(e, tb, nq, q4);
#pragma unroll
        for (int j = 0; j < 16; ++j) sm[17 * tid + j] = e[j];
        if constexpr (LOGA > 0) {
            ulonglong2 t[GA * (EA - 1)];
#pragma unroll
            for (int h = 0; h < GA; ++h) load_tw<LOGA>(t + h * (EA - 1), tab, 256u + chunk * (4096 / R) + (u32)((h * 256 + tid) >> 4));
            __syncthreads();
#pragma unroll
            for (int h = 0; h < GA; ++h) {
                const int g = h * 256 + tid, row = g >> 4, i = g & 15;
#pragma unroll
                for (int k = 0; k < EA; ++k) e[h * EA + k] = sm[row * RP + 17 * k + i];
                gs_block<LOGA>(e + h * EA, t + h * (EA - 1), nq, q4);
#pragma unroll
                for (int k = 0; k < EA; ++k) a[row * R + 16 * k + i] = e[h * EA + k];   // lazy < 4q, consumed by the head kernel
            }
        } else {
            __syncthreads();
#pragma unroll
            for (int h = 0; h < 16; ++h) { const int g = h * 256 + tid; a[g] = sm[(g >> 4) * RP + (g & 15)]; }
        }
    }
}

template <bool FWD>
void launch_head(const DevTables& t, u64* data, const LimbSel& sel, int batch, size_t bs, const u64* post, const u64* post_sh, cudaStream_t s) {
    dim3 grid((t.N >> 8) / kTile, sel.n, batch);
    ntt_head_kernel<FWD><<<grid, 256, 0, s>>>(data, t, sel, bs, post, post_sh);
}
template <bool FWD>
void launch_tail(const DevTables& t, u64* data, const LimbSel& sel, int batch, size_t bs, cudaStream_t s) {
    dim3 grid(t.N / 4096, sel.n, batch);
    switch (t.logN - 8) {
#define FLK_CASE(X) case X: ntt_tail_kernel<X, FWD><<<grid, 256, 0, s>>>(data, t, sel, bs); break;
        FLK_CASE(4) FLK_CASE(5) FLK_CASE(6) FLK_CASE(7) FLK_CASE(8)
#undef FLK_CASE
        default: throw std::invalid_argument("radix-16 transform: logN must be 12..16");
    }
}
// The radix-16 pair executes 25 % fewer instructions per butterfly (23 vs 31) but holds 128 registers per thread (4 warps per
// sub-partition) and measured no faster on B200 (993 vs 1022 GB/s at N = 2^16, 2.37 s vs 2.22 s for the forward at N = 2^15):
// it stays selectable (FLK_NTT_RADIX16=1) as the base for the asynchronous-prefetch version, the default is the
// column + chunk pair.
bool use_radix16(const DevTables& t) {
    static const bool on = [] { const char* e = std::getenv("FLK_NTT_RADIX16"); return e && e[0] == '1'; }();
    return on && t.logN >= 12;
}

}  // namespace

void launch_ntt(const DevTables& t, u64* data, const LimbSel& sel, int batch, size_t batch_stride, cudaStream_t s) {
    if (sel.n == 0 || batch == 0) return;
    if (use_radix16(t)) {
        launch_head<true>(t, data, sel, batch, batch_stride, nullptr, nullptr, s);
        launch_tail<true>(t, data, sel, batch, batch_stride, s);
    } else {
        launch_column<true>(t, data, sel, batch, batch_stride, nullptr, nullptr, s);
        launch_chunk<true>(t, data, sel, batch, batch_stride, s);
    }
    FLK_CUDA(cudaGetLastError());
}

void launch_intt(const DevTables& t, u64* data, const LimbSel& sel, int batch, size_t batch_stride, const u64* post,
                 const u64* post_sh, cudaStream_t s) {
    if (sel.n == 0 || batch == 0) return;
    if (use_radix16(t)) {
        launch_tail<false>(t, data, sel, batch, batch_stride, s);
        launch_head<false>(t, data, sel, batch, batch_stride, post, post_sh, s);
    } else {
        launch_chunk<false>(t, data, sel, batch, batch_stride, s);
        launch_column<false>(t, data, sel, batch, batch_stride, post, post_sh, s);
    }
    FLK_CUDA(cudaGetLastError());
}

}  // namespace flk
