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
#include <atomic>
#include <cstdlib>

#include "device_ctx.h"
#include "kernels.cuh"
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
                                                         const u64* __restrict__ post, const u64* __restrict__ post_sh,
                                                         const u64* __restrict__ xsrc = nullptr, int xsrc_mod = -1) {
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
    if (FWD && xsrc) {
        // forward transform of a polynomial given in coefficient form modulo ANOTHER limb (rescale: the dropped limb): the centred
        // switch to this limb's modulus happens on load, so the switched copy is never written to or read from HBM
        const u64 qs = T.q[xsrc_mod], half = qs >> 1, ml = T.mu_lo[m], mh = T.mu_hi[m];
        const u64 qs_here = barrett128(U128{qs, 0}, q, ml, mh);
        const u64* x = xsrc + (size_t)blockIdx.z * T.N + c;
#pragma unroll
        for (int k = 0; k < R1; ++k) {
            const u64 v0 = x[(size_t)k * cols];
            u64 v = barrett128(U128{v0, 0}, q, ml, mh);
            if (v0 > half) v = submod(v, qs_here, q);
            e[k] = v;
        }
    } else {
#pragma unroll
        for (int k = 0; k < R1; ++k) e[k] = a[(size_t)k * cols];
    }
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

// what the last round of a forward chunk pass does with its 8 canonical results per thread
struct StoreEpi {            // plain transform: four 16-byte stores
    static constexpr bool kFinish = false;
};
struct FinishEpi {           // ModDown finish (kernels.cuh NttFinish): the results stay in registers
    static constexpr bool kFinish = true;
    u64* out; const u64* acc; const u64* add; const u64* plus; const uint32_t* imap;
    u64 q, pinv, pinv_sh;
    const u64* s_acc; const u64* s_add;     // this thread's 8 accumulator / addend words in the CTA's shared-memory staging area
    u32 mbar;                               // shared-memory address of the mbarrier the bulk copies complete on
    // The operands of the finish do not depend on the transform: the chunk's accumulator and addend words (contiguous, C words each)
    // are fetched by the TMA engine -- one 1-D bulk copy each (cp.async.bulk, UBLKCP), issued by one thread before the first round and
    // completing on an mbarrier -- so their HBM latency is spent under twelve stages of butterflies and no thread issues a load for them.
    __device__ __forceinline__ void prefetch(u64* sa, u64* sd, u64* bar, size_t chunk_first, int chunk_words, int tid) {
        s_acc = sa + tid * 8; s_add = sd + tid * 8;
        mbar = (u32)__cvta_generic_to_shared(bar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            const u32 bytes = (u32)chunk_words * 8u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(add ? 2u * bytes : bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((u32)__cvta_generic_to_shared(sa)),
                         "l"(acc + chunk_first), "r"(bytes), "r"(mbar) : "memory");
            if (add)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((u32)__cvta_generic_to_shared(sd)),
                             "l"(add + chunk_first), "r"(bytes), "r"(mbar) : "memory");
        }
    }
    __device__ __forceinline__ void store(const u64* e, size_t p) const {
        asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(mbar) : "memory");
        // two words at a time: the chunk kernel lives in 64 registers
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(s_acc + k);
            uint2 to = imap ? *reinterpret_cast<const uint2*>(imap + p + k) : make_uint2((uint32_t)p + k, (uint32_t)p + k + 1);
            u64 v0 = mul_shoup(submod(a.x, e[k], q), pinv, pinv_sh, q), v1 = mul_shoup(submod(a.y, e[k + 1], q), pinv, pinv_sh, q);
            if (add) {
                const ulonglong2 d = *reinterpret_cast<const ulonglong2*>(s_add + k);
                v0 = addmod(v0, d.x, q); v1 = addmod(v1, d.y, q);
            }
            if (plus) { v0 = addmod(v0, plus[to.x], q); v1 = addmod(v1, plus[to.y], q); }
            out[to.x] = v0; out[to.y] = v1;
        }
    }
};

template <int S2, bool FWD, int R, bool WIDE, class EPI = StoreEpi>
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
                                               ulonglong2* tw, u64 q, u64 nq, u64 q4, u64 qinv64, const EPI& epi = EPI(), const u64* __restrict__ src = nullptr) {
        u64 e[8];
        if constexpr (FWD && FIRST) {
#pragma unroll
            for (int k = 0; k < 8; ++k) e[k] = a[tid + k * S::NT];
        } else if constexpr (!FWD && FIRST) {
            // the inverse starts with the unit-stride round: a thread's 8 values are contiguous in global memory
            // (out of place when src is given: the digits of a key switch are read where the ciphertext lies)
            const ulonglong2* in = reinterpret_cast<const ulonglong2*>((src ? src : a) + tid * 8);
#pragma unroll
            for (int k = 0; k < 4; ++k) { const ulonglong2 v = in[k]; e[2 * k] = v.x; e[2 * k + 1] = v.y; }
        } else {
#pragma unroll
            for (int h = 0; h < G::G; ++h) lds_group<LOG, ULOG, 0>(e + h * G::E, sm, SBase(G::base(tid, h)));
        }
#pragma unroll
        for (int h = 0; h < G::G; ++h) {
            if (FWD) {
                ct_block<LOG, WIDE>(e + h * G::E, tw + h * G::TW, nq, q4);
            } else {
                gs_block<LOG>(e + h * G::E, tw + h * G::TW, nq, q4);
            }
        }
        if constexpr (FWD && LAST) {
#pragma unroll
            for (int k = 0; k < 8; ++k) e[k] = final_reduce<WIDE>(e[k], q, q4, nq, qinv64);
        }
        if constexpr (!LAST) {
            // prefetch the next round's twiddles so their L2 latency overlaps the exchange and the barrier
            Rounds<S2, FWD, R + 1, WIDE, EPI>::load_twiddles(tw, tab, tid, chunk, logN);
        }
        if constexpr (!FWD && LAST) {
#pragma unroll
            for (int k = 0; k < 8; ++k) a[tid + k * S::NT] = e[k];    // lazy < 4q, consumed by the column pass
        } else if constexpr (FWD && LAST) {
            if constexpr (EPI::kFinish) {
                epi.store(e, (size_t)chunk * S::C + (size_t)tid * 8);
            } else {
                // unit-stride round: the thread's 8 results are contiguous, four 16-byte stores
                ulonglong2* out = reinterpret_cast<ulonglong2*>(a + tid * 8);
#pragma unroll
                for (int k = 0; k < 4; ++k) out[k] = make_ulonglong2(e[2 * k], e[2 * k + 1]);
            }
        } else {
#pragma unroll
            for (int h = 0; h < G::G; ++h) sts_group<LOG, ULOG, 0>(e + h * G::E, sm, SBase(G::base(tid, h)));
            // A round only exchanges values inside aligned groups of 2^max(ULOG, ULOG_next) threads (the thread that reads
            // index hi' U + lo' + k U/8 next round finds it written by threads (hi'/8) U + lo' + k U/8 of this round), so the
            // barrier shrinks with the stride: whole CTA, then a named barrier per group, then a warp.
            constexpr int NEXT_ULOG = Rounds<S2, FWD, R + 1, WIDE, EPI>::ULOG;
            constexpr int SCOPE = 1 << (ULOG > NEXT_ULOG ? ULOG : NEXT_ULOG);
            if constexpr (SCOPE >= S::NT) __syncthreads();
            else if constexpr (SCOPE <= 32) __syncwarp();
            else asm volatile("bar.sync %0, %1;" ::"r"(1 + tid / SCOPE), "n"(SCOPE) : "memory");
        }
        if constexpr (!LAST) Rounds<S2, FWD, R + 1, WIDE, EPI>::run(a, sm, tid, chunk, logN, tab, tw, q, nq, q4, qinv64, epi, src);
    }
};

template <int S2, bool FWD>
__global__ void __launch_bounds__(Sched<S2>::NT, Sched<S2>::MINB) ntt_chunk_kernel(u64* __restrict__ data, DevTables T, LimbSel sel,
                                                                                    size_t batch_stride, const u64* __restrict__ src, size_t src_bs) {
    using S = Sched<S2>;
    constexpr int C = S::C;
    __shared__ u64 sm[C + C / 8];
    const int limb = blockIdx.y, m = sel.m[limb], tid = threadIdx.x;
    const u32 chunk = blockIdx.x;
    u64* a = data + (size_t)blockIdx.z * batch_stride + (size_t)sel.pos[limb] * T.N + (size_t)chunk * C;
    const u64 q = T.q[m], nq = 0 - q, q4 = q << 2, qinv64 = T.mu_hi[m];
    const ulonglong2* tab = (FWD ? T.tw2 : T.itw2) + (size_t)m * T.N;
    ulonglong2 tw[7];
    Rounds<S2, FWD, 0, false>::load_twiddles(tw, tab, tid, chunk, T.logN);
    // the limb's width (60-bit P limbs keep a conditional subtraction per butterfly) is uniform per CTA: two straight-line bodies
    // instead of a branch inside every round.  The inverse does not depend on it.
    // inverse only: read the first round from src (same limb slots, its own batch stride), everything else works on data
    const u64* sa = (!FWD && src) ? src + (size_t)blockIdx.z * src_bs + (size_t)sel.pos[limb] * T.N + (size_t)chunk * C : nullptr;
    if (FWD && is_wide(q)) Rounds<S2, FWD, 0, true>::run(a, sm, tid, chunk, T.logN, tab, tw, q, nq, q4, qinv64);
    else Rounds<S2, FWD, 0, false>::run(a, sm, tid, chunk, T.logN, tab, tw, q, nq, q4, qinv64, StoreEpi(), sa);
}

// chunk pass of the forward transform of the ModDown conversion with the finish folded into its last round (Q limbs only: narrow
// path unless the first modulus is wide)
template <int S2>
__global__ void __launch_bounds__(Sched<S2>::NT, Sched<S2>::MINB) ntt_chunk_finish_kernel(u64* __restrict__ data, DevTables T, size_t batch_stride, NttFinish f) {
    using S = Sched<S2>;
    constexpr int C = S::C;
    extern __shared__ __align__(16) u64 dsm[];    // [C + C/8] exchange buffer, [C] accumulator words, [C] addend words, mbarrier
    u64* sm = dsm;
    const int slot = blockIdx.y, poly = slot / f.l, i = slot - poly * f.l, tid = threadIdx.x, b = blockIdx.z;
    const u32 chunk = blockIdx.x;
    u64* a = data + (size_t)b * batch_stride + (size_t)slot * T.N + (size_t)chunk * C;
    const u64 q = T.q[i], nq = 0 - q, q4 = q << 2, qinv64 = T.mu_hi[i];
    const ulonglong2* tab = T.tw2 + (size_t)i * T.N;
    FinishEpi epi;
    const size_t lo = (size_t)i * T.N;
    epi.out = f.a.out + (size_t)b * f.a.out_bs + (size_t)slot * T.N;
    epi.acc = f.a.acc + (size_t)b * f.a.acc_bs + (size_t)poly * f.a.acc_ps + lo;
    const u64* addp = poly == 0 ? f.a.add0 : f.a.add1;
    epi.add = addp ? addp + (size_t)b * (poly == 0 ? f.a.add0_bs : f.a.add1_bs) + lo : nullptr;
    epi.plus = f.a.plus ? f.a.plus + (size_t)b * f.a.plus_bs + (size_t)slot * T.N : nullptr;
    epi.imap = f.imap; epi.q = q; epi.pinv = f.pinv[i]; epi.pinv_sh = f.pinv_sh[i];
    epi.prefetch(dsm + (C + C / 8), dsm + (C + C / 8) + C, dsm + (C + C / 8) + 2 * C, (size_t)chunk * C, C, tid);
    ulonglong2 tw[7];
    Rounds<S2, true, 0, false, FinishEpi>::load_twiddles(tw, tab, tid, chunk, T.logN);
    if (is_wide(q)) Rounds<S2, true, 0, true, FinishEpi>::run(a, sm, tid, chunk, T.logN, tab, tw, q, nq, q4, qinv64, epi);
    else Rounds<S2, true, 0, false, FinishEpi>::run(a, sm, tid, chunk, T.logN, tab, tw, q, nq, q4, qinv64, epi);
}

template <int S2, bool FWD>
void launch_chunk_s(const DevTables& t, u64* data, const LimbSel& sel, dim3 grid, size_t bs, cudaStream_t s, const u64* src, size_t src_bs) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(ntt_chunk_kernel<S2, FWD>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured = true;
    }
    ntt_chunk_kernel<S2, FWD><<<grid, Sched<S2>::NT, 0, s>>>(data, t, sel, bs, src, src_bs);
}

template <bool FWD>
void launch_chunk(const DevTables& t, u64* data, const LimbSel& sel, int batch, size_t bs, cudaStream_t s, const u64* src = nullptr, size_t src_bs = 0) {
    const int S2 = t.logN - kRadix1Log;
    dim3 grid(1u << kRadix1Log, sel.n, batch);
    switch (S2) {
#define FLK_CASE(X) case X: launch_chunk_s<X, FWD>(t, data, sel, grid, bs, s, src, src_bs); break;
        FLK_CASE(6) FLK_CASE(7) FLK_CASE(8) FLK_CASE(9) FLK_CASE(10) FLK_CASE(11) FLK_CASE(12)
#undef FLK_CASE
        default: throw std::invalid_argument("unsupported ring dimension (logN must be 10..16)");
    }
}

template <bool FWD>
void launch_column(const DevTables& t, u64* data, const LimbSel& sel, int batch, size_t bs, const u64* post, const u64* post_sh,
                   cudaStream_t s, const u64* xsrc = nullptr, int xsrc_mod = -1) {
    const int cols = t.N >> kRadix1Log, threads = cols < 256 ? cols : 256;
    dim3 grid((cols + threads - 1) / threads, sel.n, batch);
    ntt_column_kernel<FWD><<<grid, threads, 0, s>>>(data, t, sel, bs, post, post_sh, xsrc, xsrc_mod);
}


}  // namespace

void launch_ntt(const DevTables& t, u64* data, const LimbSel& sel, int batch, size_t batch_stride, cudaStream_t s) {
    if (sel.n == 0 || batch == 0) return;
    launch_column<true>(t, data, sel, batch, batch_stride, nullptr, nullptr, s);
    launch_chunk<true>(t, data, sel, batch, batch_stride, s);
    FLK_CUDA(cudaGetLastError());
}

void launch_ntt_finish(const DevTables& t, u64* tq, int batch, size_t tq_bs, const NttFinish& f, cudaStream_t s) {
    if (batch == 0) return;
    LimbSel sq;
    for (int i = 0; i < f.polys * f.l; ++i) sq.push(i % f.l, i);
    launch_column<true>(t, tq, sq, batch, tq_bs, nullptr, nullptr, s, f.switch_src, f.switch_mod);
    const int S2 = t.logN - kRadix1Log;
    const dim3 grid(1u << kRadix1Log, f.polys * f.l, batch);
    switch (S2) {
#define FLK_CASE(X) case X: { constexpr size_t shm = (size_t)((1 << X) + (1 << X) / 8 + 2 * (1 << X) + 2) * 8;                                                      \
                              static std::atomic<unsigned long long> cfg{0};   /* the attribute belongs to the device: one bit per device */             \
                              int dev = 0; FLK_CUDA(cudaGetDevice(&dev));                                                                                  \
                              if (!((cfg.load() >> (dev & 63)) & 1ull)) {                                                                                  \
                                  FLK_CUDA(cudaFuncSetAttribute(ntt_chunk_finish_kernel<X>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));       \
                                  cudaFuncSetAttribute(ntt_chunk_finish_kernel<X>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
                                  cfg.fetch_or(1ull << (dev & 63)); }                                                                                      \
                              ntt_chunk_finish_kernel<X><<<grid, Sched<X>::NT, shm, s>>>(tq, t, tq_bs, f); } break;
        FLK_CASE(6) FLK_CASE(7) FLK_CASE(8) FLK_CASE(9) FLK_CASE(10) FLK_CASE(11) FLK_CASE(12)
#undef FLK_CASE
        default: throw std::invalid_argument("unsupported ring dimension (logN must be 10..16)");
    }
    FLK_CUDA(cudaGetLastError());
}

void launch_intt(const DevTables& t, u64* data, const LimbSel& sel, int batch, size_t batch_stride, const u64* post,
                 const u64* post_sh, cudaStream_t s, const u64* src, size_t src_bs) {
    if (sel.n == 0 || batch == 0) return;
    launch_chunk<false>(t, data, sel, batch, batch_stride, s, src, src_bs);
    launch_column<false>(t, data, sel, batch, batch_stride, post, post_sh, s);
    FLK_CUDA(cudaGetLastError());
}

}  // namespace flk
