// kernels.cu -- element-wise, automorphism, base-conversion, inner-product, ModDown and rescale kernels
// (K2-K7 of SURVEY.md 2.1) for sm_100a.  All are HBM-streaming integer kernels: limb-major layout,
// adjacent threads on adjacent coefficients, 128-bit accumulators reduced once.
#include <atomic>
#include <cstdlib>
#include "kernels.cuh"
#include "modarith.cuh"

namespace flk {
namespace {
using namespace dev;

constexpr int kThreads = 256;
constexpr int kAlphaMax = 8;

// ---------------- element-wise ----------------
template <int OP>
__global__ void __launch_bounds__(kThreads) ew_kernel(u64* __restrict__ out, const u64* __restrict__ a, const u64* __restrict__ b, DevTables T,
                                                      LimbSel sel, int polys, size_t a_bs, size_t b_bs, size_t b_ps) {
    const size_t per_poly = (size_t)sel.n * T.N;
    const size_t e = ((size_t)blockIdx.x * kThreads + threadIdx.x) * 2;
    if (e >= per_poly * polys) return;
    const int p = (int)(e / per_poly);
    const size_t r = e - (size_t)p * per_poly;
    const int limb = (int)(r >> T.logN), m = sel.m[limb];
    const u64 q = T.q[m];
    const ulonglong2 x = *reinterpret_cast<const ulonglong2*>(a + blockIdx.y * a_bs + e);
    const ulonglong2 y = *reinterpret_cast<const ulonglong2*>(b + blockIdx.y * b_bs + (size_t)p * b_ps + r);
    ulonglong2 z;
    if (OP == 0) { z.x = addmod(x.x, y.x, q); z.y = addmod(x.y, y.y, q); }
    else if (OP == 1) { z.x = submod(x.x, y.x, q); z.y = submod(x.y, y.y, q); }
    else { const u64 ml = T.mu_lo[m], mh = T.mu_hi[m]; z.x = mulmod(x.x, y.x, q, ml, mh); z.y = mulmod(x.y, y.y, q, ml, mh); }
    *reinterpret_cast<ulonglong2*>(out + blockIdx.y * a_bs + e) = z;
}

__global__ void __launch_bounds__(kThreads) ew_muladd_kernel(u64* __restrict__ out, const u64* __restrict__ a, const u64* __restrict__ b,
                                                             const u64* __restrict__ c, DevTables T, LimbSel sel, size_t a_bs, size_t b_bs, size_t c_bs) {
    const size_t per_poly = (size_t)sel.n * T.N;
    const size_t e = ((size_t)blockIdx.x * kThreads + threadIdx.x) * 2;
    if (e >= per_poly) return;
    const int m = sel.m[(int)(e >> T.logN)];
    const u64 q = T.q[m], ml = T.mu_lo[m], mh = T.mu_hi[m];
    const ulonglong2 x = *reinterpret_cast<const ulonglong2*>(a + blockIdx.y * a_bs + e);
    const ulonglong2 y = *reinterpret_cast<const ulonglong2*>(b + blockIdx.y * b_bs + e);
    const ulonglong2 z = *reinterpret_cast<const ulonglong2*>(c + blockIdx.y * c_bs + e);
    *reinterpret_cast<ulonglong2*>(out + blockIdx.y * a_bs + e) = make_ulonglong2(addmod(mulmod(x.x, y.x, q, ml, mh), z.x, q), addmod(mulmod(x.y, y.y, q, ml, mh), z.y, q));
}

__global__ void __launch_bounds__(kThreads) mul_scalar_kernel(u64* __restrict__ out, const u64* __restrict__ a, DevTables T, LimbSel sel,
                                                              ScalarSet sc, int polys) {
    const size_t per_poly = (size_t)sel.n * T.N;
    const size_t e = ((size_t)blockIdx.x * kThreads + threadIdx.x) * 2;
    if (e >= per_poly * polys) return;
    const size_t r = e % per_poly;
    const int limb = (int)(r >> T.logN);
    const u64 q = T.q[sel.m[limb]], c = sc.c[limb], cs = sc.c_sh[limb];
    const ulonglong2 x = *reinterpret_cast<const ulonglong2*>(a + e);
    ulonglong2 z; z.x = mul_shoup(x.x, c, cs, q); z.y = mul_shoup(x.y, c, cs, q);
    *reinterpret_cast<ulonglong2*>(out + e) = z;
}

__global__ void __launch_bounds__(kThreads) add_scalar_kernel(u64* __restrict__ out, const u64* __restrict__ a, DevTables T, LimbSel sel,
                                                              ScalarSet sc, size_t bs) {
    const size_t e = ((size_t)blockIdx.x * kThreads + threadIdx.x) * 2;
    if (e >= (size_t)sel.n * T.N) return;
    out += blockIdx.y * bs; a += blockIdx.y * bs;
    const int limb = (int)(e >> T.logN);
    const u64 q = T.q[sel.m[limb]], c = sc.c[limb];
    const ulonglong2 x = *reinterpret_cast<const ulonglong2*>(a + e);
    ulonglong2 z; z.x = addmod(x.x, c, q); z.y = addmod(x.y, c, q);
    *reinterpret_cast<ulonglong2*>(out + e) = z;
}

__global__ void __launch_bounds__(kThreads) automorph_kernel(u64* __restrict__ out, const u64* __restrict__ in, const uint32_t* __restrict__ map,
                                                             int N) {
    const int j = blockIdx.x * kThreads + threadIdx.x;
    if (j >= N) return;
    const size_t o = (size_t)blockIdx.y * N;
    out[o + j] = in[o + map[j]];
}

__global__ void __launch_bounds__(kThreads) tensor_kernel(u64* __restrict__ d0, u64* __restrict__ d1, u64* __restrict__ d2,
                                                          const u64* __restrict__ a, const u64* __restrict__ b, DevTables T, int l, size_t d_bs,
                                                          size_t a_bs, size_t b_bs) {
    const size_t e = (size_t)blockIdx.x * kThreads + threadIdx.x, pl = (size_t)l * T.N;
    if (e >= pl) return;
    d0 += blockIdx.y * d_bs; d1 += blockIdx.y * d_bs; d2 += blockIdx.y * d_bs; a += blockIdx.y * a_bs; b += blockIdx.y * b_bs;
    const int m = (int)(e >> T.logN);
    const u64 q = T.q[m], ml = T.mu_lo[m], mh = T.mu_hi[m];
    const u64 a0 = a[e], a1 = a[pl + e], b0 = b[e], b1 = b[pl + e];
    d0[e] = mulmod(a0, b0, q, ml, mh);
    U128 x{0, 0}; mad128(x, a0, b1); mad128(x, a1, b0);
    d1[e] = barrett128(x, q, ml, mh);
    d2[e] = mulmod(a1, b1, q, ml, mh);
}

// ---------------- fast basis conversion (ModUp / ModDown) ----------------
__device__ __forceinline__ RedC load_redc(const DevTables& T, int m) {
    const ulonglong2* p = reinterpret_cast<const ulonglong2*>(T.redc) + (size_t)m * 4;
    const ulonglong2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
    return RedC{a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
}

// One thread per coefficient: the A source residues stay in registers (as 30-bit halves) while the thread walks its strip
// of targets; products accumulate carry-free (dev::mac3, 4 IMAD.WIDE each), one reduction per output (dev::reduce3).
// sh[i * ext + t] = (S/s_i) mod target_t as 30-bit halves (zero for the missing sources of a short last digit).
// grid: (N / 256, target groups, batch * beta); buffers carry a batch stride.
template <int A>
__global__ void __launch_bounds__(kThreads) modup_conv_kernel(u64* __restrict__ up, const u64* __restrict__ dcoef, DevTables T, KsLevel ks,
                                                              size_t up_bs, size_t dco_bs, LimbRange rg) {
    extern __shared__ uint4 shs[];
    const int d = blockIdx.z % ks.beta, b = blockIdx.z / ks.beta, l = ks.l, ext = l + T.K;
    const int lo = d * A, hi = min(lo + A, l), ns = hi - lo;
    for (int i = threadIdx.x; i < A * ext; i += kThreads) {
        const Split30 h = split30(ks.hm[(size_t)d * A * ext + i]), g = split30(ks.hm30[(size_t)d * A * ext + i]);
        shs[i] = make_uint4(h.lo, h.hi, g.lo, g.hi);
    }
    __syncthreads();
    const int j = (blockIdx.x * kThreads + threadIdx.x) * 2;   // two adjacent coefficients per thread: 16-byte accesses, and the
    if (j >= T.N) return;                                        // conversion constants of a target are fetched once for both
    Split30 y0[A], y1[A];
#pragma unroll
    for (int i = 0; i < A; ++i) {
        const ulonglong2 v = i < ns ? *reinterpret_cast<const ulonglong2*>(dcoef + (size_t)b * dco_bs + (size_t)(lo + i) * T.N + j) : make_ulonglong2(0, 0);
        y0[i] = split30(v.x); y1[i] = split30(v.y);
    }
    const int tg = gridDim.y, per = (rg.count + tg - 1) / tg;   // targets rg.first .. rg.first + rg.count - 1 (all of them on one GPU)
    const int t0 = rg.first + blockIdx.y * per, t1 = min(t0 + per, rg.first + rg.count);
    u64* dst = up + (size_t)b * up_bs + (size_t)d * ext * T.N + j;
    for (int t = t0; t < t1; ++t) {
        if (t >= lo && t < hi) continue;
        const RedC rc = load_redc(T, t < l ? t : T.L + (t - l));
        Acc2 a0{0, 0}, a1{0, 0};
#pragma unroll
        for (int i = 0; i < A; ++i) {
            const uint4 h = shs[i * ext + t];
            mac2(a0, y0[i], h);
            mac2(a1, y1[i], h);
        }
        // the forward NTT that follows takes lazily reduced operands (< 8q)
        *reinterpret_cast<ulonglong2*>(dst + (size_t)t * T.N) = make_ulonglong2(reduce2_lazy(a0.b0, a0.b1, rc), reduce2_lazy(a1.b0, a1.b1, rc));
    }
}

// acc{0,1}[b][t] = sum_d U_d[b][t] * evk_{b,a}[d][mod(t)].  A thread owns two adjacent coefficients (16-byte accesses) of one
// extended limb for IPB ciphertexts of the batch, so every evaluation-key word it loads (the largest stream of a key switch)
// is used IPB times.  The digit loop is unrolled (BETA is a template parameter) so all loads of a thread are in flight together.
constexpr int kIpb = 8;
template <int BETA>
__global__ void __launch_bounds__(kThreads) inner_product_kernel(IpJobs jobs, DevTables T, KsLevel ks, int batch, size_t acc_bs, size_t up_bs, size_t c_bs,
                                                                 int t_first) {
    // grid (coefficients, jobs x batch groups, limbs): CTAs are issued x, then y, then z, so all batch groups of one limb run back to
    // back and that limb's key words (2 beta x N, streamed from HBM once) are served from L2 to every group after the first.  Several
    // jobs (operand, key, output) share one launch -- the baby steps of a BSGS transform multiply ONE extended operand by up to 15 keys:
    // with the job index fastest in y, that operand's limb is read from HBM once and from L2 by the other keys.
    const int job = blockIdx.y % jobs.n;
    u64* __restrict__ acc = jobs.acc[job];
    const u64* __restrict__ up = jobs.up[job];
    const u64* __restrict__ c_eval = jobs.c[job];
    const u64* __restrict__ evk = jobs.evk[job];
    const int t = t_first + blockIdx.z, l = ks.l, ext = l + T.K, b0 = (blockIdx.y / jobs.n) * kIpb;
    const int j = (blockIdx.x * kThreads + threadIdx.x) * 2;
    if (j >= T.N) return;
    const int m = t < l ? t : T.L + (t - l);
    const size_t kpoly = (size_t)(T.L + T.K) * T.N;
    const int own_d = t < l ? t / ks.alpha : -1;
    const int nb = min(kIpb, batch - b0);
    ulonglong2 k0[BETA], k1[BETA], u[BETA], un[BETA];
    const u64* src[BETA];
    size_t sbs[BETA];
#pragma unroll
    for (int d = 0; d < BETA; ++d) {
        const u64* kb = evk + (size_t)d * 2 * kpoly + (size_t)m * T.N + j;
        k0[d] = __ldg(reinterpret_cast<const ulonglong2*>(kb));
        k1[d] = __ldg(reinterpret_cast<const ulonglong2*>(kb + kpoly));
        src[d] = d == own_d ? c_eval + (size_t)t * T.N + j : up + ((size_t)d * ext + t) * T.N + j;
        sbs[d] = d == own_d ? c_bs : up_bs;
        un[d] = *reinterpret_cast<const ulonglong2*>(src[d] + (size_t)b0 * sbs[d]);
    }
    const RedC rc = load_redc(T, m);
    for (int i = 0; i < nb; ++i) {
#pragma unroll
        for (int d = 0; d < BETA; ++d) {
            u[d] = un[d];
            if (i + 1 < nb) un[d] = *reinterpret_cast<const ulonglong2*>(src[d] + (size_t)(b0 + i + 1) * sbs[d]);   // next ciphertext's operands
        }
        Acc3 s0x{0, 0, 0}, s0y{0, 0, 0}, s1x{0, 0, 0}, s1y{0, 0, 0};
#pragma unroll
        for (int d = 0; d < BETA; ++d) {
            const Split30 ux = split30(u[d].x), uy = split30(u[d].y);
            mac3(s0x, ux, split30(k0[d].x)); mac3(s0y, uy, split30(k0[d].y));
            mac3(s1x, ux, split30(k1[d].x)); mac3(s1y, uy, split30(k1[d].y));
        }
        u64* o = acc + (size_t)(b0 + i) * acc_bs + (size_t)t * T.N + j;
        *reinterpret_cast<ulonglong2*>(o) = make_ulonglong2(reduce3(s0x, rc), reduce3(s0y, rc));
        *reinterpret_cast<ulonglong2*>(o + (size_t)ext * T.N) = make_ulonglong2(reduce3(s1x, rc), reduce3(s1y, rc));
    }
}

// ---- hoisted multi-rotation (rotate-and-sum ladders): acc[b][{0,1}][t][j] = sum_k IP_k(U[b])[t][map_k[j]] ----
// One ModUp serves nk rotations of the same ciphertext; the automorphism of rotation k is applied while accumulating
// (gather at map_k[j]), so a single ModDown finishes all of them.  Per key the digit products accumulate carry-free and
// are reduced once; the nk reduced values add up modulo q.
struct MultiKeys {
    const u64* evk[kHoistMax];
    const uint32_t* map[kHoistMax];
    int n;
};
template <int BETA, int KPR_MAX>
__global__ void __launch_bounds__(kThreads, KPR_MAX > 1 ? 4 : 1) inner_product_multi_kernel(u64* __restrict__ acc, const u64* __restrict__ up, const u64* __restrict__ c_eval,
                                                                       MultiKeys mk, DevTables T, KsLevel ks, int batch, size_t acc_bs, size_t up_bs,
                                                                       size_t c_bs) {
    constexpr int IPB = 2;
    // grid (coefficients, batch groups, limbs): for one limb the nk keys' words (nk x 2 beta x N) and the limb of every ciphertext's
    // extended digits stay in L2 while all batch groups gather from them (the batch used to be the slowest grid dimension: every
    // group re-streamed the keys from HBM, 15 GB per launch at batch 64)
    const int t = blockIdx.z, l = ks.l, ext = l + T.K, b0 = blockIdx.y * IPB;
    const int j = blockIdx.x * kThreads + threadIdx.x;
    if (j >= T.N) return;
    const int m = t < l ? t : T.L + (t - l);
    const size_t kpoly = (size_t)(T.L + T.K) * T.N;
    const int own_d = t < l ? t / ks.alpha : -1;
    const RedC rc = load_redc(T, m);
    // The carry-free accumulators take 8 products before they must be reduced, so the digit products of KPR = 8 / BETA keys
    // (gathered at different source positions, but summed into the same output) share one reduction -- the reduction costs
    // more instructions than a key's products.
    constexpr int KPR_WIDE = (BETA >= 8 ? 1 : 8 / BETA) < KPR_MAX ? (BETA >= 8 ? 1 : 8 / BETA) : KPR_MAX;
    // Below 2^56 (every Q limb) the halves of a residue are 30 + 26 bits: the low accumulator still takes 2^60 per product, so 15
    // products fit it together with the fold of the high one (15 2^60 + 4q < 2^64), the other two stay far below their bounds --
    // 15 / BETA keys share a reduction there instead of 8 / BETA (the reductions were 39 % of this kernel's instructions).
    constexpr int KPR_NARROW = KPR_MAX == 1 ? 1 : (BETA >= 15 ? 1 : 15 / BETA);
    const int KPR = (rc.q >> 56) != 0 ? KPR_WIDE : KPR_NARROW;
    u64 r0[IPB], r1[IPB];
    Acc3 s0[IPB], s1[IPB];
#pragma unroll
    for (int i = 0; i < IPB; ++i) { r0[i] = r1[i] = 0; s0[i] = Acc3{0, 0, 0}; s1[i] = Acc3{0, 0, 0}; }
    for (int k = 0; k < mk.n; ++k) {
        const uint32_t src = mk.map[k][j];
        Split30 k0[BETA], k1[BETA];
#pragma unroll
        for (int d = 0; d < BETA; ++d) {
            const u64* kb = mk.evk[k] + (size_t)d * 2 * kpoly + (size_t)m * T.N + src;
            k0[d] = split30(__ldg(kb)); k1[d] = split30(__ldg(kb + kpoly));
        }
        const bool flush = (k % KPR) == KPR - 1 || k == mk.n - 1;
#pragma unroll
        for (int i = 0; i < IPB; ++i) {
            if (b0 + i >= batch) break;
#pragma unroll
            for (int d = 0; d < BETA; ++d) {
                const u64 u = d == own_d ? c_eval[(size_t)(b0 + i) * c_bs + (size_t)t * T.N + src]
                                         : up[(size_t)(b0 + i) * up_bs + ((size_t)d * ext + t) * T.N + src];
                const Split30 us = split30(u);
                mac3(s0[i], us, k0[d]); mac3(s1[i], us, k1[d]);
            }
            if (flush) {
                r0[i] = addmod(r0[i], reduce3(s0[i], rc), rc.q);
                r1[i] = addmod(r1[i], reduce3(s1[i], rc), rc.q);
                s0[i] = Acc3{0, 0, 0}; s1[i] = Acc3{0, 0, 0};
            }
        }
    }
#pragma unroll
    for (int i = 0; i < IPB; ++i) {
        if (b0 + i >= batch) break;
        u64* o = acc + (size_t)(b0 + i) * acc_bs + (size_t)t * T.N + j;
        o[0] = r0[i];
        o[(size_t)ext * T.N] = r1[i];
    }
}

// s0[b][i][j] = (self ? c0[b][i][j] : 0) + sum_k c0[b][i][map_k[j]]     grid (N / 256, l, batch)
__global__ void __launch_bounds__(kThreads) gather_sum_kernel(u64* __restrict__ s0, const u64* __restrict__ c0, MultiKeys mk, DevTables T, int l,
                                                              size_t s0_bs, size_t c_bs, int self) {
    const int j = blockIdx.x * kThreads + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    if (j >= T.N) return;
    const u64 q = T.q[i];
    const u64* src = c0 + (size_t)b * c_bs + (size_t)i * T.N;
    u64 v = self ? src[j] : 0;
    for (int k = 0; k < mk.n; ++k) v = addmod(v, src[mk.map[k][j]], q);
    s0[(size_t)b * s0_bs + (size_t)i * T.N + j] = v;
}

// ---- BSGS diagonal linear transform, double hoisting (CoeffsToSlots / SlotsToCoeffs, ct x pt matrix products) ----
// W[(j B + b)][p][t][x] = sum_i pt[j][i][t][x] * u_i[b][p][t][x] over the extended basis Q_l u P, where u_i is baby rotation i of
// ciphertext b BEFORE its ModDown: u_0 = P * ct (Q limbs; zero on the P limbs), u_i = sigma_i(IP_i(ModUp(c1)) + (P c0, 0)).
// The automorphism is the gather x -> map_i[x]; one ModDown per giant step j then serves all n1 products (Bossuat et al.).
// A thread owns coefficient x of limb t of BOTH polynomials of one ciphertext for every giant step, so a baby value is
// fetched once and a plaintext word is used twice; products accumulate carry-free (mac3), one reduction per 8 terms.
// grid (B, N / 256, ext): the B ciphertexts of a batch run side by side and share each plaintext word through L2.
template <int N1>
__global__ void __launch_bounds__(kThreads) bsgs_inner_kernel(u64* __restrict__ W, const u64* __restrict__ pts, const u64* __restrict__ pc,
                                                              BsgsArgs a, DevTables T, int l, int B, size_t acc_bs, size_t pc_bs) {
    const int b = blockIdx.x, t = blockIdx.z, ext = l + T.K;
    const int x = blockIdx.y * kThreads + threadIdx.x;
    if (x >= T.N) return;
    const int m = t < l ? t : T.L + (t - l);
    const RedC rc = load_redc(T, m);
    const size_t row = (size_t)t * T.N, ppoly = (size_t)ext * T.N, cpoly = (size_t)l * T.N;
    Split30 u0[N1], u1[N1];
#pragma unroll
    for (int i = 0; i < N1; ++i) {
        u64 v0 = 0, v1 = 0;
        if (i < a.n1) {
            if (i == 0) {
                if (t < l) { v0 = pc[(size_t)b * pc_bs + row + x]; v1 = pc[(size_t)b * pc_bs + cpoly + row + x]; }
            } else if (a.accb[i]) {
                const uint32_t src = a.map[i][x];
                const u64* ab = a.accb[i] + (size_t)b * acc_bs + row + src;
                v0 = ab[0]; v1 = ab[ppoly];
                if (t < l) v0 = addmod(v0, pc[(size_t)b * pc_bs + row + src], rc.q);
            }
        }
        u0[i] = split30(v0); u1[i] = split30(v1);
    }
    for (int j = 0; j < a.n2; ++j) {
        const uint32_t mask = a.mask[j];
        const u64* pj = pts + ((size_t)j * a.n1 * ext + t) * T.N + x;
        u64 r0 = 0, r1 = 0;
#pragma unroll
        for (int i0 = 0; i0 < N1; i0 += 8) {
            if (i0 >= a.n1) break;
            Acc3 s0{0, 0, 0}, s1{0, 0, 0};
#pragma unroll
            for (int i = i0; i < i0 + 8 && i < N1; ++i) {
                if (i < a.n1 && ((mask >> i) & 1u)) {
                    const Split30 ps = split30(__ldg(pj + (size_t)i * ppoly));
                    mac3(s0, u0[i], ps); mac3(s1, u1[i], ps);
                }
            }
            r0 = i0 ? addmod(r0, reduce3(s0, rc), rc.q) : reduce3(s0, rc);
            r1 = i0 ? addmod(r1, reduce3(s1, rc), rc.q) : reduce3(s1, rc);
        }
        u64* o = W + ((size_t)j * B + b) * acc_bs + row + x;
        o[0] = r0; o[ppoly] = r1;
    }
}

// The same product as a TMA-fed pipeline (the kernel every transform runs on).  A CTA owns 128 coefficients of one extended limb of
// one ciphertext, both polynomials (threads 0..127 / 128..255), so a plaintext word fetched once serves two multiplications.  The
// plaintext rows of giant step j -- N1 diagonals x 128 coefficients, 1 KiB each, contiguous in HBM -- are brought into shared memory
// by the TMA engine (cp.async.bulk, one 1-D copy per used diagonal, issued by one thread, completing on an mbarrier) kStages giant
// steps ahead of their use, so no thread ever issues or waits for a plaintext load: the baby values stay in registers (one
// polynomial per thread: 2 N1 registers), the multiplier runs while the next rows arrive.  The kernel above (loads issued by the
// threads, 74 / 156 registers) ran at 0.6 TB/s for 16 baby steps and 1.3 TB/s for 8: latency-bound.
template <int N1, int kStages>
__global__ void __launch_bounds__(kThreads, 3) bsgs_inner_tma_kernel(u64* __restrict__ W, const u64* __restrict__ pts, const u64* __restrict__ pc,
                                                                     BsgsArgs a, DevTables T, int l, int B, size_t acc_bs, size_t pc_bs) {
    constexpr int kCols = kThreads / 2;                       // coefficients per CTA
    extern __shared__ __align__(128) u64 bsgs_smem[];         // [kStages][N1][kCols] plaintext rows, then kStages mbarriers
    u64(*rows)[N1][kCols] = reinterpret_cast<u64(*)[N1][kCols]>(bsgs_smem);
    u64* full = bsgs_smem + (size_t)kStages * N1 * kCols;
    const int b = blockIdx.x, t = blockIdx.z, ext = l + T.K;
    const int pol = threadIdx.x / kCols, xl = threadIdx.x % kCols;
    const int x0 = blockIdx.y * kCols, x = x0 + xl;
    const size_t row = (size_t)t * T.N, ppoly = (size_t)ext * T.N, cpoly = (size_t)l * T.N;
    auto fill = [&](int j) {                                  // thread 0: the used rows of giant step j into stage j % kStages
        const int st = j % kStages;
        const uint32_t mask = a.mask[j], bar = (uint32_t)__cvta_generic_to_shared(&full[st]);
        const uint32_t bytes = (uint32_t)__popc(mask & ((1u << a.n1) - 1u)) * (uint32_t)(kCols * 8);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        const u64* pj = pts + ((size_t)j * a.n1 * ext + t) * T.N + x0;
        for (int i = 0; i < a.n1; ++i)
            if ((mask >> i) & 1u)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 (uint32_t)__cvta_generic_to_shared(&rows[st][i][0])),
                             "l"(pj + (size_t)i * ppoly), "r"((uint32_t)(kCols * 8)), "r"(bar)
                             : "memory");
    };
    if (threadIdx.x == 0) {
        for (int st = 0; st < kStages; ++st) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&full[st])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int j = 0; j < kStages && j < a.n2; ++j) fill(j);
    }
    const int m = t < l ? t : T.L + (t - l);
    const RedC rc = load_redc(T, m);
    Split30 u[N1];
#pragma unroll
    for (int i = 0; i < N1; ++i) {
        u64 v = 0;
        if (i < a.n1) {
            if (i == 0) {
                if (t < l) v = pc[(size_t)b * pc_bs + (size_t)pol * cpoly + row + x];
            } else if (a.accb[i]) {
                const uint32_t src = a.map[i][x];
                v = a.accb[i][(size_t)b * acc_bs + (size_t)pol * ppoly + row + src];
                if (pol == 0 && t < l) v = addmod(v, pc[(size_t)b * pc_bs + row + src], rc.q);
            }
        }
        u[i] = split30(v);
    }
    __syncthreads();                                          // the barriers are initialised before anybody waits on them
    for (int j = 0; j < a.n2; ++j) {
        const int st = j % kStages;
        const uint32_t mask = a.mask[j], bar = (uint32_t)__cvta_generic_to_shared(&full[st]), parity = (uint32_t)(j / kStages) & 1u;
        asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar),
                     "r"(parity)
                     : "memory");
        u64 r = 0;
#pragma unroll
        for (int i0 = 0; i0 < N1; i0 += 8) {
            Acc3 s{0, 0, 0};
#pragma unroll
            for (int i = i0; i < i0 + 8; ++i)
                if ((mask >> i) & 1u) mac3(s, u[i], split30(rows[st][i][xl]));      // warp-uniform branch; an unused row was not fetched
            r = i0 ? addmod(r, reduce3(s, rc), rc.q) : reduce3(s, rc);
        }
        W[((size_t)j * B + b) * acc_bs + (size_t)pol * ppoly + row + x] = r;
        __syncthreads();                                      // everybody has read stage st: it may be refilled
        if (threadIdx.x == 0 && j + kStages < a.n2) fill(j + kStages);
    }
}

// out[b][r][x] = sum_k src_k[b][r][map_k[x]]  (rows r of rpp limbs per polynomial, modulus by extended-basis index)   grid (N / 256, rows, B)
__global__ void __launch_bounds__(kThreads) gather_multi_kernel(u64* __restrict__ out, GatherArgs g, DevTables T, int l, int rpp, size_t out_bs,
                                                                size_t src_bs) {
    const int x = blockIdx.x * kThreads + threadIdx.x, r = blockIdx.y, b = blockIdx.z;
    if (x >= T.N) return;
    const int t = r % rpp;
    const u64 q = T.q[t < l ? t : T.L + (t - l)];
    const size_t o = (size_t)b * src_bs + (size_t)r * T.N;
    u64 v = 0;
    for (int k = 0; k < g.n; ++k) v = addmod(v, g.src[k][o + g.map[k][x]], q);
    out[(size_t)b * out_bs + (size_t)r * T.N + x] = v;
}

// grid: (N / 256, target groups, batch * polys); pcoef = P part (coefficient form, pre-scaled) of accumulator (b, p)
template <int KK>
__global__ void __launch_bounds__(kThreads) moddown_conv_kernel(u64* __restrict__ tq, const u64* __restrict__ pcoef, size_t pstride, DevTables T,
                                                                MdConst md, int l, int polys, size_t tq_bs, size_t p_bs, LimbRange rg) {
    extern __shared__ uint4 shs[];
    const int p = blockIdx.z % polys, b = blockIdx.z / polys;
    for (int i = threadIdx.x; i < KK * l; i += kThreads) {
        const Split30 h = split30(md.phm[(size_t)(i / l) * T.L + (i % l)]), g = split30(md.phm30[(size_t)(i / l) * T.L + (i % l)]);
        shs[i] = make_uint4(h.lo, h.hi, g.lo, g.hi);
    }
    __syncthreads();
    const int j = (blockIdx.x * kThreads + threadIdx.x) * 2;
    if (j >= T.N) return;
    Split30 y0[KK], y1[KK];
#pragma unroll
    for (int k = 0; k < KK; ++k) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(pcoef + (size_t)b * p_bs + (size_t)p * pstride + (size_t)k * T.N + j);
        y0[k] = split30(v.x); y1[k] = split30(v.y);
    }
    const int tg = gridDim.y, per = (rg.count + tg - 1) / tg;
    const int t0 = rg.first + blockIdx.y * per, t1 = min(t0 + per, rg.first + rg.count);
    u64* dst = tq + (size_t)b * tq_bs + (size_t)p * l * T.N + j;
    for (int t = t0; t < t1; ++t) {
        const RedC rc = load_redc(T, t);
        Acc2 a0{0, 0}, a1{0, 0};
#pragma unroll
        for (int k = 0; k < KK; ++k) {
            const uint4 h = shs[k * l + t];
            mac2(a0, y0[k], h);
            mac2(a1, y1[k], h);
        }
        *reinterpret_cast<ulonglong2*>(dst + (size_t)t * T.N) = make_ulonglong2(reduce2_lazy(a0.b0, a0.b1, rc), reduce2_lazy(a1.b0, a1.b1, rc));   // < 5q: NTT operand
    }
}

// out[b][p][i][j] = ((acc - tq) P^-1 + add_p)[b][i][map ? map[j] : j];  grid (N / 256, l, batch * polys)
__global__ void __launch_bounds__(kThreads) moddown_finish_kernel(u64* __restrict__ out, const u64* __restrict__ acc, size_t acc_ps,
                                                                  const u64* __restrict__ tq, const u64* __restrict__ add0,
                                                                  const u64* __restrict__ add1, const uint32_t* __restrict__ map, DevTables T,
                                                                  MdConst md, int l, int polys, size_t out_bs, size_t acc_bs, size_t tq_bs,
                                                                  size_t add0_bs, size_t add1_bs, const u64* __restrict__ plus, size_t plus_bs,
                                                                  int i_first) {
    const int i = i_first + blockIdx.y, p = blockIdx.z % polys, b = blockIdx.z / polys;
    const int j = blockIdx.x * kThreads + threadIdx.x;
    if (j >= T.N) return;
    const int src = map ? map[j] : j;
    const u64 q = T.q[i];
    const size_t o = (size_t)i * T.N + src;
    u64 v = submod(acc[(size_t)b * acc_bs + (size_t)p * acc_ps + o], tq[(size_t)b * tq_bs + (size_t)p * l * T.N + o], q);
    v = mul_shoup(v, md.pinv[i], md.pinv_sh[i], q);
    const u64* add = p == 0 ? add0 : add1;
    if (add) v = addmod(v, add[(size_t)b * (p == 0 ? add0_bs : add1_bs) + o], q);
    const size_t oo = ((size_t)p * l + i) * T.N + j;
    if (plus) v = addmod(v, plus[(size_t)b * plus_bs + oo], q);   // unpermuted addend: out = plus + rotate(...) (rotate-and-add ladders)
    out[(size_t)b * out_bs + oo] = v;
}

// ---------------- plaintext-weighted sums of ciphertexts (EvalLinearWSum) ----------------
// out[o] = sum_t k[o][t] * in[t], limb-wise.  A thread owns one coefficient of one limb of one polynomial for kOt outputs, so
// every input word is read n_out / kOt times.  k: [n_out][n_in][l] residues with Shoup companions ([.][.][.][2]).
constexpr int kOt = 8;
// rows = polynomials x limbs of one operand (2 l for a ciphertext, 2 l B for a batched one): limb = row mod l
// ptrs != nullptr: input t starts at ptrs[t] (operands scattered over the pool) instead of in + t * rows * N
__global__ void __launch_bounds__(kThreads) lincomb_kernel(u64* __restrict__ out, const u64* __restrict__ in, const u64* const* __restrict__ ptrs,
                                                           const ulonglong2* __restrict__ k, DevTables T, int l, int rows, int n_in, int n_out) {
    const int j = blockIdx.x * kThreads + threadIdx.x;
    if (j >= T.N) return;
    const int limb = blockIdx.y % l, o0 = blockIdx.z * kOt;
    const size_t poly_off = (size_t)blockIdx.y * T.N + j, ct = (size_t)rows * T.N;
    const u64 q = T.q[limb];
    u64 acc[kOt];
#pragma unroll
    for (int o = 0; o < kOt; ++o) acc[o] = 0;
    for (int t = 0; t < n_in; ++t) {
        const u64 x = ptrs ? ptrs[t][poly_off] : in[(size_t)t * ct + poly_off];
#pragma unroll
        for (int o = 0; o < kOt; ++o) {
            if (o0 + o < n_out) {
                const ulonglong2 c = __ldg(k + ((size_t)(o0 + o) * n_in + t) * l + limb);
                acc[o] = addmod(acc[o], mul_shoup(x, c.x, c.y, q), q);
            }
        }
    }
#pragma unroll
    for (int o = 0; o < kOt; ++o)
        if (o0 + o < n_out) out[(size_t)(o0 + o) * ct + poly_off] = acc[o];
}

// ---------------- modulus switch of one coefficient-form limb (ModRaise of the bootstrap; the rescale does its switch inside ntt.cu) ----------------
__global__ void __launch_bounds__(kThreads) mod_switch_kernel(u64* __restrict__ out, const u64* __restrict__ x, DevTables T, int src_mod, LimbSel sel) {
    const int p = blockIdx.z, i = blockIdx.y;
    const int j = blockIdx.x * kThreads + threadIdx.x;
    if (j >= T.N) return;
    const int m = sel.m[i];
    const u64 qs = T.q[src_mod], half = qs >> 1, q = T.q[m], ml = T.mu_lo[m], mh = T.mu_hi[m];
    const u64 v0 = x[(size_t)p * T.N + j];
    u64 v = barrett128(U128{v0, 0}, q, ml, mh);
    if (v0 > half) v = submod(v, barrett128(U128{qs, 0}, q, ml, mh), q);
    out[((size_t)p * sel.n + i) * T.N + j] = v;
}

__global__ void __launch_bounds__(kThreads) mul_i_kernel(u64* __restrict__ out, const u64* __restrict__ a, DevTables T, LimbSel sel, ScalarSet sc,
                                                         int polys) {
    const size_t per_poly = (size_t)sel.n * T.N;
    const size_t e = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (e >= per_poly * polys) return;
    const size_t r = e % per_poly;
    const int limb = (int)(r >> T.logN), j = (int)(r & (T.N - 1));
    const u64 q = T.q[sel.m[limb]];
    u64 v = mul_shoup(a[e], sc.c[limb], sc.c_sh[limb], q);
    if (j >= (T.N >> 1) && v) v = q - v;
    out[e] = v;
}

// ---------------- integer coefficients -> residues ----------------
__global__ void __launch_bounds__(kThreads) reduce_i64_kernel(u64* __restrict__ out, const int64_t* __restrict__ coef, DevTables T, LimbSel sel) {
    const int j = blockIdx.x * kThreads + threadIdx.x, limb = blockIdx.y;
    if (j >= T.N) return;
    const int m = sel.m[limb];
    const u64 q = T.q[m];
    const int64_t v = coef[j];
    const u64 a = v < 0 ? (u64)(-(v + 1)) + 1 : (u64)v;
    u64 r = barrett128(U128{a, 0}, q, T.mu_lo[m], T.mu_hi[m]);
    if (v < 0 && r) r = q - r;
    out[(size_t)limb * T.N + j] = r;
}
__global__ void __launch_bounds__(kThreads) reduce_i128_kernel(u64* __restrict__ out, const int64_t* __restrict__ coef, DevTables T, LimbSel sel) {
    const int j = blockIdx.x * kThreads + threadIdx.x, limb = blockIdx.y;
    if (j >= T.N) return;
    const int m = sel.m[limb];
    const u64 q = T.q[m];
    u64 lo = (u64)coef[2 * j];
    int64_t hi = coef[2 * j + 1];
    const bool neg = hi < 0;
    if (neg) { lo = ~lo + 1; hi = ~hi + (lo == 0); }
    u64 r = barrett128(U128{lo, (u64)hi}, q, T.mu_lo[m], T.mu_hi[m]);
    if (neg && r) r = q - r;
    out[(size_t)limb * T.N + j] = r;
}
__global__ void __launch_bounds__(kThreads) reduce_i8_kernel(u64* __restrict__ out, const int8_t* __restrict__ coef, DevTables T, LimbSel sel) {
    const int j = blockIdx.x * kThreads + threadIdx.x, limb = blockIdx.y;
    if (j >= T.N) return;
    const u64 q = T.q[sel.m[limb]];
    const int v = coef[j];
    out[(size_t)limb * T.N + j] = v >= 0 ? (u64)v : q - (u64)(-v);
}

inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace

void launch_ew(const DevTables& t, EwOp op, u64* out, const u64* a, const u64* b, const LimbSel& sel, int polys, int batch, size_t a_bs,
               size_t b_bs, size_t b_ps, cudaStream_t s) {
    dim3 grid(cdiv((size_t)polys * sel.n * t.N / 2, kThreads), batch);
    switch (op) {
        case EwOp::Add: ew_kernel<0><<<grid, kThreads, 0, s>>>(out, a, b, t, sel, polys, a_bs, b_bs, b_ps); break;
        case EwOp::Sub: ew_kernel<1><<<grid, kThreads, 0, s>>>(out, a, b, t, sel, polys, a_bs, b_bs, b_ps); break;
        case EwOp::Mul: ew_kernel<2><<<grid, kThreads, 0, s>>>(out, a, b, t, sel, polys, a_bs, b_bs, b_ps); break;
    }
    FLK_CUDA(cudaGetLastError());
}
void launch_ew_muladd(const DevTables& t, u64* out, const u64* a, const u64* b, const u64* c, const LimbSel& sel, int batch, size_t a_bs, size_t b_bs,
                      size_t c_bs, cudaStream_t s) {
    ew_muladd_kernel<<<dim3(cdiv((size_t)sel.n * t.N / 2, kThreads), batch), kThreads, 0, s>>>(out, a, b, c, t, sel, a_bs, b_bs, c_bs);
    FLK_CUDA(cudaGetLastError());
}
void launch_mul_scalar(const DevTables& t, u64* out, const u64* a, const ScalarSet& sc, const LimbSel& sel, int polys, cudaStream_t s) {
    mul_scalar_kernel<<<cdiv((size_t)polys * sel.n * t.N / 2, kThreads), kThreads, 0, s>>>(out, a, t, sel, sc, polys);
    FLK_CUDA(cudaGetLastError());
}
void launch_add_scalar(const DevTables& t, u64* out, const u64* a, const ScalarSet& sc, const LimbSel& sel, int batch, size_t bs, cudaStream_t s) {
    add_scalar_kernel<<<dim3(cdiv((size_t)sel.n * t.N / 2, kThreads), batch), kThreads, 0, s>>>(out, a, t, sel, sc, bs);
    FLK_CUDA(cudaGetLastError());
}
void launch_automorph(u64* out, const u64* in, const uint32_t* map, int N, int limbs, cudaStream_t s) {
    automorph_kernel<<<dim3(cdiv(N, kThreads), limbs), kThreads, 0, s>>>(out, in, map, N);
    FLK_CUDA(cudaGetLastError());
}
void launch_tensor(const DevTables& t, u64* d0, u64* d1, u64* d2, const u64* a, const u64* b, int l, int batch, size_t d_bs, size_t a_bs, size_t b_bs,
                   cudaStream_t s) {
    tensor_kernel<<<dim3(cdiv((size_t)l * t.N, kThreads), batch), kThreads, 0, s>>>(d0, d1, d2, a, b, t, l, d_bs, a_bs, b_bs);
    FLK_CUDA(cudaGetLastError());
}
void launch_modup_conv(const DevTables& t, const KsLevel& ks, u64* up, const u64* dcoef, int batch, size_t up_bs, size_t dco_bs, cudaStream_t s,
                       LimbRange rg) {
    if (ks.alpha > kAlphaMax) throw std::invalid_argument("digit size above 8 limbs is not supported");
    const int ext = ks.l + t.K;
    if (rg.count < 0) rg = LimbRange{0, ext};
    if (rg.count == 0) return;
    const int tg = rg.count >= 16 ? 4 : 1;
    const dim3 grid(cdiv(t.N / 2, kThreads), tg, ks.beta * batch);
    const size_t shm = (size_t)ks.alpha * ext * 16;
    switch (ks.alpha) {
#define FLK_CASE(X) case X: modup_conv_kernel<X><<<grid, kThreads, shm, s>>>(up, dcoef, t, ks, up_bs, dco_bs, rg); break;
        FLK_CASE(1) FLK_CASE(2) FLK_CASE(3) FLK_CASE(4) FLK_CASE(5) FLK_CASE(6) FLK_CASE(7) FLK_CASE(8)
#undef FLK_CASE
    }
    FLK_CUDA(cudaGetLastError());
}
void launch_inner_product(const DevTables& t, const KsLevel& ks, u64* acc, const u64* up, const u64* c_eval, const u64* evk, int batch,
                          size_t acc_bs, size_t up_bs, size_t c_bs, cudaStream_t s, LimbRange rg) {
    IpJobs one{};
    one.n = 1; one.acc[0] = acc; one.up[0] = up; one.c[0] = c_eval; one.evk[0] = evk;
    launch_inner_product_jobs(t, ks, one, batch, acc_bs, up_bs, c_bs, s, rg);
}
void launch_inner_product_jobs(const DevTables& t, const KsLevel& ks, const IpJobs& jobs, int batch, size_t acc_bs, size_t up_bs, size_t c_bs,
                               cudaStream_t s, LimbRange rg) {
    if (jobs.n < 1 || jobs.n > kIpJobsMax) throw std::invalid_argument("key inner product: 1..16 jobs per launch");
    if (rg.count < 0) rg = LimbRange{0, ks.l + t.K};
    if (rg.count == 0) return;
    const dim3 grid(cdiv(t.N / 2, kThreads), jobs.n * ((batch + kIpb - 1) / kIpb), rg.count);
    switch (ks.beta) {
#define FLK_CASE(X) case X: inner_product_kernel<X><<<grid, kThreads, 0, s>>>(jobs, t, ks, batch, acc_bs, up_bs, c_bs, rg.first); break;
        FLK_CASE(1) FLK_CASE(2) FLK_CASE(3) FLK_CASE(4) FLK_CASE(5) FLK_CASE(6) FLK_CASE(7) FLK_CASE(8)
#undef FLK_CASE
        default: throw std::invalid_argument("more than 8 key-switch digits is not supported");
    }
    FLK_CUDA(cudaGetLastError());
}
void launch_inner_product_multi(const DevTables& t, const KsLevel& ks, u64* acc, const u64* up, const u64* c_eval, const u64* const* evks,
                                const uint32_t* const* maps, int nk, int batch, size_t acc_bs, size_t up_bs, size_t c_bs, cudaStream_t s) {
    if (nk < 1 || nk > kHoistMax) throw std::invalid_argument("hoisted rotation sum: 1..15 rotations per call");
    MultiKeys mk{};
    mk.n = nk;
    for (int k = 0; k < nk; ++k) { mk.evk[k] = evks[k]; mk.map[k] = maps[k]; }
    const dim3 grid(cdiv(t.N, kThreads), (batch + 1) / 2, ks.l + t.K);
    static const bool group = [] { const char* e = std::getenv("FLK_IPM_GROUP"); return !e || e[0] != '0'; }();
    switch (ks.beta) {
#define FLK_CASE(X) case X: if (group) inner_product_multi_kernel<X, 8><<<grid, kThreads, 0, s>>>(acc, up, c_eval, mk, t, ks, batch, acc_bs, up_bs, c_bs); \
                            else inner_product_multi_kernel<X, 1><<<grid, kThreads, 0, s>>>(acc, up, c_eval, mk, t, ks, batch, acc_bs, up_bs, c_bs); break;
        FLK_CASE(1) FLK_CASE(2) FLK_CASE(3) FLK_CASE(4) FLK_CASE(5) FLK_CASE(6) FLK_CASE(7) FLK_CASE(8)
#undef FLK_CASE
        default: throw std::invalid_argument("more than 8 key-switch digits is not supported");
    }
    FLK_CUDA(cudaGetLastError());
}
void launch_gather_sum(const DevTables& t, u64* s0, const u64* c0, const uint32_t* const* maps, int nk, int l, int batch, size_t s0_bs, size_t c_bs,
                       bool self, cudaStream_t s) {
    MultiKeys mk{};
    mk.n = nk;
    for (int k = 0; k < nk; ++k) mk.map[k] = maps[k];
    gather_sum_kernel<<<dim3(cdiv(t.N, kThreads), l, batch), kThreads, 0, s>>>(s0, c0, mk, t, l, s0_bs, c_bs, self ? 1 : 0);
    FLK_CUDA(cudaGetLastError());
}
void launch_moddown_conv(const DevTables& t, const MdConst& md, u64* tq, const u64* pcoef, size_t pstride, int l, int polys, int batch,
                         size_t tq_bs, size_t p_bs, cudaStream_t s, LimbRange rg) {
    if (t.K > kAlphaMax) throw std::invalid_argument("more than 8 P limbs is not supported");
    if (rg.count < 0) rg = LimbRange{0, l};
    if (rg.count == 0) return;
    const int tg = rg.count >= 16 ? 4 : 1;
    const dim3 grid(cdiv(t.N / 2, kThreads), tg, polys * batch);
    const size_t shm = (size_t)t.K * l * 16;
    switch (t.K) {
#define FLK_CASE(X) case X: moddown_conv_kernel<X><<<grid, kThreads, shm, s>>>(tq, pcoef, pstride, t, md, l, polys, tq_bs, p_bs, rg); break;
        FLK_CASE(1) FLK_CASE(2) FLK_CASE(3) FLK_CASE(4) FLK_CASE(5) FLK_CASE(6) FLK_CASE(7) FLK_CASE(8)
#undef FLK_CASE
    }
    FLK_CUDA(cudaGetLastError());
}
void launch_moddown_finish(const DevTables& t, const MdConst& md, const FinishArgs& a, const uint32_t* map, int l, int polys, int batch,
                           cudaStream_t s, LimbRange rg) {
    if (rg.count < 0) rg = LimbRange{0, l};
    if (rg.count == 0) return;
    moddown_finish_kernel<<<dim3(cdiv(t.N, kThreads), rg.count, polys * batch), kThreads, 0, s>>>(a.out, a.acc, a.acc_ps, a.tq, a.add0, a.add1, map, t, md, l,
                                                                                                 polys, a.out_bs, a.acc_bs, a.tq_bs, a.add0_bs,
                                                                                                 a.add1_bs, a.plus, a.plus_bs, rg.first);
    FLK_CUDA(cudaGetLastError());
}
void launch_bsgs_inner(const DevTables& t, u64* W, const u64* pts, const u64* pc, const BsgsArgs& a, int l, int B, size_t acc_bs, size_t pc_bs,
                       cudaStream_t s) {
    if (a.n1 < 1 || a.n1 > kBsgsMax || a.n2 < 1 || a.n2 > kBsgsMax) throw std::invalid_argument("linear transform: 1..16 baby and giant steps");
    static const bool by_threads = [] { const char* e = std::getenv("FLK_BSGS_LOADS"); return e && e[0] == '1'; }();   // the former kernel, for comparison
    if (by_threads || t.N % (kThreads / 2) != 0) {
        const dim3 grid(B, cdiv(t.N, kThreads), l + t.K);
        if (a.n1 <= 8) bsgs_inner_kernel<8><<<grid, kThreads, 0, s>>>(W, pts, pc, a, t, l, B, acc_bs, pc_bs);
        else bsgs_inner_kernel<16><<<grid, kThreads, 0, s>>>(W, pts, pc, a, t, l, B, acc_bs, pc_bs);
    } else {
        const dim3 grid(B, t.N / (kThreads / 2), l + t.K);
        constexpr size_t shm8 = (size_t)4 * 8 * (kThreads / 2) * 8 + 4 * 8, shm16 = (size_t)3 * 16 * (kThreads / 2) * 8 + 3 * 8;
        static std::atomic<unsigned long long> configured{0};   // the attribute belongs to the device: one bit per device of this process
        int dev = 0;
        FLK_CUDA(cudaGetDevice(&dev));
        if (!((configured.load() >> (dev & 63)) & 1ull)) {
            FLK_CUDA(cudaFuncSetAttribute(bsgs_inner_tma_kernel<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm8));
            FLK_CUDA(cudaFuncSetAttribute(bsgs_inner_tma_kernel<16, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm16));
            configured.fetch_or(1ull << (dev & 63));
        }
        if (a.n1 <= 8) bsgs_inner_tma_kernel<8, 4><<<grid, kThreads, shm8, s>>>(W, pts, pc, a, t, l, B, acc_bs, pc_bs);
        else bsgs_inner_tma_kernel<16, 3><<<grid, kThreads, shm16, s>>>(W, pts, pc, a, t, l, B, acc_bs, pc_bs);
    }
    FLK_CUDA(cudaGetLastError());
}
void launch_gather_multi(const DevTables& t, u64* out, const GatherArgs& g, int l, int rows_per_poly, int rows, int batch, size_t out_bs, size_t src_bs,
                         cudaStream_t s) {
    gather_multi_kernel<<<dim3(cdiv(t.N, kThreads), rows, batch), kThreads, 0, s>>>(out, g, t, l, rows_per_poly, out_bs, src_bs);
    FLK_CUDA(cudaGetLastError());
}
void launch_lincomb(const DevTables& t, u64* out, const u64* in, const u64* k, int l, int rows, int n_in, int n_out, cudaStream_t s,
                    const u64* const* in_ptrs) {
    lincomb_kernel<<<dim3(cdiv(t.N, kThreads), rows, (n_out + kOt - 1) / kOt), kThreads, 0, s>>>(out, in, in_ptrs, reinterpret_cast<const ulonglong2*>(k),
                                                                                               t, l, rows, n_in, n_out);
    FLK_CUDA(cudaGetLastError());
}
void launch_mod_switch(const DevTables& t, u64* out, const u64* x, int src_mod, const LimbSel& sel, int polys, cudaStream_t s) {
    mod_switch_kernel<<<dim3(cdiv(t.N, kThreads), sel.n, polys), kThreads, 0, s>>>(out, x, t, src_mod, sel);
    FLK_CUDA(cudaGetLastError());
}
void launch_mul_i(const DevTables& t, u64* out, const u64* a, const ScalarSet& sc, const LimbSel& sel, int polys, cudaStream_t s) {
    mul_i_kernel<<<cdiv((size_t)polys * sel.n * t.N, kThreads), kThreads, 0, s>>>(out, a, t, sel, sc, polys);
    FLK_CUDA(cudaGetLastError());
}
void launch_reduce_i64(const DevTables& t, u64* out, const int64_t* coef, const LimbSel& sel, cudaStream_t s) {
    reduce_i64_kernel<<<dim3(cdiv(t.N, kThreads), sel.n), kThreads, 0, s>>>(out, coef, t, sel);
    FLK_CUDA(cudaGetLastError());
}
void launch_reduce_i128(const DevTables& t, u64* out, const int64_t* coef, const LimbSel& sel, cudaStream_t s) {
    reduce_i128_kernel<<<dim3(cdiv(t.N, kThreads), sel.n), kThreads, 0, s>>>(out, coef, t, sel);
    FLK_CUDA(cudaGetLastError());
}
void launch_reduce_i8(const DevTables& t, u64* out, const int8_t* coef, const LimbSel& sel, cudaStream_t s) {
    reduce_i8_kernel<<<dim3(cdiv(t.N, kThreads), sel.n), kThreads, 0, s>>>(out, coef, t, sel);
    FLK_CUDA(cudaGetLastError());
}

}  // namespace flk
