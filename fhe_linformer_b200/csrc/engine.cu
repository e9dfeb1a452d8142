// engine.cu -- device engine: table upload and the hybrid key-switch / rescale / rotate pipelines.
// Reference call sites replaced: context->EvalRotate (FHEController.cpp:435,833,843), the relinearisation
// inside context->EvalMult(ct,ct) (:431) and the implicit FLEXIBLEAUTO rescale (:18); algorithms per
// SURVEY.md Appendix A.5-A.7.
#include "engine.h"

#include <cstdlib>
#include <cstring>

namespace flk {

template <class V>
V* Engine::to_device(const std::vector<V>& h) {
    V* d = nullptr;
    FLK_CUDA(cudaMalloc(&d, std::max<size_t>(1, h.size()) * sizeof(V)));
    // pageable H2D copies may still be in flight on the legacy stream when cudaMemcpy returns; the engine stream is
    // non-blocking, so copy on it and wait.
    FLK_CUDA(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(V), cudaMemcpyHostToDevice, stream));
    FLK_CUDA(cudaStreamSynchronize(stream));
    owned_.push_back(d);
    return d;
}

Engine::Engine(const ParamSpec& spec, int device) : P(spec) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw std::runtime_error("fhe_linformer_b200: no CUDA device available (this engine has no CPU fallback)");
    if (device >= 0) FLK_CUDA(cudaSetDevice(device));
    FLK_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    int dev = 0;
    FLK_CUDA(cudaGetDevice(&dev));
    device_id = dev;
    cudaMemPool_t pool;
    FLK_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t keep = ~0ull;
    FLK_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));

    if (const char* ev = std::getenv("FLK_CACHE_GB")) cache_cap_bytes_ = (size_t)std::atol(ev) << 30;   // start-up default only; fl_ctx_set_cache_bytes is the interface
    if (const char* ev = std::getenv("FLK_L2_BUDGET_MB")) l2_budget_bytes = std::atol(ev) << 20;   // tuning knob (bench sweeps it)
    const int Tn = P.T, N = P.N;
    {   // twiddles interleaved with their Shoup companions so one 16-byte load fetches both
        std::vector<u64> tw(N), tws(N), itw(N), itws(N);
        std::vector<ulonglong2> f((size_t)Tn * N), b((size_t)Tn * N);
        for (int m = 0; m < Tn; ++m) {
            P.twiddles(m, tw.data(), tws.data(), itw.data(), itws.data());
            for (int i = 0; i < N; ++i) {
                f[(size_t)m * N + i] = make_ulonglong2(tw[i], tws[i]);
                b[(size_t)m * N + i] = make_ulonglong2(itw[i], itws[i]);
            }
        }
        T.tw2 = to_device(f); T.itw2 = to_device(b);
    }
    T.q = to_device(P.q); T.mu_lo = to_device(P.mu_lo); T.mu_hi = to_device(P.mu_hi);
    T.ninv = to_device(P.ninv); T.ninv_sh = to_device(P.ninv_sh);
    {
        std::vector<u64> rc((size_t)Tn * 8, 0);
        for (int m = 0; m < Tn; ++m) {
            const u64 qq = P.q[m];
            if ((qq >> 60) || qq < (1ull << 33)) throw std::invalid_argument("moduli must lie between 2^33 and 2^60");
            const u64 c30 = (1ull << 30) % qq, c60 = (1ull << 60) % qq;
            u64* r = &rc[(size_t)m * 8];
            r[0] = qq; r[1] = 0 - qq; r[2] = P.mu_hi[m]; r[3] = c30; r[4] = nt::shoup(c30, qq); r[5] = c60; r[6] = nt::shoup(c60, qq);
            r[7] = (u64)(((unsigned __int128)1 << 94) / qq);
        }
        T.redc = to_device(rc);
    }
    T.logN = P.logN; T.N = N; T.L = P.L; T.K = P.K;

    // ModDown constants (A.6)
    {
        std::vector<int> sm(P.K);
        for (int k = 0; k < P.K; ++k) sm[k] = P.L + k;
        std::vector<u64> hatinv(P.K), post(Tn, 0), post_sh(Tn, 0), phm((size_t)P.K * P.L), phm30(phm.size()), pinv(P.L), pinv_sh(P.L);
        P.conv_hatinv(sm.data(), P.K, hatinv.data());
        for (int k = 0; k < P.K; ++k) {
            u64 qq = P.q[P.L + k];
            post[P.L + k] = nt::mulmod(P.ninv[P.L + k], hatinv[k], qq);
            post_sh[P.L + k] = nt::shoup(post[P.L + k], qq);
            for (int i = 0; i < P.L; ++i) {
                phm[(size_t)k * P.L + i] = P.conv_hat_mod(sm.data(), P.K, k, P.q[i]);
                phm30[(size_t)k * P.L + i] = nt::mulmod(phm[(size_t)k * P.L + i], (1ull << 30) % P.q[i], P.q[i]);
            }
        }
        for (int i = 0; i < P.L; ++i) {
            pinv[i] = nt::invmod(P.P_mod(P.q[i]), P.q[i]);
            pinv_sh[i] = nt::shoup(pinv[i], P.q[i]);
            pmod_.c[i] = P.P_mod(P.q[i]);
            pmod_.c_sh[i] = nt::shoup(pmod_.c[i], P.q[i]);
        }
        md_.post = to_device(post); md_.post_sh = to_device(post_sh); md_.phm = to_device(phm); md_.phm30 = to_device(phm30);
        md_.pinv = to_device(pinv); md_.pinv_sh = to_device(pinv_sh);
    }
    // rescale constants (A.7)
    {
        std::vector<u64> inv((size_t)P.L * P.L, 0), inv_sh(inv.size(), 0);
        for (int r = 0; r < P.L; ++r)
            for (int i = 0; i < r; ++i) {
                inv[(size_t)r * P.L + i] = nt::invmod(P.q[r] % P.q[i], P.q[i]);
                inv_sh[(size_t)r * P.L + i] = nt::shoup(inv[(size_t)r * P.L + i], P.q[i]);
            }
        rs_.qlinv = to_device(inv); rs_.qlinv_sh = to_device(inv_sh);
    }
}

Engine::~Engine() {
    cudaSetDevice(device_id);
    trim_cache(0);
    cudaStreamSynchronize(stream);
    for (auto& kv : maps_) cudaFree(kv.second);
    for (void* p : owned_) cudaFree(p);
    for (auto& sl : host_slots_) {
        if (sl.in) { cudaFree(sl.in); cudaFree(sl.out); }
        if (sl.in_consumed) { cudaEventDestroy(sl.in_consumed); cudaEventDestroy(sl.out_drained); }
    }
    if (h2d_stream) cudaStreamDestroy(h2d_stream);
    if (d2h_stream) cudaStreamDestroy(d2h_stream);
    cudaStreamDestroy(stream);
    // hand the pool's unused reservations back to the driver: the next context on this device (another ring, another controller)
    // starts from free memory instead of a fragmented pool
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
}

// Device memory comes from CUDA's stream-ordered pool, fronted by an exact-size cache: a forward pass asks for the same few
// hundred block sizes over and over, and recycling a block of exactly that size on the one engine stream is free, whereas the
// pool alone fragments under the mixed sizes and sporadically stalls for 0.2-1.4 s mapping new physical memory
// (measured: scripts/forward_repeat.py).  Blocks are handed back to the pool only when the cache exceeds its cap.
u64* Engine::alloc(size_t words) {
    const size_t bytes = (std::max<size_t>(words, 1) * 8 + 0xFFFF) & ~(size_t)0xFFFF;
    u64* p = nullptr;
    // smallest cached block that fits, as long as it wastes at most a quarter (adjacent levels differ by one limb in ~25)
    auto it = free_blocks_.lower_bound(bytes);
    while (it != free_blocks_.end() && it->second.empty()) ++it;
    if (it != free_blocks_.end() && it->first <= bytes + bytes / 4) {
        p = it->second.back();
        it->second.pop_back();
        cached_bytes_ -= it->first;
        block_size_[p] = it->first;
        return p;
    } else {
        FLK_CUDA(cudaMallocAsync(&p, bytes, stream));
        ++pool_allocs;
    }
    block_size_[p] = bytes;
    return p;
}
void Engine::release(u64* p) {
    if (!p) return;
    auto it = block_size_.find(p);
    if (it == block_size_.end()) { FLK_CUDA(cudaFreeAsync(p, stream)); return; }
    const size_t bytes = it->second;
    block_size_.erase(it);
    free_blocks_[bytes].push_back(p);
    cached_bytes_ += bytes;
    if (cached_bytes_ > cache_cap_bytes_) trim_cache(cache_cap_bytes_ / 8 * 7);
}
void Engine::trim_cache(size_t keep_bytes) {
    cudaSetDevice(device_id);   // may run from fl_elem_free on a thread bound to another device
    ++cache_trims;
    for (auto it = free_blocks_.rbegin(); it != free_blocks_.rend() && cached_bytes_ > keep_bytes; ++it)   // largest classes first
        while (!it->second.empty() && cached_bytes_ > keep_bytes) {
            cudaFreeAsync(it->second.back(), stream);
            it->second.pop_back();
            cached_bytes_ -= it->first;
        }
}
void Engine::upload(u64* dst, const u64* src, size_t words) { FLK_CUDA(cudaMemcpyAsync(dst, src, words * 8, cudaMemcpyHostToDevice, stream)); }
void Engine::download(u64* dst, const u64* src, size_t words) {
    FLK_CUDA(cudaMemcpyAsync(dst, src, words * 8, cudaMemcpyDeviceToHost, stream));
    FLK_CUDA(cudaStreamSynchronize(stream));
}
void Engine::copy(u64* dst, const u64* src, size_t words) { FLK_CUDA(cudaMemcpyAsync(dst, src, words * 8, cudaMemcpyDeviceToDevice, stream)); }
void Engine::sync() {
    FLK_CUDA(cudaStreamSynchronize(stream));
    if (d2h_stream) { FLK_CUDA(cudaStreamSynchronize(h2d_stream)); FLK_CUDA(cudaStreamSynchronize(d2h_stream)); }
}

void Engine::ntt(u64* data, const LimbSel& sel, int batch, size_t bs) { launch_ntt(T, data, sel, batch, bs, stream); }
void Engine::intt(u64* data, const LimbSel& sel, int batch, size_t bs) { launch_intt(T, data, sel, batch, bs, nullptr, nullptr, stream); }

void Engine::ew(EwOp op, u64* out, const u64* a, const u64* b, int l, int polys, bool broadcast_b) {
    launch_ew(T, op, out, a, b, sel_range(0, l), polys, 1, 0, 0, broadcast_b ? 0 : (size_t)l * P.N, stream);
}
void Engine::ew_sel(EwOp op, u64* out, const u64* a, const u64* b, const LimbSel& sel) {
    launch_ew(T, op, out, a, b, sel, 1, 1, 0, 0, 0, stream);
}

const uint32_t* Engine::automorph_map(uint32_t g) {
    auto it = maps_.find(g);
    if (it != maps_.end()) return it->second;
    std::vector<uint32_t> h(P.N);
    P.automorph_map(g, h.data());
    uint32_t* d = nullptr;
    FLK_CUDA(cudaMalloc(&d, (size_t)P.N * 4));
    FLK_CUDA(cudaMemcpyAsync(d, h.data(), (size_t)P.N * 4, cudaMemcpyHostToDevice, stream));
    FLK_CUDA(cudaStreamSynchronize(stream));
    maps_[g] = d;
    return d;
}

// Forward transform of the ModDown conversion tq ([polys][l][N] per batch element) with the finish -- (acc - tq) P^-1 + addends,
// permuted by the automorphism of g -- folded into its last pass (launch_ntt_finish): tq is never written back in evaluation form.
void Engine::ntt_finish(u64* tq, size_t tq_bs, const FinishArgs& fa, uint32_t g, int l, int polys, int B) {
    const uint32_t* imap = nullptr;
    if (g) {
        uint32_t gi = 1;                                   // g^-1 mod 2N: the scatter map of sigma_g is the gather map of sigma_{g^-1}
        for (int i = 0; i < 6; ++i) gi = gi * (2 - g * gi);
        imap = automorph_map(gi & (uint32_t)(2 * P.N - 1));
    }
    NttFinish f{fa, imap, md_.pinv, md_.pinv_sh, l, polys};
    launch_ntt_finish(T, tq, B, tq_bs, f, stream);
}

void Engine::automorph(u64* out, const u64* in, uint32_t g, int limbs) { launch_automorph(out, in, automorph_map(g), P.N, limbs, stream); }

const KsLevel& Engine::ks_level(int l) {
    auto it = ks_.find(l);
    if (it != ks_.end()) return it->second;
    if (l < 1 || l > P.L) throw std::invalid_argument("key switch: limb count out of range");
    const int beta = P.beta(l), ext = l + P.K, a = P.alpha;
    if (beta * ext - l > kMaxLimbSel)
        throw std::invalid_argument("key switch at " + std::to_string(l) + " limbs needs " + std::to_string(beta * ext - l) + " extended limbs per ModUp; one launch addresses " +
                                    std::to_string(kMaxLimbSel) + " (fewer digits or a shorter chain)");
    std::vector<u64> post(P.T, 0), post_sh(P.T, 0), hm((size_t)beta * a * ext, 0), hm30(hm.size(), 0);
    for (int d = 0; d < beta; ++d) {
        const int lo = d * a, hi = std::min(lo + a, l), ns = hi - lo;
        std::vector<int> sm(ns);
        for (int i = 0; i < ns; ++i) sm[i] = lo + i;
        std::vector<u64> hatinv(ns);
        P.conv_hatinv(sm.data(), ns, hatinv.data());
        for (int i = 0; i < ns; ++i) {
            const int m = lo + i;
            post[m] = nt::mulmod(P.ninv[m], hatinv[i], P.q[m]);
            post_sh[m] = nt::shoup(post[m], P.q[m]);
            for (int t = 0; t < ext; ++t) {
                const u64 qt = P.q[P.mod_index_ext(l, t)];
                const u64 h = P.conv_hat_mod(sm.data(), ns, i, qt);
                hm[((size_t)d * a + i) * ext + t] = h;
                hm30[((size_t)d * a + i) * ext + t] = nt::mulmod(h, (1ull << 30) % qt, qt);
            }
        }
    }
    KsLevel k;
    k.post = to_device(post); k.post_sh = to_device(post_sh); k.hm = to_device(hm); k.hm30 = to_device(hm30);
    k.l = l; k.beta = beta; k.alpha = a;
    return ks_.emplace(l, k).first->second;
}

// DropLastElementAndScale of `polys` polynomials stored back to back ([polys][l][N] -> [polys][l-1][N]); a batch of
// ciphertexts is simply polys = 2 * batch.
void Engine::rescale(u64* out, const u64* in, int l, int polys) {
    if (l < 2) throw std::invalid_argument("rescale: no limb left to drop");
    const int N = P.N;
    // Four launches, each polynomial a batch element: the inverse transform reads the last limb where it lies, and the
    // (x_i - switched) q_last^-1 epilogue rides on the last pass of the forward transform (the ModDown finish with P = q_last).
    u64* xlast = alloc((size_t)polys * N);
    LimbSel s1; s1.push(l - 1, 0);
    launch_intt(T, xlast, s1, polys, (size_t)N, nullptr, nullptr, stream, in + (size_t)(l - 1) * N, (size_t)l * N);
    u64* tq = alloc((size_t)polys * (l - 1) * N);
    FinishArgs fa{};
    fa.out = out; fa.out_bs = (size_t)(l - 1) * N;
    fa.acc = in; fa.acc_ps = 0; fa.acc_bs = (size_t)l * N;
    fa.tq = tq; fa.tq_bs = (size_t)(l - 1) * N;
    NttFinish f{fa, nullptr, rs_.qlinv + (size_t)(l - 1) * P.L, rs_.qlinv_sh + (size_t)(l - 1) * P.L, l - 1, 1};
    f.switch_src = xlast; f.switch_mod = l - 1;           // the centred switch of the dropped limb happens in the first pass's loads
    launch_ntt_finish(T, tq, polys, (size_t)(l - 1) * N, f, stream);
    release(xlast); release(tq);
    if (ledger_on) ledger.add("rescale", l, (32.0 * l - 16.0) * N, polys / 2 > 0 ? polys / 2 : 1);
}

// Forward NTT of a batch in sub-batches small enough that the first pass's output is still in L2 (126 MB) when the second
// pass reads it: the transform is latency-bound, and an L2 hit costs less than half of an HBM access.
void Engine::ntt_l2(u64* data, const LimbSel& sel, int batch, size_t bs) {
    const size_t per = (size_t)sel.n * P.N * 8;
    int sub = (int)std::max<size_t>(1, (size_t)l2_budget_bytes / std::max<size_t>(per, 1));
    for (int b0 = 0; b0 < batch; b0 += sub) launch_ntt(T, data + (size_t)b0 * bs, sel, std::min(sub, batch - b0), bs, stream);
}

// ModUp of INTT'd, pre-scaled digits: base conversion to the limbs outside each digit, then the forward transform of those limbs.
// (Folding the conversion into the loads of the transform's first pass -- VERDICT r01 item 2(ii) -- was built and measured in round 2:
// bit-exact, but 2 464 instead of 3 981 rotations/s.  A column-pass thread owns 16 coefficients of ONE target limb, so every one of
// the l + K - alpha targets of a digit re-reads the digit's alpha source limbs through L2, 28 times the loads of the conversion
// kernel, which keeps the sources of two coefficients in registers across all targets.  DESIGN.md section 4.)
void Engine::modup_ntt(const KsLevel& ks, u64* up, const u64* dco, int B, size_t up_bs, size_t dco_bs) {
    const int l = ks.l, ext = l + P.K;
    LimbSel su;
    for (int d = 0; d < ks.beta; ++d) {
        const int lo = d * ks.alpha, hi = std::min(lo + ks.alpha, l);
        for (int t = 0; t < ext; ++t) {
            if (t >= lo && t < hi) continue;
            su.push(P.mod_index_ext(l, t), d * ext + t);
        }
    }
    launch_modup_conv(T, ks, up, dco, B, up_bs, dco_bs, stream);
    ntt_l2(up, su, B, up_bs);
}

// Hybrid key switch of a batch of polynomials with one evaluation key (HYBRID KeySwitch of EvalRotate / EvalMult, A.6):
//   INTT digits (pre-scaled) -> ModUp base conversion -> NTT -> inner product with the key over Q_l u P -> INTT of the P part
//   -> ModDown conversion -> NTT -> (acc - conv) P^-1 + addends, optionally permuted by the automorphism of g.
// Every stage is one launch over the whole batch; work buffers are batch-contiguous.
void Engine::keyswitch(const KsBatch& io, const u64* evk, uint32_t g) {
    const int l = io.l, B = io.B;
    if (B <= 0) return;
    const KsLevel& ks = ks_level(l);
    const int N = P.N, K = P.K, ext = l + K, beta = ks.beta;
    const size_t dco_bs = (size_t)l * N, up_bs = (size_t)beta * ext * N, acc_bs = (size_t)2 * ext * N, tq_bs = (size_t)2 * l * N;
    // 1. digits to coefficient form, pre-scaled by (Q_d/q_i)^-1
    u64* dco = alloc(dco_bs * B);
    launch_intt(T, dco, sel_range(0, l), B, dco_bs, ks.post, ks.post_sh, stream, io.c, io.c_bs);   // reads c1 in place: no gather copy
    // 2. ModUp: basis-extend every digit to the limbs outside it, back to evaluation form
    u64* up = alloc(up_bs * B);
    modup_ntt(ks, up, dco, B, up_bs, dco_bs);
    // 3. inner product with the evaluation key over Q_l u P
    u64* acc = alloc(acc_bs * B);
    launch_inner_product(T, ks, acc, up, io.c, evk, B, acc_bs, up_bs, io.c_bs, stream);
    // 4. ModDown both accumulators
    LimbSel sp;
    for (int p = 0; p < 2; ++p)
        for (int k = 0; k < K; ++k) sp.push(P.L + k, p * ext + l + k);
    launch_intt(T, acc, sp, B, acc_bs, md_.post, md_.post_sh, stream);
    u64* tq = alloc(tq_bs * B);
    launch_moddown_conv(T, md_, tq, acc + (size_t)l * N, (size_t)ext * N, l, 2, B, tq_bs, acc_bs, stream);
    FinishArgs fa{io.out, io.out_bs, acc, (size_t)ext * N, acc_bs, tq, tq_bs, io.add0, io.add0_bs, io.add1, io.add1_bs, io.plus, io.plus_bs};
    ntt_finish(tq, tq_bs, fa, g, l, 2, B);
    release(dco); release(up); release(acc); release(tq);
}

void Engine::keyswitch(u64* out, const u64* c, const u64* evk, int l, const u64* add0, const u64* add1, uint32_t g) {
    KsBatch io{1, l, c, 0, out, 0, add0, 0, add1, 0, nullptr, 0};
    keyswitch(io, evk, g);
}

void Engine::rotate(u64* out, const u64* ct, int l, uint32_t g, const u64* evk) { rotate_batch(out, ct, l, g, evk, 1, false); }
void Engine::rotate_add(u64* out, const u64* ct, int l, uint32_t g, const u64* evk) { rotate_batch(out, ct, l, g, evk, 1, true); }

// out[b] = rotate(ct[b]) (+ ct[b] when `accumulate`) for B ciphertexts stored back to back ([B][2][l][N])
void Engine::rotate_batch(u64* out, const u64* ct, int l, uint32_t g, const u64* evk, int B, bool accumulate) {
    const size_t cs = (size_t)2 * l * P.N;
    KsBatch io{B, l, ct + (size_t)l * P.N, cs, out, cs, ct, cs, nullptr, 0, accumulate ? ct : nullptr, cs};
    keyswitch(io, evk, g);
    if (ledger_on) {
        ledger.add("rotate", l, (4.0 * l + 2.0 * P.beta(l) * (l + P.K)) * 8.0 * P.N, B);
        if (accumulate) ledger.add("add", l, 48.0 * l * P.N, B);
    }
}

// ---- limb-sharded key switch: the stages of keyswitch() restricted to limb ranges of the extended basis ----
// G ranks hold the same input polynomial and split the l + K limbs of Q_l u P between them (sharded.py).  Every stage is
// limb-wise independent except the two base conversions, which need all source limbs: the digits (cheap, recomputed by every
// rank) and the P limbs of the accumulators (exchanged over NVLink between ks_pcoef_part and ks_moddown_part).
void Engine::ks_digits(u64* dco, const u64* c, int l) { ks_digits_part(dco, c, l, 0, l); }
// digits first .. first + count - 1 only (a rank's share when the digit limbs are exchanged instead of recomputed)
void Engine::ks_digits_part(u64* dco, const u64* c, int l, int first, int count) {
    if (count <= 0) return;
    const KsLevel& ks = ks_level(l);
    LimbSel s;
    for (int i = first; i < first + count; ++i) s.push(i, i);
    launch_intt(T, dco, s, 1, 0, ks.post, ks.post_sh, stream, c, 0);
}
void Engine::ks_modup_part(u64* up, const u64* dco, int l, int first, int count) {
    const KsLevel& ks = ks_level(l);
    const int ext = l + P.K;
    launch_modup_conv(T, ks, up, dco, 1, 0, 0, stream, LimbRange{first, count});
    LimbSel su;
    for (int d = 0; d < ks.beta; ++d) {
        const int lo = d * ks.alpha, hi = std::min(lo + ks.alpha, l);
        for (int t = first; t < first + count; ++t) {
            if (t >= lo && t < hi) continue;
            su.push(P.mod_index_ext(l, t), d * ext + t);
        }
    }
    launch_ntt(T, up, su, 1, 0, stream);
}
void Engine::ks_inner_part(u64* acc, const u64* up, const u64* c, const u64* evk, int l, int first, int count) {
    launch_inner_product(T, ks_level(l), acc, up, c, evk, 1, 0, 0, 0, stream, LimbRange{first, count});
}
void Engine::ks_pcoef_part(u64* acc, int l, int first, int count) {   // first, count: a range of the K special limbs (0-based inside P)
    const int ext = l + P.K;
    LimbSel sp;
    for (int p = 0; p < 2; ++p)
        for (int k = first; k < first + count; ++k) sp.push(P.L + k, p * ext + l + k);
    launch_intt(T, acc, sp, 1, 0, md_.post, md_.post_sh, stream);
}
void Engine::ks_moddown_part(u64* out, u64* tq, const u64* acc, int l, int first, int count, const u64* add0, const u64* add1, uint32_t g) {
    const int N = P.N, ext = l + P.K;
    launch_moddown_conv(T, md_, tq, acc + (size_t)l * N, (size_t)ext * N, l, 2, 1, 0, 0, stream, LimbRange{first, count});
    LimbSel sq;
    for (int p = 0; p < 2; ++p)
        for (int i = first; i < first + count; ++i) sq.push(i, p * l + i);
    launch_ntt(T, tq, sq, 1, 0, stream);
    FinishArgs fa{out, 0, acc, (size_t)ext * N, 0, tq, 0, add0, 0, add1, 0, nullptr, 0};
    launch_moddown_finish(T, md_, fa, g ? automorph_map(g) : nullptr, l, 2, 1, stream, LimbRange{first, count});
}

// Hoisted rotate-and-sum: out = (self ? ct : 0) + sum_k sigma_k(ct) for nk rotations of the SAME ciphertexts.  The digit
// decomposition / ModUp (45 % of a key switch) is done once, the nk evaluation-key inner products accumulate -- each gathered
// through its automorphism map -- in the extended basis Q_l u P, and a single ModDown brings the sum back (double hoisting,
// Bossuat et al.).  Used by the rotate-and-add ladders (two doubling steps = three rotations of one operand).
void Engine::rotate_sum_batch(u64* out, const u64* ct, int l, const uint32_t* gs, const u64* const* evks, int nk, int B, bool self) {
    if (B <= 0) return;
    if (nk < 1 || nk > kHoistMax) throw std::invalid_argument("hoisted rotation sum: 1..15 rotations per call");
    const KsLevel& ks = ks_level(l);
    const int N = P.N, K = P.K, ext = l + K, beta = ks.beta;
    const size_t cs = (size_t)2 * l * N, dco_bs = (size_t)l * N, up_bs = (size_t)beta * ext * N, acc_bs = (size_t)2 * ext * N, tq_bs = cs;
    const u64* c1 = ct + (size_t)l * N;
    const uint32_t* maps[kHoistMax];
    for (int k = 0; k < nk; ++k) maps[k] = automorph_map(gs[k]);
    u64* dco = alloc(dco_bs * B);
    launch_intt(T, dco, sel_range(0, l), B, dco_bs, ks.post, ks.post_sh, stream, c1, cs);
    u64* up = alloc(up_bs * B);
    modup_ntt(ks, up, dco, B, up_bs, dco_bs);
    u64* acc = alloc(acc_bs * B);
    launch_inner_product_multi(T, ks, acc, up, c1, evks, maps, nk, B, acc_bs, up_bs, cs, stream);
    u64* s0 = alloc(dco_bs * B);
    launch_gather_sum(T, s0, ct, maps, nk, l, B, dco_bs, cs, self, stream);
    LimbSel sp;
    for (int p = 0; p < 2; ++p)
        for (int k = 0; k < K; ++k) sp.push(P.L + k, p * ext + l + k);
    launch_intt(T, acc, sp, B, acc_bs, md_.post, md_.post_sh, stream);
    u64* tq = alloc(tq_bs * B);
    launch_moddown_conv(T, md_, tq, acc + (size_t)l * N, (size_t)ext * N, l, 2, B, tq_bs, acc_bs, stream);
    FinishArgs fa{out, cs, acc, (size_t)ext * N, acc_bs, tq, tq_bs, s0, dco_bs, self ? c1 : nullptr, cs, nullptr, 0};
    ntt_finish(tq, tq_bs, fa, 0, l, 2, B);
    release(dco); release(up); release(acc); release(s0); release(tq);
    if (ledger_on) {
        // the ledger follows the reference's operation census, not the work done here: a hoisted group of
        // 2^s - 1 rotations stands for s rotate-and-add steps of the sequential ladder
        int s_steps = 0;
        while ((1 << s_steps) - 1 < nk) ++s_steps;
        ledger.add("rotate", l, (4.0 * l + 2.0 * P.beta(l) * (l + P.K)) * 8.0 * P.N, B * s_steps);
        ledger.add("add", l, 48.0 * l * P.N, B * s_steps);
    }
}

void Engine::modup_batch(u64* up, const u64* c, size_t c_bs, int Bn, int l) {
    const KsLevel& ks = ks_level(l);
    const int N = P.N, ext = l + P.K, beta = ks.beta;
    const size_t dco_bs = (size_t)l * N, up_bs = (size_t)beta * ext * N;
    u64* dco = alloc(dco_bs * Bn);
    launch_intt(T, dco, sel_range(0, l), Bn, dco_bs, ks.post, ks.post_sh, stream, c, c_bs);
    modup_ntt(ks, up, dco, Bn, up_bs, dco_bs);
    release(dco);
}

void Engine::moddown_acc(u64* out, size_t out_bs, u64* acc, int Bn, int l, const u64* add0, size_t add0_bs, const u64* add1, size_t add1_bs,
                         const u64* plus, size_t plus_bs, uint32_t g) {
    const int N = P.N, K = P.K, ext = l + K;
    const size_t acc_bs = (size_t)2 * ext * N, tq_bs = (size_t)2 * l * N;
    LimbSel sp;
    for (int p = 0; p < 2; ++p)
        for (int k = 0; k < K; ++k) sp.push(P.L + k, p * ext + l + k);
    launch_intt(T, acc, sp, Bn, acc_bs, md_.post, md_.post_sh, stream);
    u64* tq = alloc(tq_bs * Bn);
    launch_moddown_conv(T, md_, tq, acc + (size_t)l * N, (size_t)ext * N, l, 2, Bn, tq_bs, acc_bs, stream);
    FinishArgs fa{out, out_bs, acc, (size_t)ext * N, acc_bs, tq, tq_bs, add0, add0_bs, add1, add1_bs, plus, plus_bs};
    ntt_finish(tq, tq_bs, fa, g, l, 2, Bn);
    release(tq);
}

// BSGS diagonal linear transform with double hoisting (the ct x pt matrix product of CoeffsToSlots / SlotsToCoeffs and of the
// packed linear layers).  Against n1 - 1 + r separate rotations (r = rotating giant steps) it runs 1 + r ModUps instead of
// n1 - 1 + r and n2 + 1 ModDowns instead of n1 - 1 + r, and the n1 n2 plaintext products + sums become one kernel that
// streams each plaintext diagonal once per batch.
void Engine::linear_transform(u64* out, const u64* ct, int B, const LtPlan& p) {
    if (B <= 0) return;
    const int l = p.l, n1 = p.n1, n2 = p.n2;
    const KsLevel& ks = ks_level(l);
    const int N = P.N, ext = l + P.K, beta = ks.beta;
    const size_t cs = (size_t)2 * l * N, dco_bs = (size_t)l * N, up_bs = (size_t)beta * ext * N, acc_bs = (size_t)2 * ext * N;
    const u64* c1 = ct + (size_t)l * N;
    // u_0 = P * ct on the Q limbs
    u64* pc = alloc(cs * B);
    launch_mul_scalar(T, pc, ct, pmod_, sel_range(0, l), 2 * B, stream);
    // baby steps: one ModUp, one key inner product per rotation, no ModDown
    BsgsArgs a{};
    a.n1 = n1; a.n2 = n2;
    for (int j = 0; j < n2; ++j) a.mask[j] = p.mask[j];
    uint32_t used = 0;
    for (int j = 0; j < n2; ++j) used |= p.mask[j];
    u64* accb = nullptr;
    int nb = 0;
    if (used >> 1) {
        u64* up = alloc(up_bs * B);
        modup_batch(up, c1, cs, B, l);
        accb = alloc(acc_bs * B * (size_t)(n1 - 1));
        IpJobs babies{};                                       // every baby rotation's key product in one launch (the operand is shared)
        for (int i = 1; i < n1; ++i) {
            if (!((used >> i) & 1u)) continue;
            u64* dst = accb + (size_t)(i - 1) * B * acc_bs;
            babies.acc[babies.n] = dst; babies.up[babies.n] = up; babies.c[babies.n] = c1; babies.evk[babies.n] = p.baby_evk[i];
            ++babies.n;
            a.accb[i] = dst; a.map[i] = automorph_map(p.baby_g[i]);
            ++nb;
        }
        if (babies.n) launch_inner_product_jobs(T, ks, babies, B, acc_bs, up_bs, cs, stream);
        release(up);
    }
    u64* W = alloc(acc_bs * B * (size_t)n2);
    launch_bsgs_inner(T, W, p.pts, pc, a, l, B, acc_bs, cs, stream);
    if (accb) release(accb);
    release(pc);
    // one ModDown per giant step (all of them as one batch)
    int r = 0;
    for (int j = 0; j < n2; ++j) if (p.giant_g[j] != 1) ++r;
    const bool ident = r < n2;     // the non-rotating giant step is the last one
    if (r == 0) {
        moddown_acc(out, cs, W, B, l, nullptr, 0, nullptr, 0, nullptr, 0, 0);
        release(W);
    } else {
        u64* w = alloc(cs * B * (size_t)n2);
        moddown_acc(w, cs, W, B * n2, l, nullptr, 0, nullptr, 0, nullptr, 0, 0);
        release(W);
        // giant steps: different operands, different rotations -- batched ModUp, one inner product each, summed through their
        // automorphism maps in the extended basis, a single ModDown
        u64* up2 = alloc(up_bs * B * (size_t)r);
        modup_batch(up2, w + (size_t)l * N, cs, B * r, l);
        u64* acc2 = alloc(acc_bs * B * (size_t)r);
        GatherArgs gz{}, g0{};
        gz.n = g0.n = r;
        IpJobs giants{};                                       // one launch for all giant steps (each has its own operand and key)
        for (int j = 0; j < r; ++j) {
            const u64* wj = w + (size_t)j * B * cs;
            giants.acc[giants.n] = acc2 + (size_t)j * B * acc_bs; giants.up[giants.n] = up2 + (size_t)j * B * up_bs;
            giants.c[giants.n] = wj + (size_t)l * N; giants.evk[giants.n] = p.giant_evk[j];
            ++giants.n;
            gz.src[j] = acc2 + (size_t)j * B * acc_bs; g0.src[j] = wj;
            gz.map[j] = g0.map[j] = automorph_map(p.giant_g[j]);
        }
        launch_inner_product_jobs(T, ks, giants, B, acc_bs, up_bs, cs, stream);
        release(up2);
        u64* Z = alloc(acc_bs * B);
        launch_gather_multi(T, Z, gz, l, ext, 2 * ext, B, acc_bs, acc_bs, stream);
        u64* s0 = alloc(dco_bs * B);
        launch_gather_multi(T, s0, g0, l, l, l, B, dco_bs, cs, stream);
        release(acc2);
        moddown_acc(out, cs, Z, B, l, s0, dco_bs, nullptr, 0, ident ? w + (size_t)r * B * cs : nullptr, cs, 0);
        release(Z); release(s0); release(w);
    }
    if (ledger_on) {
        // the reference's census of the same transform: one EvalRotate per baby / giant rotation, one ct x pt product and one
        // addition per diagonal
        ledger.add("rotate", l, (4.0 * l + 2.0 * P.beta(l) * (l + P.K)) * 8.0 * P.N, B * (nb + r));
        ledger.add("mul_plain", l, 40.0 * P.N * l, B * p.ndiag);
        ledger.add("add", l, 48.0 * l * P.N, B * std::max(0, p.ndiag - 1));
    }
}

// Host-resident caller (the reference keeps ciphertexts in host memory): chunk c+1 is uploaded while chunk c is key
// switched and chunk c-1 is downloaded, so a PCIe-bound call costs max(H2D, D2H, compute) instead of their sum.
// The device-side staging areas are two persistent slots used alternately by successive calls, each guarded by its own
// events, so with wait = false call n+1 starts uploading while call n is still computing / downloading (the caller
// synchronises once with sync()); wait = true returns with the results in out_host.
void Engine::rotate_batch_host(u64* out_host, const u64* ct_host, int l, uint32_t g, const u64* evk, int B, int chunk, bool wait) {
    if (B <= 0) return;
    if (!h2d_stream) {
        FLK_CUDA(cudaStreamCreateWithFlags(&h2d_stream, cudaStreamNonBlocking));
        FLK_CUDA(cudaStreamCreateWithFlags(&d2h_stream, cudaStreamNonBlocking));
        for (auto& sl : host_slots_) {
            FLK_CUDA(cudaEventCreateWithFlags(&sl.in_consumed, cudaEventDisableTiming));
            FLK_CUDA(cudaEventCreateWithFlags(&sl.out_drained, cudaEventDisableTiming));
        }
    }
    chunk = std::max(1, std::min(chunk, B));
    const size_t cs = (size_t)2 * l * P.N;
    if (host_slots_[0].words < cs * B) {           // grow both slots: only after everything that uses them has finished
        sync();
        for (auto& s2 : host_slots_) {
            if (s2.in) { FLK_CUDA(cudaFree(s2.in)); FLK_CUDA(cudaFree(s2.out)); }
            FLK_CUDA(cudaMalloc(&s2.in, cs * B * 8)); FLK_CUDA(cudaMalloc(&s2.out, cs * B * 8));
            s2.words = cs * B;
        }
    }
    HostSlot& sl = host_slots_[host_next_++ & 1];
    automorph_map(g);   // table upload (first use of g) happens before the pipeline starts
    ks_level(l);
    FLK_CUDA(cudaStreamWaitEvent(h2d_stream, sl.in_consumed, 0));     // the slot's previous key switches have read their inputs
    FLK_CUDA(cudaStreamWaitEvent(stream, sl.out_drained, 0));         // ... and its previous results have left for the host
    const int nch = (B + chunk - 1) / chunk;
    std::vector<cudaEvent_t> up(nch), done(nch);
    for (int c = 0; c < nch; ++c) {
        FLK_CUDA(cudaEventCreateWithFlags(&up[c], cudaEventDisableTiming));
        FLK_CUDA(cudaEventCreateWithFlags(&done[c], cudaEventDisableTiming));
    }
    for (int c = 0; c < nch; ++c) {
        const int b0 = c * chunk, nb = std::min(chunk, B - b0);
        FLK_CUDA(cudaMemcpyAsync(sl.in + b0 * cs, ct_host + b0 * cs, cs * nb * 8, cudaMemcpyHostToDevice, h2d_stream));
        FLK_CUDA(cudaEventRecord(up[c], h2d_stream));
        FLK_CUDA(cudaStreamWaitEvent(stream, up[c], 0));
        rotate_batch(sl.out + b0 * cs, sl.in + b0 * cs, l, g, evk, nb, false);
        FLK_CUDA(cudaEventRecord(done[c], stream));
        FLK_CUDA(cudaStreamWaitEvent(d2h_stream, done[c], 0));
        FLK_CUDA(cudaMemcpyAsync(out_host + b0 * cs, sl.out + b0 * cs, cs * nb * 8, cudaMemcpyDeviceToHost, d2h_stream));
    }
    FLK_CUDA(cudaEventRecord(sl.in_consumed, stream));
    FLK_CUDA(cudaEventRecord(sl.out_drained, d2h_stream));
    for (int c = 0; c < nch; ++c) { cudaEventDestroy(up[c]); cudaEventDestroy(done[c]); }   // released when they complete
    if (wait) sync();
}

void Engine::mul_relin(u64* out, const u64* a, const u64* b, int l, const u64* evk) { mul_relin_batch(out, a, b, l, evk, 1, 0); }

// EvalMult(ct, ct) for B ciphertext pairs stored back to back (b_bs = 0 broadcasts one right operand): tensor product, then
// one batched key switch of the d2 components with the relinearisation key
void Engine::mul_relin_batch(u64* out, const u64* a, const u64* b, int l, const u64* evk, int B, size_t b_bs) {
    const size_t pl = (size_t)l * P.N;
    u64* d = alloc(3 * pl * B);
    launch_tensor(T, d, d + pl, d + 2 * pl, a, b, l, B, 3 * pl, 2 * pl, b_bs, stream);
    KsBatch io{B, l, d + 2 * pl, 3 * pl, out, 2 * pl, d, 3 * pl, d + pl, 3 * pl, nullptr, 0};
    keyswitch(io, evk, 0);
    release(d);
    if (ledger_on) ledger.add("mul_relin", l, (6.0 * l + 2.0 * P.beta(l) * (l + P.K)) * 8.0 * P.N, B);
}

void Engine::modup(u64* out_ext, const u64* c_eval, int l, int digit) {
    const KsLevel& ks = ks_level(l);
    const int N = P.N, ext = l + P.K;
    u64* dco = alloc((size_t)l * N);
    copy(dco, c_eval, (size_t)l * N);
    launch_intt(T, dco, sel_range(0, l), 1, 0, ks.post, ks.post_sh, stream);
    u64* up = alloc((size_t)ks.beta * ext * N);
    launch_modup_conv(T, ks, up, dco, 1, 0, 0, stream);
    const int lo = digit * ks.alpha, hi = std::min(lo + ks.alpha, l);
    LimbSel su;
    for (int t = 0; t < ext; ++t) {
        if (t >= lo && t < hi) continue;
        su.push(P.mod_index_ext(l, t), digit * ext + t);
    }
    ntt(up, su);
    copy(out_ext, up + (size_t)digit * ext * N, (size_t)ext * N);
    copy(out_ext + (size_t)lo * N, c_eval + (size_t)lo * N, (size_t)(hi - lo) * N);
    release(dco); release(up);
}

void Engine::moddown(u64* out, const u64* in_ext, int l) {
    const int N = P.N, K = P.K, ext = l + K;
    u64* acc = alloc((size_t)ext * N);
    copy(acc, in_ext, (size_t)ext * N);
    LimbSel sp;
    for (int k = 0; k < K; ++k) sp.push(P.L + k, l + k);
    launch_intt(T, acc, sp, 1, 0, md_.post, md_.post_sh, stream);
    u64* tq = alloc((size_t)l * N);
    launch_moddown_conv(T, md_, tq, acc + (size_t)l * N, (size_t)ext * N, l, 1, 1, 0, 0, stream);
    FinishArgs fa{out, 0, acc, (size_t)ext * N, 0, tq, 0, nullptr, 0, nullptr, 0, nullptr, 0};
    ntt_finish(tq, 0, fa, 0, l, 1, 1);
    release(acc); release(tq);
}

}  // namespace flk
