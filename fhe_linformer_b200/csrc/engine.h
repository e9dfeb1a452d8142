// engine.h -- device engine: owns the parameter tables in HBM, the stream, and the primitive
// operations on raw limb-major device buffers.  The CKKS scheme layer (scheme.h) and the C-ABI sit on top.
#pragma once
#include <map>
#include <memory>
#include <unordered_map>
#include <vector>

#include "device_ctx.h"
#include "kernels.cuh"

namespace flk {

struct OpLedger {   // algorithmic-byte accounting per SURVEY.md section 8(d)
    struct Row { long count = 0; double bytes = 0; };
    std::map<std::string, Row> rows;
    void add(const std::string& op, int l, double bytes, int times = 1) {
        Row& r = rows[op + "@" + std::to_string(l)];
        r.count += times; r.bytes += bytes * times;
    }
    void reset() { rows.clear(); }
};

// A batch of key switches sharing one evaluation key; *_bs are batch strides in words (0 when B = 1).
struct KsBatch {
    int B, l;
    const u64* c; size_t c_bs;          // polynomials to switch, [l][N] each
    u64* out; size_t out_bs;            // results, [2][l][N] each
    const u64* add0; size_t add0_bs;    // optional [l][N] addends of the two result polynomials (permuted with them)
    const u64* add1; size_t add1_bs;
    const u64* plus; size_t plus_bs;    // optional [2][l][N] addend that is NOT permuted (rotate-and-add)
};

// BSGS plan of a diagonal linear transform  out = sum_j sigma_{G_j}( sum_i pt[j][i] * sigma_{g_i}(ct) )  (double hoisting).
// Baby step 0 is the identity; a giant step without rotation (g = 1), if any, is stored last.
struct LtPlan {
    int n1 = 1, n2 = 1, l = 0;
    uint32_t baby_g[kBsgsMax] = {};  const u64* baby_evk[kBsgsMax] = {};
    uint32_t giant_g[kBsgsMax] = {}; const u64* giant_evk[kBsgsMax] = {};
    uint32_t mask[kBsgsMax] = {};    // per giant step: bit i set when plaintext (j, i) is present
    const u64* pts = nullptr;        // [n2][n1][l+K][N] plaintext diagonals over Q_l u P, evaluation form
    int ndiag = 0;                   // present diagonals (ledger)
};

class Engine {
public:
    explicit Engine(const ParamSpec& spec, int device = -1);
    ~Engine();
    Engine(const Engine&) = delete;

    const Params P;
    DevTables T{};
    cudaStream_t stream = nullptr;
    int device_id = 0;          // the device this engine lives on; every C-ABI entry binds the calling thread to it (capi.cpp)
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;   // copy engines of the host-operand pipeline (created on first use)
    OpLedger ledger;
    bool ledger_on = false;

    // memory (stream-ordered pool)
    u64* alloc(size_t words);
    void release(u64* p);
    void upload(u64* dst, const u64* src, size_t words);
    void download(u64* dst, const u64* src, size_t words);
    void copy(u64* dst, const u64* src, size_t words);
    void sync();

    // --- primitives on device buffers (all asynchronous on `stream`) ---
    void ntt(u64* data, const LimbSel& sel, int batch = 1, size_t batch_stride = 0);
    void ntt_finish(u64* tq, size_t tq_bs, const FinishArgs& fa, uint32_t g, int l, int polys, int B);   // NTT of the ModDown conversion + fused finish
    void intt(u64* data, const LimbSel& sel, int batch = 1, size_t batch_stride = 0);
    void ew(EwOp op, u64* out, const u64* a, const u64* b, int l, int polys, bool broadcast_b);
    void ew_sel(EwOp op, u64* out, const u64* a, const u64* b, const LimbSel& sel);
    void automorph(u64* out, const u64* in, uint32_t g, int limbs);
    void rescale(u64* out, const u64* in, int l, int polys);
    void ntt_l2(u64* data, const LimbSel& sel, int batch, size_t batch_stride);   // batched forward NTT in L2-sized sub-batches
    long l2_budget_bytes = 1l << 40;   // sub-batching measured slower on B200 (2460 vs 2625 rot/s): off unless FLK_L2_BUDGET_MB is set
    // out[2][l][N] = KeySwitch(c) (+add0 / +add1), optionally permuted by the automorphism map of g (0 = none)
    void keyswitch(u64* out, const u64* c, const u64* evk, int l, const u64* add0, const u64* add1, uint32_t g);
    void keyswitch(const KsBatch& io, const u64* evk, uint32_t g);
    void rotate_batch(u64* out, const u64* ct, int l, uint32_t g, const u64* evk, int B, bool accumulate);   // ct, out: [B][2][l][N]
    // out[b] = plan(ct[b]) for B ciphertexts stored back to back: one ModUp for all baby rotations, plaintext products in the
    // extended basis, one ModDown per giant step, giant rotations summed before a single final ModDown
    void linear_transform(u64* out, const u64* ct, int B, const LtPlan& plan);
    void modup_ntt(const KsLevel& ks, u64* up, const u64* dco, int B, size_t up_bs, size_t dco_bs);   // base conversion + NTT of all digits
    // out[b] = (self ? ct[b] : 0) + sum_k rotate(ct[b], g_k): the rotations share one ModUp and one ModDown (hoisting); nk <= kHoistMax
    void rotate_sum_batch(u64* out, const u64* ct, int l, const uint32_t* gs, const u64* const* evks, int nk, int B, bool self);
    // the same with HOST operands: uploads, key switches and downloads of successive chunks overlap on three streams
    void rotate_batch_host(u64* out_host, const u64* ct_host, int l, uint32_t g, const u64* evk, int B, int chunk, bool wait = true);
    void rotate(u64* out, const u64* ct, int l, uint32_t g, const u64* evk);
    void rotate_add(u64* out, const u64* ct, int l, uint32_t g, const u64* evk);   // out = ct + rotate(ct)
    void mul_relin(u64* out, const u64* a, const u64* b, int l, const u64* evk);
    void mul_relin_batch(u64* out, const u64* a, const u64* b, int l, const u64* evk, int B, size_t b_bs);   // a, out: [B][2][l][N]
    // limb-sharded key switch: one stage each, restricted to [first, first + count) of the extended basis (engine.cu)
    void ks_digits_part(u64* dco, const u64* c, int l, int first, int count);
    void ks_digits(u64* dco, const u64* c, int l);
    void ks_modup_part(u64* up, const u64* dco, int l, int first, int count);
    void ks_inner_part(u64* acc, const u64* up, const u64* c, const u64* evk, int l, int first, int count);
    void ks_pcoef_part(u64* acc, int l, int first, int count);
    void ks_moddown_part(u64* out, u64* tq, const u64* acc, int l, int first, int count, const u64* add0, const u64* add1, uint32_t g);
    // pieces exposed for parity tests
    void modup(u64* out_ext, const u64* c_eval, int l, int digit);     // out: (l+K) limbs eval
    void moddown(u64* out, const u64* in_ext, int l);                  // in: (l+K) limbs eval

    const uint32_t* automorph_map(uint32_t g);
    const KsLevel& ks_level(int l);
    const MdConst& md() const { return md_; }
    const RsConst& rs() const { return rs_; }
    size_t evk_words() const { return (size_t)P.dnum * 2 * P.T * P.N; }

    void trim_cache(size_t keep_bytes);
    // block-cache cap (a context parameter: fl_ctx_set_cache_bytes); several controllers on one GPU share its 180 GB
    void set_cache_cap(size_t bytes) { cache_cap_bytes_ = bytes; if (cached_bytes_ > cache_cap_bytes_) trim_cache(cache_cap_bytes_ / 8 * 7); }
    size_t cache_cap() const { return cache_cap_bytes_; }
    long pool_allocs = 0, cache_trims = 0;               // allocator statistics (fl_ctx_info slots 5-7)
    size_t cached_bytes() const { return cached_bytes_; }

private:
    // the tail of a key switch: acc [Bn][2][l+K][N] (eval) -> out [Bn][2][l][N] = (acc - conv_P->Q(acc_P)) P^-1 + addends
    void moddown_acc(u64* out, size_t out_bs, u64* acc, int Bn, int l, const u64* add0, size_t add0_bs, const u64* add1, size_t add1_bs,
                     const u64* plus, size_t plus_bs, uint32_t g);
    void modup_batch(u64* up, const u64* c, size_t c_bs, int Bn, int l);   // digits + ModUp + NTT of Bn polynomials [l][N] -> [Bn][beta][l+K][N]
    ScalarSet pmod_{};           // P mod q_i with Shoup companions
    struct HostSlot { u64* in = nullptr; u64* out = nullptr; size_t words = 0; cudaEvent_t in_consumed = nullptr, out_drained = nullptr; };
    HostSlot host_slots_[2];     // device staging of the host-operand pipeline, alternated by successive calls
    unsigned host_next_ = 0;
    std::map<size_t, std::vector<u64*>> free_blocks_;    // exact-size cache in front of the stream-ordered pool
    std::unordered_map<u64*, size_t> block_size_;
    size_t cached_bytes_ = 0, cache_cap_bytes_ = (size_t)96 << 30;   // of the 180 GB; set_cache_cap() / fl_ctx_set_cache_bytes
    std::vector<void*> owned_;
    template <class V> V* to_device(const std::vector<V>& h);
    std::unordered_map<int, KsLevel> ks_;
    std::unordered_map<uint32_t, uint32_t*> maps_;
    MdConst md_{};
    RsConst rs_{};
};

}  // namespace flk
