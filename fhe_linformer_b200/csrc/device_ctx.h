// device_ctx.h -- device-resident tables of one CKKS parameter set + launch helpers shared by kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "params.h"

namespace flk {

#define FLK_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " + \
                                     __FILE__ + ":" + std::to_string(__LINE__));                   \
    } while (0)

constexpr int kMaxLimbSel = 256;   // limbs one launch may address: beta (l + K) - l of a ModUp; checked in LimbSel::push / Engine::ks_level
constexpr int kRadix1Log = 4;   // stages done by the register-only column pass (strides N/2 .. N/16)

// Passed by value to kernels.
struct DevTables {
    const u64* q;        // [T]
    const u64* mu_lo;    // [T] floor(2^128/q) low / high words
    const u64* mu_hi;
    const ulonglong2* tw2;    // [T][N] {psi^bitrev(i), its Shoup companion}
    const ulonglong2* itw2;   // [T][N] {psi^-bitrev(i), Shoup companion}
    const u64* ninv;     // [T]
    const u64* ninv_sh;
    const u64* redc;     // [T][8]: dev::RedC per modulus (q, -q, floor(2^64/q), 2^30 mod q, 2^60 mod q with Shoup companions)
    int logN, N, L, K;
};

// modulus index of every limb of a buffer (limb-major [n][N])
struct LimbSel {
    int n = 0;
    uint8_t m[kMaxLimbSel];      // modulus index (a chain holds at most 255 moduli: Params checks L + K)
    uint16_t pos[kMaxLimbSel];   // limb slot inside the buffer (NTT kernels only; identity elsewhere)
    // bounds-checked append: a parameter set whose key switch needs more limbs than one launch can address is an error, not a
    // silent overrun of this by-value struct
    void push(int mod, int slot) {
        if (n >= kMaxLimbSel || mod < 0 || mod > 255 || slot < 0 || slot > 65535)
            throw std::invalid_argument("limb selection overflow: " + std::to_string(n + 1) + " limbs (modulus " + std::to_string(mod) + ", slot " +
                                        std::to_string(slot) + ") exceed the " + std::to_string(kMaxLimbSel) + "-limb launch descriptor");
        m[n] = (uint8_t)mod; pos[n] = (uint16_t)slot; ++n;
    }
};

inline LimbSel sel_range(int first_mod, int count) {
    LimbSel s;
    for (int i = 0; i < count; ++i) s.push(first_mod + i, i);
    return s;
}
// Q limbs 0..l-1 followed by the K P limbs
inline LimbSel sel_ext(int l, int L, int K) {
    LimbSel s;
    for (int i = 0; i < l; ++i) s.push(i, i);
    for (int k = 0; k < K; ++k) s.push(L + k, l + k);
    return s;
}

// ---- kernel launchers (definitions in the .cu files) ----
// In-place negacyclic NTT of `batch` buffers of sel.n limbs each (buffers batch_stride words apart).
void launch_ntt(const DevTables& t, u64* data, const LimbSel& sel, int batch, size_t batch_stride, cudaStream_t s);
// Inverse; multiplies by post[m] (Shoup pair post_sh[m]), indexed by modulus, instead of N^-1 when post != nullptr.
// src != nullptr: out of place -- the input limbs are read from src (same limb slots, batches src_bs words apart), data receives the result.
void launch_intt(const DevTables& t, u64* data, const LimbSel& sel, int batch, size_t batch_stride, const u64* post,
                 const u64* post_sh, cudaStream_t s, const u64* src = nullptr, size_t src_bs = 0);

}  // namespace flk
