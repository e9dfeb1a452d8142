// modarith.cuh -- 64-bit modular arithmetic for sm_100a (integer pipes only: IMAD on fma, IADD3/SHF on alu).
// All moduli are < 2^61 so lazily reduced values in [0,4q) fit a 64-bit word.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace flk {
using u64 = uint64_t;
using u32 = uint32_t;

namespace dev {

__device__ __forceinline__ u64 csub(u64 a, u64 q) { return a >= q ? a - q : a; }
__device__ __forceinline__ u64 addmod(u64 a, u64 b, u64 q) { return csub(a + b, q); }
__device__ __forceinline__ u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }

// a*w mod q with Shoup companion ws = floor(w*2^64/q); any a < 2^64; result in [0,2q)
__device__ __forceinline__ u64 mul_shoup_lazy(u64 a, u64 w, u64 ws, u64 q) {
    return a * w - __umul64hi(a, ws) * q;
}
__device__ __forceinline__ u64 mul_shoup(u64 a, u64 w, u64 ws, u64 q) { return csub(mul_shoup_lazy(a, w, ws, q), q); }

struct U128 {
    u64 lo, hi;
};
__device__ __forceinline__ void mad128(U128& acc, u64 a, u64 b) {
    u64 lo = a * b, hi = __umul64hi(a, b);
    acc.lo += lo;
    acc.hi += hi + (acc.lo < lo);
}
// x mod q for any 128-bit x; mu = floor(2^128/q) = (mu_hi, mu_lo).  Quotient estimate is low by at most 3.
__device__ __forceinline__ u64 barrett128(U128 x, u64 q, u64 mu_lo, u64 mu_hi) {
    u64 qh = x.hi * mu_hi + __umul64hi(x.hi, mu_lo) + __umul64hi(x.lo, mu_hi);
    u64 r = x.lo - qh * q;
    r = csub(r, q << 1);
    r = csub(r, q << 1);
    return csub(r, q);
}
__device__ __forceinline__ u64 mulmod(u64 a, u64 b, u64 q, u64 mu_lo, u64 mu_hi) {
    U128 x{a * b, __umul64hi(a, b)};
    return barrett128(x, q, mu_lo, mu_hi);
}

// Harvey lazy butterflies.  Cooley-Tukey (forward): x,y in [0,4q) -> [0,4q)
__device__ __forceinline__ void ct_bfly(u64& x, u64& y, u64 w, u64 ws, u64 q) {
    const u64 q2 = q << 1;
    u64 u = csub(x, q2);
    u64 v = mul_shoup_lazy(y, w, ws, q);
    x = u + v;
    y = u - v + q2;
}
// Gentleman-Sande (inverse): x,y in [0,2q) -> [0,2q)
__device__ __forceinline__ void gs_bfly(u64& x, u64& y, u64 w, u64 ws, u64 q) {
    const u64 q2 = q << 1;
    u64 s = csub(x + y, q2);
    u64 d = x - y + q2;
    x = s;
    y = mul_shoup_lazy(d, w, ws, q);
}

}  // namespace dev
}  // namespace flk
