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

// floor(a*b / 2^64) - e with e in {0,1,2}: the three partial products that reach the high word (IMAD.WIDE.U32 x3);
// the low x low product and the low halves of the cross products are dropped.
__device__ __forceinline__ u64 mulhi_lazy(u64 a, u64 b) {
    const u32 al = (u32)a, ah = (u32)(a >> 32), bl = (u32)b, bh = (u32)(b >> 32);
    const u64 t = (u64)ah * bl;
    const u64 u = (u64)al * bh;
    return (u64)ah * bh + (t >> 32) + (u >> 32);
}

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

// ---- carry-free multiply-accumulate for the base conversions and the evaluation-key inner product ----
// Every modulus is below 2^60, so a residue splits into two 30-bit halves and a product into four partial products below
// 2^60.  Up to 8 products (16 cross terms) accumulate in plain 64-bit registers -- one IMAD.WIDE.U32 each, no carry chain:
//   sum = a0 + a1 2^30 + a2 2^60,  a0 < 2^63, a1 < 2^64, a2 < 2^63.
struct Split30 {
    u32 lo, hi;
};
__device__ __forceinline__ Split30 split30(u64 x) { return Split30{(u32)x & 0x3fffffffu, (u32)(x >> 30)}; }
struct Acc3 {
    u64 a0, a1, a2;
};
// acc += a * b as one IMAD.WIDE.U32 (written in PTX: the C form compiles to a product plus separate 64-bit additions)
__device__ __forceinline__ void wmad(u64& acc, u32 a, u32 b) { asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b)); }
__device__ __forceinline__ void mac3(Acc3& s, Split30 x, Split30 y) {
    wmad(s.a0, x.lo, y.lo);
    wmad(s.a1, x.lo, y.hi);
    wmad(s.a1, x.hi, y.lo);
    wmad(s.a2, x.hi, y.hi);
}
// conditional subtraction for operands below 2^63 (sign test instead of a 64-bit compare)
__device__ __forceinline__ u64 csub_s(u64 a, u64 q) {
    const u64 t = a - q;
    return (long long)t < 0 ? a : t;
}
// a*w - qhat*q (mod 2^64) with qhat = the three high partial products of a*ws (low by at most 2): any a < 2^64, result in
// [0,4q); nq = -q.  Written on 32-bit halves in PTX: 5 IMAD.WIDE.U32 + 4 IMAD + 4 carry adds, nothing else -- the C form
// costs ~3 more SASS instructions per call in zero-extension moves.
__device__ __forceinline__ u64 shoup_lazy4(u64 a, u64 w, u64 ws, u64 nq) {
    u64 r;
    asm("{\n\t.reg .u32 yl, yh, wl, wh, sl, sh, nl, nh, t1, t2, ql, qh, c;\n\t.reg .u64 t, u, h, p;\n\t"
        "mov.b64 {yl, yh}, %1; mov.b64 {wl, wh}, %2; mov.b64 {sl, sh}, %3; mov.b64 {nl, nh}, %4;\n\t"
        "mul.wide.u32 t, yh, sl;\n\t"
        "mul.wide.u32 u, yl, sh;\n\t"
        "mul.wide.u32 h, yh, sh;\n\t"
        "mov.b64 {c, t1}, t; mov.b64 {c, t2}, u; mov.b64 {ql, qh}, h;\n\t"
        "add.cc.u32 ql, ql, t1; addc.u32 qh, qh, 0;\n\t"
        "add.cc.u32 ql, ql, t2; addc.u32 qh, qh, 0;\n\t"
        "mul.wide.u32 p, yl, wl;\n\t"
        "mad.wide.u32 p, ql, nl, p;\n\t"
        "mov.b64 {c, t1}, p;\n\t"
        "mad.lo.u32 t1, yl, wh, t1;\n\t"
        "mad.lo.u32 t1, yh, wl, t1;\n\t"
        "mad.lo.u32 t1, ql, nh, t1;\n\t"
        "mad.lo.u32 t1, qh, nl, t1;\n\t"
        "mov.b64 %0, {c, t1};\n\t}"
        : "=l"(r) : "l"(a), "l"(w), "l"(ws), "l"(nq));
    return r;
}
// per-modulus constants of the accumulator reductions
struct __align__(16) RedC {
    u64 q, nq, mu64;        // mu64 = floor(2^64 / q)
    u64 c30, c30s;          // 2^30 mod q and its Shoup companion
    u64 c60, c60s;          // 2^60 mod q
    u64 mu94;               // floor(2^94 / q)  (q > 2^30)
};
// Two-accumulator sums X = b0 + b1 2^30 < 2^94 (b1 + (b0 >> 30) < 2^64): one Barrett step on the top 64 bits.
// floor(X / 2^30) = b1 + (b0 >> 30) exactly; qhat = hi64(that * floor(2^94/q)) from three partial products is
// floor(X/q) - e with e in 0..4, so X - qhat q (computed modulo 2^64) lies in [0, 5q): 6 IMAD-class + ~12 ALU instructions
// against 24 + 14 for the three-accumulator form below.
struct Acc2 {
    u64 b0, b1;
};
// acc += y * h with y as 30-bit halves and the constant given twice: h = (h.x, h.y) and h 2^30 mod q = (h.z, h.w), all 30-bit
// halves.  Up to 8 terms (16 products below 2^60 per accumulator) fit without carries.
__device__ __forceinline__ void mac2(Acc2& s, Split30 y, uint4 h) {
    wmad(s.b0, y.lo, h.x);
    wmad(s.b1, y.lo, h.y);
    wmad(s.b0, y.hi, h.z);
    wmad(s.b1, y.hi, h.w);
}
__device__ __forceinline__ u64 reduce2_lazy(u64 b0, u64 b1, const RedC& k) {    // [0, 5q)
    const u64 xh = b1 + (b0 >> 30);
    const u64 xl = b0 + (b1 << 30);
    return xl + mulhi_lazy(xh, k.mu94) * k.nq;
}
__device__ __forceinline__ u64 reduce2(u64 b0, u64 b1, const RedC& k) {         // canonical
    u64 r = reduce2_lazy(b0, b1, k);
    r = csub_s(r, k.q << 2);
    r = csub_s(r, k.q << 1);
    return csub_s(r, k.q);
}
// (a0 + a1 2^30 + a2 2^60) mod q for sums of up to 8 products (mac3): the top accumulator is folded through a lazy Shoup
// product with 2^60 mod q (< 4q < 2^62, so a0 + it stays below 2^64), the rest is the two-accumulator step above
// (a1 + ((a0 + fold) >> 30) < 2^64 because a1 <= 16 (2^30 - 1)^2).
__device__ __forceinline__ u64 reduce3_lazy(const Acc3& s, const RedC& k) {
    return reduce2_lazy(s.a0 + shoup_lazy4(s.a2, k.c60, k.c60s, k.nq), s.a1, k);
}
__device__ __forceinline__ u64 reduce3(const Acc3& s, const RedC& k) {
    return reduce2(s.a0 + shoup_lazy4(s.a2, k.c60, k.c60s, k.nq), s.a1, k);
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
