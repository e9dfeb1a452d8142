// Cost (cycles per warp-op per SM sub-partition) of the pieces of a 64-bit Shoup butterfly on sm_100a, 8 independent chains per thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
__device__ __forceinline__ u64 mulhi_lazy(u64 a, u64 b) {
    const u32 al = (u32)a, ah = (u32)(a >> 32), bl = (u32)b, bh = (u32)(b >> 32);
    const u64 t = (u64)ah * bl; const u64 u = (u64)al * bh;
    return (u64)ah * bh + (t >> 32) + (u >> 32);
}
template <int OP> __global__ void k(u64* out, u64 c0, u64 c1, u64 c2, int iters) {
    u64 x[8], y[8];
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 8 + i + c0; y[i] = x[i] * 3 + c1; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) x[i] = x[i] * c1 + c2;                           // mul.lo.u64 + add (3 IMAD-class)
            if (OP == 1) x[i] = mulhi_lazy(x[i], c1) + c2;                // lazy high product
            if (OP == 2) x[i] = __umul64hi(x[i], c1) + c2;                // exact high product
            if (OP == 3) x[i] = x[i] + y[i];                              // 64-bit add
            if (OP == 4) { u64 t = x[i]; x[i] = t + y[i]; y[i] = t - y[i] + c2; }   // butterfly add/sub
            if (OP == 5) x[i] = x[i] * c1 + mulhi_lazy(x[i], c2) * c0;    // lazy Shoup product
            if (OP == 6) { const u64 v = y[i] * c1 + mulhi_lazy(y[i], c2) * c0; const u64 t = x[i]; x[i] = t + v; y[i] = t - v + c2; }  // whole butterfly
            if (OP == 7) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(x[i]) : "r"((u32)x[i]), "r"((u32)c1)); }    // IMAD.WIDE, RZ addend
            if (OP == 8) { u64 t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"((u32)y[i]), "r"((u32)c1)); x[i] += (t >> 32); }
        }
    }
    u64 s = 0; for (int i = 0; i < 8; ++i) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char* name) {
    const int threads = 512, bps = 2, iters = 4000;
    u64* out; cudaMalloc(&out, (size_t)148 * bps * threads * 8);
    k<OP><<<148 * bps, threads>>>(out, 0xFFFFFFF000000123ull, 0x000FFFFF12345677ull, 0x3FFFFFFFFFFF14ull, 10);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<OP><<<148 * bps, threads>>>(out, 0xFFFFFFF000000123ull, 0x000FFFFF12345677ull, 0x3FFFFFFFFFFF14ull, iters); cudaEventRecord(b);
    cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b);
    double ops_per_smsp = (double)bps * threads / 32 / 4 * iters * 8;
    printf("%-44s %6.2f cycles per warp-op per SMSP\n", name, ms * 1e-3 * 1.965e9 / ops_per_smsp);
    cudaFree(out);
}
int main() {
    run<0>("64-bit a*b+c low word (mul.lo.u64)"); run<1>("lazy high product (3 IMAD.WIDE)"); run<2>("exact high product (__umul64hi)");
    run<3>("64-bit add"); run<4>("butterfly add/sub (x+y, x-y+4q)"); run<5>("lazy Shoup product"); run<6>("whole lazy butterfly");
    run<7>("IMAD.WIDE.U32 (mul.wide)"); run<8>("mul.wide + add of its high word");
    return 0;
}
