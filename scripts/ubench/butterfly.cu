// Arithmetic-only throughput of the lazy Shoup butterfly on sm_100a: radix-8 register blocks in a loop, no memory traffic.
// Reports cycles per warp-butterfly per SM sub-partition at several occupancies.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
__device__ __forceinline__ u64 mulhi_lazy(u64 a, u64 b) {
    const u32 al = (u32)a, ah = (u32)(a >> 32), bl = (u32)b, bh = (u32)(b >> 32);
    const u64 t = (u64)ah * bl; const u64 u = (u64)al * bh;
    return (u64)ah * bh + (t >> 32) + (u >> 32);
}
template <int MODE> __device__ __forceinline__ u64 shoup(u64 a, u64 w, u64 ws, u64 nq) {
    if (MODE == 0) return a * w + mulhi_lazy(a, ws) * nq;
    return a * w + __umul64hi(a, ws) * nq;
}
template <int MODE> __device__ __forceinline__ void bf(u64& x, u64& y, u64 w, u64 ws, u64 nq, u64 q4) {
    const u64 v = shoup<MODE>(y, w, ws, nq); const u64 u = x; x = u + v; y = u - v + q4;
}
template <int MODE> __global__ void k(u64* out, u64 w0, u64 ws0, u64 nq, u64 q4, int iters) {
    u64 e[8]; u64 w[7], ws[7];
    for (int i = 0; i < 8; ++i) e[i] = threadIdx.x * 8 + i + w0;
    for (int i = 0; i < 7; ++i) { w[i] = w0 + i * 977; ws[i] = ws0 + i * 131; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < 3; ++s) { const int half = 4 >> s;
#pragma unroll
            for (int g = 0; g < (1 << s); ++g)
#pragma unroll
                for (int j = 0; j < half; ++j) bf<MODE>(e[g * 2 * half + j], e[g * 2 * half + half + j], w[(1 << s) - 1 + g], ws[(1 << s) - 1 + g], nq, q4);
        }
    }
    u64 s = 0; for (int i = 0; i < 8; ++i) s += e[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int threads, int blocks_per_sm) {
    u64* out; cudaMalloc(&out, (size_t)148 * blocks_per_sm * threads * 8);
    const int iters = 2000;
    k<MODE><<<148 * blocks_per_sm, threads>>>(out, 12345, 678, 0 - 0xFFFFFFFFFFFC5ull, 4 * 0xFFFFFFFFFFFC5ull, 10);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<MODE><<<148 * blocks_per_sm, threads>>>(out, 12345, 678, 0 - 0xFFFFFFFFFFFC5ull, 4 * 0xFFFFFFFFFFFC5ull, iters); cudaEventRecord(b);
    cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b);
    double warp_bf_per_smsp = (double)blocks_per_sm * threads / 32 / 4 * iters * 12;
    double cycles = ms * 1e-3 * 1.965e9;
    printf("%-22s %4d thr x %d CTA/SM (%2d warps/SMSP): %6.2f cycles per warp-butterfly per SMSP  -> %.2f T butterflies/s chip\n", name, threads, blocks_per_sm,
           blocks_per_sm * threads / 128, cycles / warp_bf_per_smsp, 148.0 * blocks_per_sm * threads * iters * 12 / (ms * 1e-3) / 1e12);
    cudaFree(out);
}
int main() {
    run<0>("lazy 3-product", 128, 1); run<0>("lazy 3-product", 256, 1); run<0>("lazy 3-product", 512, 1); run<0>("lazy 3-product", 512, 2); run<0>("lazy 3-product", 1024, 2);
    run<1>("exact __umul64hi", 512, 2); run<1>("exact __umul64hi", 1024, 2);
    return 0;
}
