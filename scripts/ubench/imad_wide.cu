// Is IMAD.WIDE.U32 (32x32 -> 64, the workhorse of every 64-bit modular multiply) half the rate of IMAD (low 32 bits) on sm_100a?
// Dependent chains whose multiplicands change every iteration (nothing for ptxas to hoist), 8 chains per thread, 16 warps per
// sub-partition, kernels long enough (>= 20 ms) for the clock to settle; the ratio of the rows is the result.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o imad_wide imad_wide.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
template <int OP> __global__ void k(u64* out, u32 c0, u32 c1, int iters) {
    u64 x[8]; u32 a[8], b[8];
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 8 + i + c0; a[i] = threadIdx.x + i * c1; b[i] = c0 + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c1));                       // IMAD
            if (OP == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"((u32)x[i]), "r"(b[i]));              // IMAD.WIDE, 64-bit addend
            if (OP == 2) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(x[i]) : "r"((u32)x[i]), "r"((u32)(x[i] >> 32) | 1u)); // IMAD.WIDE, no addend
            if (OP == 3) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));                                    // IMAD.HI
            if (OP == 4) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"((u32)x[i]), "r"(b[i]));
                           asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c1)); }                   // one of each
            if (OP == 5) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"((u32)x[i]), "r"(b[i]));
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c1));
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(b[i]) : "r"(a[i]), "r"(c0));
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c0)); }               // IMAD.WIDE + 3 LOP3 (ALU)
            if (OP == 8) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(x[i]) : "r"((u32)x[i]), "r"((u32)(x[i] >> 32) | 1u));
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c1));
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(b[i]) : "r"(a[i]), "r"(c0)); }               // IMAD.WIDE without addend + 2 LOP3
            if (OP == 9) { u64 t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[i]), "r"(b[i]));
                           asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(((u32*)&x[i])[0]), "+r"(((u32*)&x[i])[1]) : "r"((u32)t), "r"((u32)(t >> 32)));
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"((u32)x[i]), "r"(c1)); }                  // mul.wide + separate 64-bit accumulate (2 adds) + 1 LOP3
            if (OP == 6) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c1));
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(b[i]) : "r"(a[i]), "r"(c0)); }               // 2 LOP3
            if (OP == 7) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c1));
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(b[i]) : "r"(a[i]), "r"(c0)); }               // IMAD + LOP3
        }
    }
    u64 s = 0; for (int i = 0; i < 8; ++i) s += x[i] + a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> double run(const char* name, int per_iter, int iters) {
    const int threads = 1024, bps = 2, blocks = 148 * bps;
    static u64* out = nullptr; if (!out) cudaMalloc(&out, (size_t)blocks * threads * 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<OP><<<blocks, threads>>>(out, 0x9E3779B9u, 0x85EBCA6Bu, iters);   // warm-up at full length: lets the clock settle
    cudaEventRecord(a); k<OP><<<blocks, threads>>>(out, 0x9E3779B9u, 0x85EBCA6Bu, iters); cudaEventRecord(b);
    cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b);
    const double warp_inst_per_smsp = (double)bps * threads / 32 / 4 * iters * 8 * per_iter;
    printf("%-36s %8.3f ms  %6.3f ns per warp-instruction per sub-partition  (%.2f cycles at 1965 MHz)\n", name, ms, ms * 1e6 / warp_inst_per_smsp,
           ms * 1e6 / warp_inst_per_smsp * 1.965);
    return ms;
}
int main() {
    const int it = 40000;
    run<0>("IMAD (mad.lo.u32)", 1, it); run<0>("IMAD (mad.lo.u32)", 1, it);
    run<1>("IMAD.WIDE.U32 + 64-bit addend", 1, it); run<2>("IMAD.WIDE.U32 (mul.wide)", 1, it); run<3>("IMAD.HI.U32", 1, it);
    run<4>("IMAD.WIDE + IMAD", 2, it); run<6>("2 LOP3", 2, it); run<7>("IMAD + LOP3", 2, it); run<5>("IMAD.WIDE + 3 LOP3", 4, it);
    run<8>("IMAD.WIDE (no addend) + 2 LOP3", 3, it); run<9>("mul.wide + add.cc/addc + LOP3", 4, it);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
