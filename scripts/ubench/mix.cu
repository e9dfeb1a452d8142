// Does integer ALU work (IADD3 / IADD3.X / LOP3 / SHF) overlap with IMAD-class work on sm_100a, or do their issue costs add up?
// Loop body = NM independent-chain IMAD.WIDE.U32 (64-bit accumulate) + NA ALU instructions of a given kind, per thread.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mix mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned int u32; typedef unsigned long long u64;
#define ITER 2048
template <int NM, int NA, int KIND> __global__ void probe(u32* out, u32 a0, u32 b0) {
    u64 w[8]; u32 x[8], y[8], z[16];
    for (int i = 0; i < 8; ++i) { x[i] = a0 + i + threadIdx.x; y[i] = b0 * (i + 3); w[i] = ((u64)x[i] << 32) | y[i]; }
    for (int i = 0; i < 16; ++i) z[i] = a0 * (i + 7) + threadIdx.x;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < (NM > NA ? NM : NA); ++i) {
            if (i < NM) asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(w[i % 8]) : "r"(y[i % 8]));   // multiplicand = own low word: not loop-invariant
            if (i < NA) {
                if (KIND == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(z[i % 16]) : "r"(b0));                                       // IADD3, independent of the IMADs
                if (KIND == 1) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %2;" : "+r"(z[i % 16]), "+r"(z[(i + 8) % 16]) : "r"(b0));  // carry pair (2 instr)
                if (KIND == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i % 16]) : "r"(b0), "r"(a0));                  // LOP3
                if (KIND == 3) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(z[i % 16]) : "r"(z[(i + 1) % 16]), "r"(a0));     // SHF
                if (KIND == 4) { u32 hi = (u32)(w[i % 8] >> 32); asm volatile("add.u32 %0, %0, %1;" : "+r"(z[i % 16]) : "r"(hi)); }    // IADD3 fed by the IMAD chain
                if (KIND == 5) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(z[i % 16]) : "r"(b0), "r"(a0));                        // plain IMAD instead of ALU
            }
        }
    }
    u32 s = 0; for (int i = 0; i < 8; ++i) s += x[i] + y[i] + (u32)w[i] + (u32)(w[i] >> 32);
    for (int i = 0; i < 16; ++i) s += z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NM, int NA, int KIND> void run(const char* name, int per_alu) {
    u32* out; cudaMalloc(&out, 148 * 4 * 512 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<NM, NA, KIND><<<148 * 4, 512>>>(out, 1, 3); cudaDeviceSynchronize();
    cudaEventRecord(e0); probe<NM, NA, KIND><<<148 * 4, 512>>>(out, 1, 3); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3;
    const double warps_per_smsp = 4.0 * 512 / 32 / 4;
    printf("%-10s IMAD.WIDE x%2d + ALU x%2d : %6.2f cycles per loop body per warp per SMSP  (%5.2f inst/clk/SMSP)\n", name, NM, NA * per_alu,
           cycles / (ITER * warps_per_smsp), (NM + NA * per_alu) * ITER * warps_per_smsp / cycles);
    cudaFree(out);
}
int main() {
    run<8, 0, 0>("imad only", 1); run<0, 8, 0>("iadd only", 1); run<0, 16, 0>("iadd only", 1);
    run<8, 4, 0>("iadd", 1); run<8, 8, 0>("iadd", 1); run<8, 12, 0>("iadd", 1); run<8, 16, 0>("iadd", 1);
    run<8, 2, 1>("carry", 2); run<8, 4, 1>("carry", 2); run<8, 8, 1>("carry", 2); run<0, 8, 1>("carry", 2);
    run<8, 4, 2>("lop3", 1); run<8, 8, 2>("lop3", 1); run<0, 8, 2>("lop3", 1);
    run<8, 4, 3>("shf", 1); run<8, 8, 3>("shf", 1); run<0, 8, 3>("shf", 1);
    run<8, 4, 4>("iadd dep", 1); run<8, 8, 4>("iadd dep", 1);
    run<8, 4, 5>("imad", 1); run<8, 8, 5>("imad", 1);
    return 0;
}
