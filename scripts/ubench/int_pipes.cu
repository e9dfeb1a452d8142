// Integer-pipe throughput probe for sm_100a: warp-instructions per cycle per SM for the operations a 64-bit Shoup butterfly is made of.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o int_pipes int_pipes.cu ; run on the B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned int u32; typedef unsigned long long u64;
#define ITER 4096
#define CHAINS 8
template <int OP> __global__ void probe(u32* out, u32 a0, u32 b0) {
    u32 x[CHAINS], y[CHAINS]; u64 w[CHAINS];
    for (int i = 0; i < CHAINS; ++i) { x[i] = a0 + i + threadIdx.x; y[i] = b0 * (i + 3); w[i] = ((u64)x[i] << 32) | y[i]; }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(a0));                 // IMAD
            if (OP == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(a0));                 // IMAD.HI.U32
            if (OP == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(y[i]));             // IMAD.WIDE.U32
            if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));                                  // IADD3
            if (OP == 4) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %2;" : "+r"(x[i]), "+r"(y[i]) : "r"(a0)); // IADD3 + IADD3.X
            if (OP == 5) asm volatile("mul.hi.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) % CHAINS] | 1));              // 64-bit mulhi (emulated)
            if (OP == 6) asm volatile("mul.lo.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) % CHAINS] | 1));              // 64-bit mullo (emulated)
            if (OP == 7) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(a0)); asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(b0)); }  // IMAD + IADD3 mix
            if (OP == 8) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(a0));            // LOP3
            if (OP == 9) asm volatile("{.reg .pred p; setp.ge.u32 p, %0, %1; selp.u32 %0, %1, %0, p;}" : "+r"(x[i]) : "r"(y[i]));  // ISETP + SEL
            if (OP == 10) asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));                               // IMNMX/VIMNMX
            if (OP == 11) asm volatile("min.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) % CHAINS]));               // 64-bit min
            if (OP == 12) asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) % CHAINS]));               // 64-bit add
        }
    }
    u32 s = 0; for (int i = 0; i < CHAINS; ++i) s += x[i] + y[i] + (u32)w[i] + (u32)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char* name, int ops_per) {
    u32* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<OP><<<148 * 2, 1024>>>(out, 1, 3); cudaDeviceSynchronize();
    cudaEventRecord(e0); probe<OP><<<148 * 2, 1024>>>(out, 1, 3); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_inst = 148.0 * 2 * 32 * (double)ITER * CHAINS * ops_per;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s %8.3f ms  %6.2f warp-inst/clk/SM (at %d MHz nominal)\n", name, ms, warp_inst / cycles / 148.0, clk / 1000);
    cudaFree(out);
}
int main() {
    run<0>("IMAD (mad.lo.u32)", 1); run<1>("IMAD.HI.U32", 1); run<2>("IMAD.WIDE.U32", 1); run<3>("IADD3", 1); run<4>("IADD3+IADD3.X pair", 2);
    run<5>("mul.hi.u64 (per PTX op)", 1); run<6>("mul.lo.u64 (per PTX op)", 1); run<7>("IMAD+IADD3 mix", 2); run<8>("LOP3", 1);
    run<9>("ISETP+SEL", 2); run<10>("min.u32", 1); run<11>("min.u64 (per PTX op)", 1); run<12>("add.u64 (per PTX op)", 1);
    return 0;
}
