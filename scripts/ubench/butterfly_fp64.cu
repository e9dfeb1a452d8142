// Which pipe should carry the 52-bit limbs' butterflies on sm_100a (B200: FP64 at half the FP32 rate, unlike B300)?
// Register-only loops, cycles taken with clock64() inside the kernel (independent of the clock the card runs at):
//   * pipe probes: DFMA / DADD / DMUL / IMAD / IMAD.WIDE with a 64-bit addend / IADD3 and pairwise mixes;
//   * the integer lazy Shoup butterfly as the NTT kernels use it (PTX, 9 IMAD-class + carries);
//   * an FP64 butterfly on balanced residues held as doubles: exact product by DMUL + DFMA, quotient by a magic-number
//     rounding of y * (w/p), remainder by DFMA, one conditional +-p per output (ALU-assisted or FP64-only);
//   * both kinds in one kernel (half of the warps of every sub-partition each) -- do the pipes overlap?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o butterfly_fp64 butterfly_fp64.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;

#define MAGIC 6755399441055744.0   // 1.5 * 2^52

// ---- integer butterfly (same PTX as csrc/modarith.cuh shoup_lazy4) ----
__device__ __forceinline__ u64 shoup_lazy4(u64 a, u64 w, u64 ws, u64 nq) {
    u64 r;
    asm("{\n\t.reg .u32 yl, yh, wl, wh, sl, sh, nl, nh, t1, t2, ql, qh, c;\n\t.reg .u64 t, u, h, p;\n\t"
        "mov.b64 {yl, yh}, %1; mov.b64 {wl, wh}, %2; mov.b64 {sl, sh}, %3; mov.b64 {nl, nh}, %4;\n\t"
        "mul.wide.u32 t, yh, sl;\n\tmul.wide.u32 u, yl, sh;\n\tmul.wide.u32 h, yh, sh;\n\t"
        "mov.b64 {c, t1}, t; mov.b64 {c, t2}, u; mov.b64 {ql, qh}, h;\n\t"
        "add.cc.u32 ql, ql, t1; addc.u32 qh, qh, 0;\n\tadd.cc.u32 ql, ql, t2; addc.u32 qh, qh, 0;\n\t"
        "mul.wide.u32 p, yl, wl;\n\tmad.wide.u32 p, ql, nl, p;\n\tmov.b64 {c, t1}, p;\n\t"
        "mad.lo.u32 t1, yl, wh, t1;\n\tmad.lo.u32 t1, yh, wl, t1;\n\tmad.lo.u32 t1, ql, nh, t1;\n\tmad.lo.u32 t1, qh, nl, t1;\n\t"
        "mov.b64 %0, {c, t1};\n\t}"
        : "=l"(r) : "l"(a), "l"(w), "l"(ws), "l"(nq));
    return r;
}
// variant: Shoup companion scaled to 2^63, the three high partial products chained through multiply-add addends (no carry adds);
// the quotient is in units of 2q (nq = -2q), result in [0, 8q)
__device__ __forceinline__ u64 shoup63_lazy8(u64 a, u64 w, u64 ws63, u64 n2q) {
    u64 r;
    asm("{\n\t.reg .u32 yl, yh, wl, wh, sl, sh, nl, nh, t1, ml, mh, ql, qh, c, zr;\n\t.reg .u64 mid, h, p, z;\n\t"
        "mov.b64 {yl, yh}, %1; mov.b64 {wl, wh}, %2; mov.b64 {sl, sh}, %3; mov.b64 {nl, nh}, %4;\n\t"
        "mul.wide.u32 mid, yl, sh;\n\tmad.wide.u32 mid, yh, sl, mid;\n\t"
        "mov.b64 {ml, mh}, mid; mov.u32 zr, 0; mov.b64 z, {mh, zr};\n\t"
        "mad.wide.u32 h, yh, sh, z;\n\tmov.b64 {ql, qh}, h;\n\t"
        "mul.wide.u32 p, yl, wl;\n\tmad.wide.u32 p, ql, nl, p;\n\tmov.b64 {c, t1}, p;\n\t"
        "mad.lo.u32 t1, yl, wh, t1;\n\tmad.lo.u32 t1, yh, wl, t1;\n\tmad.lo.u32 t1, ql, nh, t1;\n\tmad.lo.u32 t1, qh, nl, t1;\n\t"
        "mov.b64 %0, {c, t1};\n\t}"
        : "=l"(r) : "l"(a), "l"(w), "l"(ws63), "l"(n2q));
    return r;
}
__device__ __forceinline__ void bf_int63(u64& x, u64& y, u64 w, u64 ws, u64 n2q, u64 q8) {
    const u64 v = shoup63_lazy8(y, w, ws, n2q); const u64 u = x; x = u + v; y = u - v + q8;
}
__device__ __forceinline__ void bf_int(u64& x, u64& y, u64 w, u64 ws, u64 nq, u64 q4) {
    const u64 v = shoup_lazy4(y, w, ws, nq); const u64 u = x; x = u + v; y = u - v + q4;
}

// ---- FP64 butterfly ----
struct FMod { double p; int thr_hi, p_hi, p_lo; double pinv; };
__device__ __forceinline__ double mulmod_f(double y, double w, double winv, double p) {
    const double h = __dmul_rn(y, w);
    const double l = __fma_rn(y, w, -h);
    const double q = __dsub_rn(__fma_rn(y, winv, MAGIC), MAGIC);
    return __dadd_rn(__fma_rn(-q, p, h), l);
}
// |x| <= 1.5 p -> |x| <= p/2 (+ 2^-20 p): threshold test on the high word, the addend built with integer selects
__device__ __forceinline__ double corr_alu(double x, const FMod& m) {
    const int hi = __double2hiint(x);
    const bool c = (hi & 0x7fffffff) > m.thr_hi;
    const int ahi = c ? (m.p_hi | (hi & 0x80000000)) : 0, alo = c ? m.p_lo : 0;
    return __dsub_rn(x, __hiloint2double(ahi, alo));
}
__device__ __forceinline__ double corr_f64(double x, const FMod& m) {
    const double k = __dsub_rn(__fma_rn(x, m.pinv, MAGIC), MAGIC);
    return __fma_rn(-k, m.p, x);
}
template <int CORR> __device__ __forceinline__ void bf_f64(double& x, double& y, double w, double winv, const FMod& m) {
    const double v = mulmod_f(y, w, winv, m.p);
    const double a = __dadd_rn(x, v), b = __dsub_rn(x, v);
    if (CORR == 0) { x = corr_alu(a, m); y = corr_alu(b, m); }
    else if (CORR == 1) { x = corr_f64(a, m); y = corr_f64(b, m); }
    else { x = a; y = b; }   // no correction: arithmetic floor only (not a valid transform)
}

__device__ __forceinline__ void radix8_int(u64* e, const u64* w, const u64* ws, u64 nq, u64 q4) {
#pragma unroll
    for (int s = 0; s < 3; ++s) { const int half = 4 >> s;
#pragma unroll
        for (int g = 0; g < (1 << s); ++g)
#pragma unroll
            for (int j = 0; j < half; ++j) bf_int(e[g * 2 * half + j], e[g * 2 * half + half + j], w[(1 << s) - 1 + g], ws[(1 << s) - 1 + g], nq, q4);
    }
}
__device__ __forceinline__ void radix8_int63(u64* e, const u64* w, const u64* ws, u64 n2q, u64 q8) {
#pragma unroll
    for (int s = 0; s < 3; ++s) { const int half = 4 >> s;
#pragma unroll
        for (int g = 0; g < (1 << s); ++g)
#pragma unroll
            for (int j = 0; j < half; ++j) bf_int63(e[g * 2 * half + j], e[g * 2 * half + half + j], w[(1 << s) - 1 + g], ws[(1 << s) - 1 + g], n2q, q8);
    }
}
template <int CORR> __device__ __forceinline__ void radix8_f64(double* e, const double* w, const double* wi, const FMod& m) {
#pragma unroll
    for (int s = 0; s < 3; ++s) { const int half = 4 >> s;
#pragma unroll
        for (int g = 0; g < (1 << s); ++g)
#pragma unroll
            for (int j = 0; j < half; ++j) bf_f64<CORR>(e[g * 2 * half + j], e[g * 2 * half + half + j], w[(1 << s) - 1 + g], wi[(1 << s) - 1 + g], m);
    }
}

// MODE 0: all warps integer; 1: all warps FP64 (CORR); 2: warps with (wid>>2)&1 integer, the others FP64
template <int MODE, int CORR> __global__ void k_bf(u64* out, long long* cyc, u64 w0, u64 ws0, u64 q, double pd, int it_int, int it_f64) {
    const int wid = threadIdx.x >> 5;
    const bool do_int = MODE == 0 || MODE == 3 || (MODE == 2 && ((wid >> 2) & 1));
    u64 acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (do_int) {
        u64 e[8], w[7], ws[7];
        for (int i = 0; i < 8; ++i) e[i] = threadIdx.x * 8 + i + w0;
        for (int i = 0; i < 7; ++i) { w[i] = w0 + i * 977; ws[i] = ws0 + i * 131; }
        const u64 nq = 0 - q, q4 = q << 2;
        if (MODE == 3) for (int it = 0; it < it_int; ++it) radix8_int63(e, w, ws, nq << 1, q << 3);
        else for (int it = 0; it < it_int; ++it) radix8_int(e, w, ws, nq, q4);
        for (int i = 0; i < 8; ++i) acc += e[i];
    } else {
        double e[8], w[7], wi[7];
        FMod m; m.p = pd; m.pinv = 1.0 / pd; m.thr_hi = __double2hiint(pd * 0.5); m.p_hi = __double2hiint(pd); m.p_lo = __double2loint(pd);
        for (int i = 0; i < 8; ++i) e[i] = (double)(threadIdx.x * 8 + i) + (double)(w0 & 0xffff);
        for (int i = 0; i < 7; ++i) { w[i] = (double)((w0 + i * 977) & 0x7ffffffffffffull) - 1125899906842624.0; wi[i] = w[i] / pd; }
        for (int it = 0; it < it_f64; ++it) radix8_f64<CORR>(e, w, wi, m);
        for (int i = 0; i < 8; ++i) acc += (u64)(long long)e[i];
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;   // warp 0; the slowest warp is bounded by the event time printed beside it
}

// ---- pipe probes: 8 independent chains per thread ----
template <int OP> __global__ void k_pipe(u64* out, long long* cyc, u64 c0, u64 c1, double d0, double d1, int iters) {
    u64 x[8]; double f[8]; u32 a[8];
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 8 + i + c0; f[i] = (double)(threadIdx.x + i) * d0; a[i] = threadIdx.x + i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) f[i] = __fma_rn(f[i], d0, d1);
            if (OP == 1) f[i] = __dadd_rn(f[i], d1);
            if (OP == 2) f[i] = __dmul_rn(f[i], d0);
            if (OP == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"((u32)c0), "r"((u32)c1));
            if (OP == 4) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(a[i]), "r"((u32)c1));
            if (OP == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"((u32)c1));
            if (OP == 6) { f[i] = __fma_rn(f[i], d0, d1); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"((u32)c0), "r"((u32)c1)); }
            if (OP == 7) { f[i] = __fma_rn(f[i], d0, d1); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"((u32)c1)); }
            if (OP == 8) { f[i] = __fma_rn(f[i], d0, d1); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"((u32)c0), "r"((u32)c1));
                           asm volatile("add.u32 %0, %0, %1;" : "+r"((u32&)x[i]) : "r"((u32)c1)); }
            if (OP == 9) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"((u32)c0), "r"((u32)c1));
                           asm volatile("add.u32 %0, %0, %1;" : "+r"((u32&)x[i]) : "r"((u32)c1)); }
            if (OP == 10) f[i] = (double)(long long)x[i] + f[i];            // I2F.F64.S64 + DADD
            if (OP == 11) { f[i] = __fma_rn(f[i], d0, d1); f[i] = __dadd_rn(f[i], d1);
                            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"((u32)c0), "r"((u32)c1));
                            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(a[i]), "r"((u32)c1)); }   // 2 FP64 + 2 IMAD
        }
    }
    const long long t1 = clock64();
    u64 s = 0; for (int i = 0; i < 8; ++i) s += x[i] + (u64)(long long)f[i] + a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static u64* g_out; static long long* g_cyc;
static double mean_cycles(int blocks) {
    static long long h[4096]; cudaMemcpy(h, g_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < blocks; ++i) s += (double)h[i]; return s / blocks;
}
template <int OP> void pipe(const char* name, int per_iter_ops) {
    const int threads = 512, bps = 2, iters = 4000, blocks = 148 * bps;
    k_pipe<OP><<<blocks, threads>>>(g_out, g_cyc, 0xFFFFFFF000000123ull, 0x000FFFFF12345677ull, 1.0000001, 0.5, 10);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k_pipe<OP><<<blocks, threads>>>(g_out, g_cyc, 0xFFFFFFF000000123ull, 0x000FFFFF12345677ull, 1.0000001, 0.5, iters); cudaEventRecord(b);
    cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b);
    const double cyc = mean_cycles(blocks), warp_ops = (double)bps * threads / 32 / 4 * iters * 8 * per_iter_ops;
    printf("%-40s %7.3f cycles per warp-instruction per SMSP   (clock64 %9.0f cycles, %.3f ms => %.0f MHz)\n", name, cyc / warp_ops, cyc, ms, cyc / ms / 1e3);
}
template <int MODE, int CORR> void bf(const char* name, int threads, int bps, int it_int, int it_f64) {
    const int blocks = 148 * bps;
    const u64 q = 0x10000002520001ull;
    k_bf<MODE, CORR><<<blocks, threads>>>(g_out, g_cyc, 12345, 678, q, (double)q, 10, 10);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k_bf<MODE, CORR><<<blocks, threads>>>(g_out, g_cyc, 12345, 678, q, (double)q, it_int, it_f64); cudaEventRecord(b);
    cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b);
    const double cyc = mean_cycles(blocks);
    const double warps = (double)bps * threads / 32 / 4;
    double nbf;   // warp-butterflies per sub-partition
    if (MODE == 0 || MODE == 3) nbf = warps * it_int * 12.0; else if (MODE == 1) nbf = warps * it_f64 * 12.0; else nbf = warps / 2 * (it_int + it_f64) * 12.0;
    const double mhz = 1965.0;   // event time -> cycles needs a clock; printed beside the in-kernel count
    printf("%-34s %4d thr x %d CTA/SM  it_int %5d it_f64 %5d : %6.2f cycles per warp-butterfly per SMSP (warp 0 clock64), %6.2f by event time at %.0f MHz, %.3f ms\n",
           name, threads, bps, it_int, it_f64, cyc / nbf, ms * 1e-3 * mhz * 1e6 / nbf, mhz, ms);
}
int main() {
    cudaMalloc(&g_out, (size_t)148 * 2 * 1024 * 8); cudaMalloc(&g_cyc, 4096 * sizeof(long long));
    pipe<0>("DFMA", 1); pipe<1>("DADD", 1); pipe<2>("DMUL", 1); pipe<3>("IMAD (mad.lo.u32)", 1); pipe<4>("IMAD.WIDE.U32 with 64-bit addend", 1);
    pipe<5>("IADD (add.u32)", 1); pipe<6>("DFMA + IMAD", 2); pipe<7>("DFMA + IADD", 2); pipe<8>("DFMA + IMAD + IADD", 3); pipe<9>("IMAD + IADD", 2);
    pipe<10>("I2F.F64.S64 + DADD", 2); pipe<11>("2 FP64 + IMAD + IMAD.WIDE", 4);
    bf<0, 0>("integer Shoup (PTX)", 512, 1, 2000, 0); bf<0, 0>("integer Shoup (PTX)", 512, 2, 2000, 0); bf<0, 0>("integer Shoup (PTX)", 1024, 2, 2000, 0);
    bf<3, 0>("integer Shoup, 2^63 companion", 512, 2, 2000, 0); bf<3, 0>("integer Shoup, 2^63 companion", 1024, 2, 2000, 0); bf<0, 0>("integer Shoup (PTX)", 1024, 2, 2000, 0);
    bf<3, 0>("integer Shoup, 2^63 companion", 1024, 2, 4000, 0); bf<0, 0>("integer Shoup (PTX)", 1024, 2, 4000, 0);
    bf<1, 0>("FP64, ALU-assisted correction", 512, 1, 0, 2000); bf<1, 0>("FP64, ALU-assisted correction", 512, 2, 0, 2000); bf<1, 0>("FP64, ALU-assisted correction", 1024, 2, 0, 2000);
    bf<1, 1>("FP64, FP64-only correction", 512, 2, 0, 2000);
    bf<1, 2>("FP64, no correction (floor)", 512, 2, 0, 2000);
    bf<2, 0>("mixed int | FP64(ALU corr)", 512, 2, 2000, 2000); bf<2, 0>("mixed int | FP64(ALU corr)", 512, 2, 1000, 2000); bf<2, 0>("mixed int | FP64(ALU corr)", 512, 2, 1000, 3000);
    bf<2, 1>("mixed int | FP64(FP64 corr)", 512, 2, 2000, 2000); bf<2, 1>("mixed int | FP64(FP64 corr)", 512, 2, 1000, 2000); bf<2, 1>("mixed int | FP64(FP64 corr)", 1024, 2, 1000, 2000);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
