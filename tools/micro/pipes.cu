// Micro-benchmarks of the sm_100a pipes the SDDE kernel leans on: dependent-issue latency and throughput of
// DFMA, IMAD.WIDE.U32, LOP3, MUFU.RSQ64H / RCP64H, and random-index LDS.128 (the log table).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

template <int NCH> __global__ void k_dfma(double *o, int iters, double a, double b) {
    double x[NCH];
    for (int i = 0; i < NCH; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < NCH; ++i) x[i] = fma(x[i], a, b);
    double s = 0; for (int i = 0; i < NCH; ++i) s += x[i];
    o[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NCH> __global__ void k_imadwide(u64 *o, int iters, u32 m) {
    u64 x[NCH];
    for (int i = 0; i < NCH; ++i) x[i] = threadIdx.x + i + 1;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < NCH; ++i) x[i] = (u64)(u32)x[i] * m + (x[i] >> 32);   // IMAD.WIDE.U32 with 64-bit addend
    u64 s = 0; for (int i = 0; i < NCH; ++i) s += x[i];
    o[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NCH> __global__ void k_lop3(u32 *o, int iters, u32 m, u32 k) {
    u32 x[NCH];
    for (int i = 0; i < NCH; ++i) x[i] = threadIdx.x + i + 1;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < NCH; ++i) x[i] = x[i] ^ m ^ (x[(i + 1) % NCH] & k);
    u32 s = 0; for (int i = 0; i < NCH; ++i) s += x[i];
    o[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NCH> __global__ void k_rsq(double *o, int iters) {
    double x[NCH];
    for (int i = 0; i < NCH; ++i) x[i] = 1.5 + threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < NCH; ++i) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[i])); x[i] = y; }
    double s = 0; for (int i = 0; i < NCH; ++i) s += x[i];
    o[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NCH> __global__ void k_lds(double *o, int iters, u32 seed) {
    __shared__ double2 tab[128];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) tab[i] = make_double2(i, 2 * i);
    __syncthreads();
    u32 idx[NCH]; double acc = 0;
    for (int i = 0; i < NCH; ++i) idx[i] = (threadIdx.x * 2654435761u + i * 40503u + seed) >> 7;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < NCH; ++i) { double2 t = tab[idx[i] & 127]; acc += t.x; idx[i] = idx[i] * 1664525u + 1013904223u + (u32)t.y; }
    o[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best = ms < best ? ms : best; }
    return best;
}
int main() {
    void *buf; cudaMalloc(&buf, 148 * 32 * 1024 * 8);
    const int iters = 1024;
    const double clk = 1.965e9;
    printf("op            chains warps/SMSP  cycles per warp-instr per SMSP (throughput)   cycles per dependent instr (1 warp, 1 chain)\n");
#define RUN(NAME, KERN, NCH, WPS)                                                                                   \
    {                                                                                                                 \
        float ms = timeit([&] { KERN<NCH><<<148, 128 * WPS>>>((decltype(KERN##_arg))buf, iters ARGS); });            \
        double instr_per_smsp = (double)WPS * iters * 8 * NCH;                                                        \
        printf("%-12s  %4d   %4d      %.2f cycles/instr\n", NAME, NCH, WPS, ms * 1e-3 * clk / instr_per_smsp);          \
    }
    double *k_dfma_arg = nullptr; u64 *k_imadwide_arg = nullptr; u32 *k_lop3_arg = nullptr; double *k_rsq_arg = nullptr; double *k_lds_arg = nullptr;
#define ARGS , 0.999, 1e-3
    RUN("DFMA", k_dfma, 1, 1) RUN("DFMA", k_dfma, 2, 1) RUN("DFMA", k_dfma, 4, 1) RUN("DFMA", k_dfma, 8, 1) RUN("DFMA", k_dfma, 2, 5) RUN("DFMA", k_dfma, 4, 5) RUN("DFMA", k_dfma, 8, 8)
#undef ARGS
#define ARGS , 0xD2511F53u
    RUN("IMAD.WIDE", k_imadwide, 1, 1) RUN("IMAD.WIDE", k_imadwide, 4, 1) RUN("IMAD.WIDE", k_imadwide, 8, 1) RUN("IMAD.WIDE", k_imadwide, 4, 5) RUN("IMAD.WIDE", k_imadwide, 8, 8)
#undef ARGS
#define ARGS , 0x9E3779B9u, 0xff00ffu
    RUN("LOP3x2", k_lop3, 1, 1) RUN("LOP3x2", k_lop3, 8, 1) RUN("LOP3x2", k_lop3, 8, 8)
#undef ARGS
#define ARGS
    RUN("MUFU.RSQ64H", k_rsq, 1, 1) RUN("MUFU.RSQ64H", k_rsq, 4, 1) RUN("MUFU.RSQ64H", k_rsq, 8, 8)
#undef ARGS
#define ARGS , 12345u
    RUN("LDS.128rnd", k_lds, 1, 1) RUN("LDS.128rnd", k_lds, 4, 1) RUN("LDS.128rnd", k_lds, 4, 5) RUN("LDS.128rnd", k_lds, 8, 8)
#undef ARGS
    return 0;
}
