// Does loading many distinct FP64 constants (LDCU -> uniform register -> DFMA operand) limit a kernel on sm_100a?
// 64 DFMA per iteration, each with its own coefficient: (a) coefficients from the kernel-parameter constant bank
// (one LDCU.64 per DFMA, cannot be hoisted: 64 x 2 uniform registers > 63), (b) the same 64 values held in 8 registers
// reused (no constant traffic).
#include <cstdio>
#include <cuda_runtime.h>
struct P { double c[64]; };
template <int MODE> __global__ void __launch_bounds__(128) k(double *o, int iters, const __grid_constant__ P p) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    double r[8];
    for (int i = 0; i < 8; ++i) r[i] = p.c[i];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
            if (MODE == 0) { x0 = fma(x0, p.c[j], 1e-3); x1 = fma(x1, p.c[j + 1], 1e-3); x2 = fma(x2, p.c[j + 2], 1e-3); x3 = fma(x3, p.c[j + 3], 1e-3); }
            else { x0 = fma(x0, r[j & 7], 1e-3); x1 = fma(x1, r[(j + 1) & 7], 1e-3); x2 = fma(x2, r[(j + 2) & 7], 1e-3); x3 = fma(x3, r[(j + 3) & 7], 1e-3); }
        }
    }
    o[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}
template <int MODE> void run(const char *name, int bps) {
    double *o; cudaMalloc(&o, 148 * 16 * 128 * 8);
    P p; for (int i = 0; i < 64; ++i) p.c[i] = 0.99 + 1e-4 * i;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2048; float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) { cudaEventRecord(e0); k<MODE><<<148 * bps, 128>>>(o, iters, p); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best = ms < best ? ms : best; }
    printf("%-28s warps/SMSP %d  %.2f cycles per DFMA per SMSP\n", name, bps, best * 1e-3 * 1.965e9 / ((double)bps * iters * 64));
    cudaFree(o);
}
int main() { for (int b : {3, 6}) { run<0>("64 distinct constants (LDCU)", b); run<1>("8 registers", b); } return 0; }
