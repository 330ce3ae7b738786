// Micro-benchmark: do FP64 (DFMA) and INT (IMAD/LOP3) instructions share issue bandwidth on sm_100a?
// Each thread runs NF independent DFMA chains and NI independent integer chains; we time three mixes.
#include <cstdio>
#include <cuda_runtime.h>
template <int NF, int NI>
__global__ void __launch_bounds__(256) mix(double *out, unsigned *outi, int iters, double a, double b, unsigned m) {
    double x[NF > 0 ? NF : 1];
    unsigned y[NI > 0 ? NI : 1];
    for (int i = 0; i < NF; ++i) x[i] = threadIdx.x + i;
    for (int i = 0; i < NI; ++i) y[i] = threadIdx.x * 7 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < NF; ++i) x[i] = fma(x[i], a, b);
#pragma unroll
            for (int i = 0; i < NI; ++i) y[i] = (y[i] * m) ^ (y[i] >> 3);   // IMAD + SHF + LOP3 (3 int instr)
        }
    }
    double s = 0; unsigned t = 0;
    for (int i = 0; i < NF; ++i) s += x[i];
    for (int i = 0; i < NI; ++i) t ^= y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s; outi[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int NF, int NI> void run(const char *name) {
    double *o; unsigned *oi; cudaMalloc(&o, 148 * 8 * 256 * 8); cudaMalloc(&oi, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2048;
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        mix<NF, NI><<<148 * 8, 256>>>(o, oi, iters, 0.999, 1e-3, 2654435761u);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best = ms < best ? ms : best;
    }
    // per SMSP: warps = 8 blocks*8 warps/4 = 16 warps; instr per warp = iters*4*(NF + 3*NI)
    double warp_instr_per_smsp = 16.0 * iters * 4 * (NF + 3.0 * NI);
    double cycles = best * 1e-3 * 1.965e9;
    printf("%-18s NF=%d NI=%d  %.3f ms  cycles/SMSP %.3e  warp-instr/SMSP %.3e  IPC/SMSP %.3f  (fp64 instr/cycle %.3f, int instr/cycle %.3f)\n",
           name, NF, NI, best, cycles, warp_instr_per_smsp, warp_instr_per_smsp / cycles, 16.0 * iters * 4 * NF / cycles,
           16.0 * iters * 4 * 3.0 * NI / cycles);
    cudaFree(o); cudaFree(oi);
}
int main() {
    run<8, 0>("fp64 only");
    run<0, 8>("int only");
    run<8, 2>("fp64 + 6 int");
    run<8, 3>("fp64 + 9 int (1:1.1)");
    run<8, 6>("fp64 + 18 int");
    run<4, 8>("4 fp64 + 24 int");
    return 0;
}
