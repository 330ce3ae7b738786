// Micro-benchmark of the in-kernel normal generator parts (device functions of csrc/sdde_kernels.cuh):
// Philox alone, Box-Muller transform alone, both back to back, and the hand-interleaved fused version.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../stochasticdelaydiffeq.jl_b200/csrc/sdde_kernels.cuh"
using namespace sdde;
template <int MODE> __global__ void __launch_bounds__(128) k(double *o, int iters, u32 k0, u32 k1) {
    __shared__ double2 ltab[SDDE_RNG_TAB_ENTRIES];
    for (int i = threadIdx.x; i < SDDE_RNG_TAB_ENTRIES; i += blockDim.x) ltab[i] = RNG_TAB[i];
    __syncthreads();
    u32 rk[20]; philox_round_keys(k0, k1, rk);
    const u64 gid = blockIdx.x * blockDim.x + threadIdx.x;
    u32 w[4] = {(u32)gid * 2654435761u, 12345u + (u32)gid, 777u, (u32)gid * 40503u};
    u32 w2[4] = {w[1], w[0], w[3], w[2]};
    double acc = 0; u32 acci = 0;
    for (int it = 0; it < iters; ++it) {
        double z0, z1, y0, y1;
        if (MODE == 0) {   // Philox only (2 independent streams like TPT=2)
            philox4x32_10((u32)gid, w[0], (u32)it, 0, rk, w); philox4x32_10((u32)gid + 1, w2[0], (u32)it, 0, rk, w2);
            acci += w[0] ^ w[1] ^ w[2] ^ w[3] ^ w2[0] ^ w2[3];
        } else if (MODE == 1) {   // transform only
            w[0] += 0x9E3779B9u; w[1] ^= w[0] >> 3; w[2] += 0x7F4A7C15u; w[3] ^= w[2] << 5;
            w2[0] += 0x9E3779B9u; w2[1] ^= w2[0] >> 3; w2[2] += 0x7F4A7C15u; w2[3] ^= w2[2] << 5;
            normal_pair_from_words(w, ltab, z0, z1); normal_pair_from_words(w2, ltab, y0, y1);
            acc += z0 + z1 + y0 + y1;
        } else if (MODE == 2) {   // back to back
            philox4x32_10((u32)gid, w[0], (u32)it, 0, rk, w); philox4x32_10((u32)gid + 1, w2[0], (u32)it, 0, rk, w2);
            normal_pair_from_words(w, ltab, z0, z1); normal_pair_from_words(w2, ltab, y0, y1);
            acc += z0 + z1 + y0 + y1;
        } else {   // fused
            normal_pair_fused(w, gid, (u64)it, rk, ltab, z0, z1); normal_pair_fused(w2, gid + 1, (u64)it, rk, ltab, y0, y1);
            acc += z0 + z1 + y0 + y1;
        }
    }
    o[gid] = acc + acci;
}
template <int MODE> void run(const char *name, int blocks_per_sm) {
    double *o; cudaMalloc(&o, 148 * 32 * 128 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096; float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) { cudaEventRecord(e0); k<MODE><<<148 * blocks_per_sm, 128>>>(o, iters, 1, 2); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best = ms < best ? ms : best; }
    // per SMSP: blocks_per_sm warps (128 threads = 4 warps = 1 per SMSP per block); each iteration = 2 pair-generations per thread
    double cyc = best * 1e-3 * 1.965e9 / ((double)blocks_per_sm * iters);
    printf("%-22s warps/SMSP %d   %.1f cycles per (warp x 2 pairs) per SMSP\n", name, blocks_per_sm, cyc);
    cudaFree(o);
}
int main() {
    for (int b : {1, 5, 8}) {
        if (b == 1) { run<0>("philox x2", 1); run<1>("box-muller x2", 1); run<2>("philox+bm x2", 1); run<3>("fused x2", 1); }
        if (b == 5) { run<0>("philox x2", 5); run<1>("box-muller x2", 5); run<2>("philox+bm x2", 5); run<3>("fused x2", 5); }
        if (b == 8) { run<0>("philox x2", 8); run<1>("box-muller x2", 8); run<2>("philox+bm x2", 8); run<3>("fused x2", 8); }
    }
    return 0;
}
