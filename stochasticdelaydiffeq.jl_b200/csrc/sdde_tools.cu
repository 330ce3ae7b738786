// sdde_tools.cu -- test hooks and micro-benchmarks of the C ABI: the normals the step kernels draw, measured DFMA and
// copy peaks of this GPU (roofline denominators), page-locked host memory helpers.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "rng_table.hpp"
#include "sdde_internal.cuh"
#include "sdde_kernels.cuh"
#define fail sdde_fail

namespace sdde {

// test hook: the normals the step kernels would draw, laid out [n][c][traj]
__global__ void philox_dump_kernel(double *out, long long n_traj, long long traj_off, long long n_steps, int nc, u32 k0,
                                   u32 k1, const double2 *__restrict__ rng_tab) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const double2 *ltab = stage_rng_table<SDDE_RNG_COPIES>(reinterpret_cast<double2 *>(smem_raw), rng_tab, (int)blockDim.x);
    __syncthreads();
    u32 rk[20];
    philox_round_keys(k0, k1, rk);
    const long long tr = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tr >= n_traj) return;
    const long long nq = n_steps * nc;
    for (long long call = 0; call * 4 < nq; ++call) {
        double z[4];
        philox_normal_quad<SDDE_RNG_COPIES>((u64)(traj_off + tr), (u64)call, rk, ltab, z);
        for (int i = 0; i < 4; ++i)
            if (4 * call + i < nq) out[(4 * call + i) * n_traj + tr] = z[i];
    }
}

// FP64 FMA throughput micro-benchmark: 8 independent chains per thread
__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

__global__ void __launch_bounds__(256) copy_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

}  // namespace sdde

extern "C" {

int sdde_b200_philox_normals(sdde_ctx *c, uint64_t seed, int64_t traj_offset, int64_t n_traj, int64_t n_steps, int32_t n_noise,
                             double *out_host) {
    if (!c || !out_host || n_traj <= 0 || n_steps <= 0 || n_noise <= 0) return fail(SDDE_ERR_BAD_DESC, "bad arguments");
    std::lock_guard<std::mutex> lk(c->mu);
    CU_CHECK(cudaSetDevice(c->device));
    const size_t bytes = (size_t)n_steps * n_noise * n_traj * 8;
    CU_CHECK(c->out[0].reserve(bytes));
    CU_CHECK(cudaFuncSetAttribute(sdde::philox_dump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SDDE_RNG_SMEM_BYTES(SDDE_RNG_COPIES)));
    sdde::philox_dump_kernel<<<(unsigned)((n_traj + 255) / 256), 256, SDDE_RNG_SMEM_BYTES(SDDE_RNG_COPIES), c->compute>>>(
        (double *)c->out[0].p, n_traj, traj_offset, n_steps, n_noise, (sdde::u32)(seed & 0xffffffffu), (sdde::u32)(seed >> 32),
        (const double2 *)c->rng_tab.p);
    CU_CHECK(cudaGetLastError());
    CU_CHECK(cudaMemcpyAsync(out_host, c->out[0].p, bytes, cudaMemcpyDeviceToHost, c->compute));
    CU_CHECK(cudaStreamSynchronize(c->compute));
    return SDDE_OK;
}

// test hook: the host copy of the generator's table (sdde_rng_tables.cuh), 2 * n_entries doubles; no device needed
int sdde_b200_rng_table(double *out, int64_t cap_doubles, int64_t *n_entries) {
    const int n = sdde::rng_table_entries();
    if (n_entries) *n_entries = n;
    if (out) {
        if (cap_doubles < 2 * (int64_t)n) return fail(SDDE_ERR_BAD_DESC, "buffer too small for %d table entries", n);
        memcpy(out, sdde::rng_table_host(), sizeof(double) * 2 * (size_t)n);
    }
    return SDDE_OK;
}

int sdde_b200_dfma_peak(sdde_ctx *c, double *tflops) {
    if (!c || !tflops) return fail(SDDE_ERR_BAD_DESC, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CU_CHECK(cudaSetDevice(c->device));
    const int blocks = c->sm_count * 8, threads = 256, iters = 4096;
    CU_CHECK(c->out[0].reserve((size_t)blocks * threads * 8));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CU_CHECK(cudaEventRecord(c->ev_k0, c->compute));
        sdde::dfma_kernel<<<blocks, threads, 0, c->compute>>>((double *)c->out[0].p, iters, 0.999999, 1e-9);
        CU_CHECK(cudaGetLastError());
        CU_CHECK(cudaEventRecord(c->ev_k1, c->compute));
        CU_CHECK(cudaStreamSynchronize(c->compute));
        float ms = 0.f;
        CU_CHECK(cudaEventElapsedTime(&ms, c->ev_k0, c->ev_k1));
        const double flops = 2.0 * 64.0 * (double)iters * blocks * threads;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    *tflops = best;
    return SDDE_OK;
}

int sdde_b200_hbm_peak(sdde_ctx *c, double *gbs) {
    if (!c || !gbs) return fail(SDDE_ERR_BAD_DESC, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CU_CHECK(cudaSetDevice(c->device));
    const size_t bytes = (size_t)1 << 30;
    CU_CHECK(c->out[0].reserve(bytes));
    CU_CHECK(c->out[1].reserve(bytes));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CU_CHECK(cudaEventRecord(c->ev_k0, c->compute));
        sdde::copy_kernel<<<c->sm_count * 16, 256, 0, c->compute>>>((const uint4 *)c->out[0].p, (uint4 *)c->out[1].p, (long long)(bytes / 16));
        CU_CHECK(cudaGetLastError());
        CU_CHECK(cudaEventRecord(c->ev_k1, c->compute));
        CU_CHECK(cudaStreamSynchronize(c->compute));
        float ms = 0.f;
        CU_CHECK(cudaEventElapsedTime(&ms, c->ev_k0, c->ev_k1));
        if (rep > 0) best = std::max(best, 2.0 * (double)bytes / (ms * 1e-3) / 1e9);
    }
    *gbs = best;
    return SDDE_OK;
}

void *sdde_b200_alloc_pinned(int64_t bytes) {
    void *p = nullptr;
    if (bytes <= 0) return nullptr;
    // SDDE_B200_PINNED_FLAGS (measurement hook, tools/e2e_ab.py): cudaHostAlloc flags, e.g. 4 = write-combined, 1 = portable
    const char *fl = getenv("SDDE_B200_PINNED_FLAGS");
    const unsigned flags = fl ? (unsigned)atoi(fl) : cudaHostAllocDefault;
    cudaError_t e = cudaHostAlloc(&p, (size_t)bytes, flags);
    if (e != cudaSuccess) { cudaGetLastError(); fail(SDDE_ERR_CUDA - (int)e, "cudaHostAlloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e)); return nullptr; }
    return p;
}

void sdde_b200_free_pinned(void *p) { if (p) cudaFreeHost(p); }

}  // extern "C"
