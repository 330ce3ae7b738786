// sdde_internal.cuh -- shared by the translation units of libsdde_b200.so: error reporting, device/pinned buffers,
// the context object behind the opaque `sdde_ctx` of include/sdde_b200.h, and the few helpers that cross files.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sdde_b200.h"
#include "schedule.hpp"
#include "user_model.hpp"

// thread-local message of the last failing call (sdde_b200_last_error); returns `code`
int sdde_fail(int code, const char *fmt, ...);
const std::string &sdde_last_error_string();

#define CU_CHECK(expr)                                                                                        \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess)                                                                                \
            return sdde_fail(SDDE_ERR_CUDA - (int)_e, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                                             \
    } while (0)

#include <memory>

#include "host_pool.hpp"

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinnedBuf {       // page-locked host staging (small D2H results: status bytes, moment triples)
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct sdde_qsel;
void sdde_qsel_free(struct sdde_qsel *q);

struct sdde_ctx {
    int device = 0;
    cudaStream_t compute = nullptr, copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_c0 = nullptr, ev_c1 = nullptr;
    cudaEvent_t ev_kdone[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr}, ev_in = nullptr;
    DevBuf ring, out[2], steps, saves, noise, u0, ptraj, status, mom;
    DevBuf rng_tab;                  // the normal generator's table (rng_table.cpp), uploaded once at creation
    DevBuf mom_part;                 // per-slice moment triples of one wave (moments_kernel -> merge_moments_kernel)
    DevBuf noise_src, noise_nodes;   // supplied NoiseGrid on its own grid: staged W rows and the node table (plan_noise_grid)
    PinnedBuf h_status, h_mom, h_count;
    PinnedBuf bounce[2];           // pageable host output: page-locked landing buffers of the D2H chunks (2 x 64 MB, on first use)
    cudaEvent_t ev_bounce[2] = {nullptr, nullptr};
    std::unique_ptr<sdde::HostPool> pool;   // host threads that scatter the bounce buffers into the caller's pageable array
    int64_t opt_host_threads = 0;  // 0 = min(16, hardware threads)
    int64_t opt_bounce_min_mb = 32; // pageable outputs below this size take the plain cudaMemcpy2DAsync (0: always bounce -- tests)
    DevBuf qs_rows, qs_hist, qs_paths, qs_cand;   // order-statistics workspaces (sdde_b200_quantiles)
    struct sdde_qsel *qsel = nullptr;   // the active order-statistics selection (sdde_summary_api.cu), freed by sdde_qsel_free
    DevBuf gather;                 // [n_gpu][rows][3] moment triples received over NCCL (sdde_b200_solve_multi)
    bool keep_moments_on_device = false;   // set by solve_multi: leave the wave-merged triples in `mom` for the all-gather
    int64_t opt_jit_linsys = 1;    // specialise the linear delay system's constants with NVRTC (zeros vanish, no LDCU per entry)
    std::vector<std::pair<std::string, int>> linsys_jit;   // (shape + parameter bytes) -> user model id
    int64_t opt_nccl = 1;          // solve_multi: 1 = merge moments with an NCCL all-gather + device merge, 0 = on the host
    std::mutex mu;
    int64_t opt_wave_traj = 0;     // 0 = auto
    int64_t opt_batch = 0;         // 0 = compiled-in depth
    int64_t opt_chunk_steps = 128;
    int64_t opt_balance = 1;       // spread a wave's last, partial round of CTAs over all SMs (0: full CTAs only)
    int64_t opt_lean = 1;          // mark uniform batches for the lean body of run_batch (0: general body everywhere; A/B and tests)
    int64_t opt_zero_copy = -1;    // page-locked host `out`: 1 = the step kernel stores the saved paths straight into it over PCIe,
                                   // 0 = staged in HBM and copied by the copy engines wave by wave, -1 = auto: small outputs
                                   // zero-copy; large ones try staged, then zero-copy, and keep the faster (device-timed) --
                                   // which one wins depends on the host (SM posted writes reach 45-50 GB/s, copy engines
                                   // 52-57 GB/s on the measured boxes) and on how many GPUs share it
    struct ZcTrial { double ms_per_gb[2] = {0.0, 0.0}; int n[2] = {0, 0}; };   // [0] staged, [1] zero-copy: device time from the
    ZcTrial zc[64];                // first kernel to the last byte in host memory, per GB; one entry per power-of-two class of
                                   // output bytes (which arm wins is a property of the host and of the size, not of the model)
    int64_t opt_ramp_start = (int64_t)1 << 16;   // trajectories of the first wave when paths leave over PCIe (host output)
    int64_t opt_mem_limit = 0;     // bytes the workspaces may use (0 = 80 % of the free device memory)
    sdde::UserModels user;
    int sm_count = 148;
    // schedule cache: repeated solves of the same problem shape (ensemble sweeps, benchmarks) reuse the host schedule
    // and the packed device tables instead of replaying the time stepping and re-uploading them
    std::string sched_key, table_key;
    size_t mem_free_cached = 0;     // cudaMemGetInfo costs up to ~1.5 ms: refreshed only when the problem shape changes
    std::string mem_key;
    sdde::Schedule sched;
};

// descriptor validation (sizes, pointers, supported combinations); 0 or an SDDE_ERR_* code with the message set
int sdde_check_desc(const sdde_desc *d);
// trajectory-independent sizes of a solve: saved rows (n_save x saved states) and supplied-noise rows
int sdde_plan_sizes(sdde_ctx *ctx, const sdde_desc *d, int64_t *rows, int64_t *noise_rows);
// dst[i] = Chan merge of src[p * part_stride + 3 i], p = 0..parts-1, in order (device pointers), on `stream`
cudaError_t sdde_launch_merge_moments(cudaStream_t stream, const double *src, long long part_stride, int parts, long long n, double *dst);
