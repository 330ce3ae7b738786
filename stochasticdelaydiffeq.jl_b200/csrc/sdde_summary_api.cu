// sdde_summary_api.cu -- sdde_b200_quantiles: exact order statistics across trajectories (kernels: sdde_summary.cuh,
// host plan/combine: summary.cpp).  Device side of [EXT] SciMLBase EnsembleSummary.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "sdde_internal.cuh"
#include "summary.hpp"
#include "sdde_summary.cuh"
#define fail sdde_fail

extern "C" {

int sdde_b200_quantiles(sdde_ctx *c, const double *paths, int32_t mem, int64_t n_rows, int64_t n_traj, int64_t pitch,
                        const double *probs, int32_t n_prob, int32_t want_median, double *result, double *kernel_ms) {
    if (!c || !paths || !result || n_rows <= 0 || n_traj <= 0 || pitch < n_traj || n_prob < 0 || (n_prob > 0 && !probs) ||
        (n_prob == 0 && !want_median))
        return fail(SDDE_ERR_BAD_DESC, "bad arguments");
    if (mem != SDDE_MEM_HOST && mem != SDDE_MEM_DEVICE) return fail(SDDE_ERR_BAD_DESC, "mem must be SDDE_MEM_HOST or SDDE_MEM_DEVICE");
    if (n_traj >= ((int64_t)1 << 32)) return fail(SDDE_ERR_UNSUPPORTED, "rows longer than 2^32 - 1 values");
    sdde::QuantilePlan plan;
    const std::string perr = sdde::plan_quantiles(n_traj, probs, n_prob, want_median != 0, plan);
    if (!perr.empty()) return fail(SDDE_ERR_BAD_DESC, "%s", perr.c_str());
    const int R = (int)plan.ranks.size();
    if (R > sdde::QS_MAX_RANK) return fail(SDDE_ERR_UNSUPPORTED, "at most %d distinct order statistics per call", sdde::QS_MAX_RANK);
    std::lock_guard<std::mutex> lk(c->mu);
    CU_CHECK(cudaSetDevice(c->device));
    const double *src = paths;
    if (mem == SDDE_MEM_HOST) {
        CU_CHECK(c->qs_paths.reserve((size_t)n_rows * pitch * 8));
        CU_CHECK(cudaMemcpyAsync(c->qs_paths.p, paths, (size_t)n_rows * pitch * 8, cudaMemcpyHostToDevice, c->compute));
        src = (const double *)c->qs_paths.p;
    }
    const size_t hist_bytes = (size_t)n_rows * R * sdde::QS_BINS * sizeof(unsigned int);
    CU_CHECK(c->qs_rows.reserve((size_t)n_rows * sizeof(sdde::QsRow)));
    CU_CHECK(c->qs_hist.reserve(hist_bytes));
    sdde::QsRanks ranks{};
    for (int r = 0; r < R; ++r) ranks.k[r] = plan.ranks[(size_t)r];
    const size_t smem = (size_t)R * sdde::QS_BINS * sizeof(unsigned int);
    CU_CHECK(c->qs_cand.reserve((size_t)n_rows * R * sdde::QS_CAP * 8));
    for (int pass = 0; pass < sdde::QS_PASSES; ++pass)
        CU_CHECK(cudaFuncSetAttribute(sdde::qs_stream_fn(pass), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // enough CTAs to fill the machine, but no slice shorter than 8192 values
    int64_t split = (8 * (int64_t)c->sm_count + n_rows - 1) / n_rows;
    split = std::max<int64_t>(1, std::min<int64_t>(split, (n_traj + 8191) / 8192));
    CU_CHECK(cudaEventRecord(c->ev_k0, c->compute));
    sdde::qs_init_kernel<<<(unsigned)n_rows, 32, 0, c->compute>>>((sdde::QsRow *)c->qs_rows.p, R, ranks);
    for (int pass = 0; pass < sdde::QS_PASSES; ++pass) {
        CU_CHECK(cudaMemsetAsync(c->qs_hist.p, 0, hist_bytes, c->compute));
        sdde::qs_stream_fn(pass)<<<dim3((unsigned)n_rows, (unsigned)split), sdde::QS_THREADS, smem, c->compute>>>(
            src, pitch, n_traj, R, (sdde::QsRow *)c->qs_rows.p, (unsigned int *)c->qs_hist.p, (unsigned long long *)c->qs_cand.p);
        sdde::qs_scan_kernel<<<(unsigned)n_rows, 32 * R, 0, c->compute>>>((sdde::QsRow *)c->qs_rows.p, (const unsigned int *)c->qs_hist.p,
                                                                          (const unsigned long long *)c->qs_cand.p, R, pass);
    }
    CU_CHECK(cudaGetLastError());
    CU_CHECK(cudaEventRecord(c->ev_k1, c->compute));
    std::vector<sdde::QsRow> rows((size_t)n_rows);
    CU_CHECK(cudaMemcpyAsync(rows.data(), c->qs_rows.p, (size_t)n_rows * sizeof(sdde::QsRow), cudaMemcpyDeviceToHost, c->compute));
    CU_CHECK(cudaStreamSynchronize(c->compute));
    if (kernel_ms) { float ms = 0.f; CU_CHECK(cudaEventElapsedTime(&ms, c->ev_k0, c->ev_k1)); *kernel_ms = ms; }
    const int ncol = (int)plan.items.size();
    for (int64_t i = 0; i < n_rows; ++i) {
        const sdde::QsRow &q = rows[(size_t)i];
        double v[sdde::QS_MAX_RANK];
        for (int r = 0; r < R; ++r) {
            const unsigned long long key = q.prefix[r];
            const unsigned long long b = (key >> 63) ? (key & 0x7fffffffffffffffull) : ~key;
            std::memcpy(&v[r], &b, 8);
        }
        for (int j = 0; j < ncol; ++j) {
            const sdde::QuantileItem &it = plan.items[(size_t)j];
            result[i * ncol + j] = q.n_nan ? std::nan("") : sdde::combine_quantile(it, v[it.ia], v[it.ib]);
        }
    }
    return SDDE_OK;
}

}  // extern "C"
