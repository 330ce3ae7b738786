// sdde_summary_api.cu -- exact order statistics across trajectories (kernels: sdde_summary.cuh, host plan/combine:
// summary.cpp).  Device side of [EXT] SciMLBase EnsembleSummary.
//   sdde_b200_quantiles          one GPU holds every trajectory of the rows: all passes inside one call
//   sdde_b200_qsel_*             the same selection in phases, for an ensemble sharded over several processes/GPUs: each
//                                rank histograms its own columns, the caller all-reduces the histograms (NCCL), every rank
//                                applies the same bin decisions -- the selected values are those of the WHOLE ensemble
#include <algorithm>
#include <cmath>
#include <cstring>

#include "sdde_internal.cuh"
#include "summary.hpp"
#include "sdde_summary.cuh"
#define fail sdde_fail

struct sdde_qsel {
    sdde::QuantilePlan plan;
    int R = 0;
    int64_t n_rows = 0, n_local = 0, pitch = 0, split = 1;
    const double *src = nullptr;
    unsigned int *hist = nullptr;          // [n_rows][R][QS_BINS], the caller's buffer or ctx->qs_hist
    size_t hist_bytes = 0, smem = 0;
    int allow_gather = 1;
    std::vector<int64_t> nan_global;       // set by sdde_b200_qsel_nan(..., set = 1)
};
void sdde_qsel_free(sdde_qsel *q) { delete q; }

namespace {

int qsel_begin(sdde_ctx *c, const double *paths, int32_t mem, int64_t n_rows, int64_t n_local, int64_t pitch, int64_t n_total,
               const double *probs, int32_t n_prob, int32_t want_median, unsigned int *d_hist, int allow_gather) {
    if (n_rows <= 0 || n_local < 0 || n_total <= 0 || n_local > n_total || pitch < n_local || n_prob < 0 || (n_prob > 0 && !probs) ||
        (n_prob == 0 && !want_median) || (n_local > 0 && !paths))
        return fail(SDDE_ERR_BAD_DESC, "bad arguments");
    if (mem != SDDE_MEM_HOST && mem != SDDE_MEM_DEVICE) return fail(SDDE_ERR_BAD_DESC, "mem must be SDDE_MEM_HOST or SDDE_MEM_DEVICE");
    if (n_total >= ((int64_t)1 << 31)) return fail(SDDE_ERR_UNSUPPORTED, "ensembles of 2^31 or more values per row");
    sdde_qsel_free(c->qsel); c->qsel = nullptr;
    sdde_qsel *q = new sdde_qsel();
    const std::string perr = sdde::plan_quantiles(n_total, probs, n_prob, want_median != 0, q->plan);
    if (!perr.empty()) { delete q; return fail(SDDE_ERR_BAD_DESC, "%s", perr.c_str()); }
    q->R = (int)q->plan.ranks.size();
    if (q->R > sdde::QS_MAX_RANK) { delete q; return fail(SDDE_ERR_UNSUPPORTED, "at most %d distinct order statistics per call", sdde::QS_MAX_RANK); }
    c->qsel = q;
    q->n_rows = n_rows; q->n_local = n_local; q->pitch = pitch; q->allow_gather = allow_gather;
    q->src = paths;
    if (mem == SDDE_MEM_HOST && n_local > 0) {
        CU_CHECK(c->qs_paths.reserve((size_t)n_rows * pitch * 8));
        CU_CHECK(cudaMemcpyAsync(c->qs_paths.p, paths, (size_t)n_rows * pitch * 8, cudaMemcpyHostToDevice, c->compute));
        q->src = (const double *)c->qs_paths.p;
    }
    q->hist_bytes = (size_t)n_rows * q->R * sdde::QS_BINS * sizeof(unsigned int);
    CU_CHECK(c->qs_rows.reserve((size_t)n_rows * sizeof(sdde::QsRow)));
    if (d_hist) q->hist = d_hist;
    else { CU_CHECK(c->qs_hist.reserve(q->hist_bytes)); q->hist = (unsigned int *)c->qs_hist.p; }
    CU_CHECK(c->qs_cand.reserve((size_t)n_rows * q->R * sdde::QS_CAP * 8));
    q->smem = (size_t)q->R * sdde::QS_BINS * sizeof(unsigned int);
    for (int pass = 0; pass < sdde::QS_PASSES; ++pass)
        CU_CHECK(cudaFuncSetAttribute(sdde::qs_stream_fn(pass), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q->smem));
    // enough CTAs to fill the machine, but no slice shorter than 8192 values
    int64_t split = (8 * (int64_t)c->sm_count + n_rows - 1) / n_rows;
    q->split = std::max<int64_t>(1, std::min<int64_t>(split, (n_local + 8191) / 8192));
    sdde::QsRanks ranks{};
    for (int r = 0; r < q->R; ++r) ranks.k[r] = q->plan.ranks[(size_t)r];
    sdde::qs_init_kernel<<<(unsigned)n_rows, 32, 0, c->compute>>>((sdde::QsRow *)c->qs_rows.p, q->R, ranks);
    CU_CHECK(cudaGetLastError());
    return SDDE_OK;
}

int qsel_local(sdde_ctx *c, int pass) {
    sdde_qsel *q = c->qsel;
    CU_CHECK(cudaMemsetAsync(q->hist, 0, q->hist_bytes, c->compute));
    sdde::qs_stream_fn(pass)<<<dim3((unsigned)q->n_rows, (unsigned)q->split), sdde::QS_THREADS, q->smem, c->compute>>>(
        q->src, q->pitch, q->n_local, q->R, (sdde::QsRow *)c->qs_rows.p, q->hist, (unsigned long long *)c->qs_cand.p);
    CU_CHECK(cudaGetLastError());
    return SDDE_OK;
}

int qsel_apply(sdde_ctx *c, int pass) {
    sdde_qsel *q = c->qsel;
    sdde::qs_scan_kernel<<<(unsigned)q->n_rows, 32 * q->R, 0, c->compute>>>((sdde::QsRow *)c->qs_rows.p, q->hist,
                                                                             (const unsigned long long *)c->qs_cand.p, q->R, pass, q->allow_gather);
    CU_CHECK(cudaGetLastError());
    return SDDE_OK;
}

int qsel_end(sdde_ctx *c, double *result) {
    sdde_qsel *q = c->qsel;
    std::vector<sdde::QsRow> rows((size_t)q->n_rows);
    CU_CHECK(cudaMemcpyAsync(rows.data(), c->qs_rows.p, (size_t)q->n_rows * sizeof(sdde::QsRow), cudaMemcpyDeviceToHost, c->compute));
    CU_CHECK(cudaStreamSynchronize(c->compute));
    const int ncol = (int)q->plan.items.size(), R = q->R;
    for (int64_t i = 0; i < q->n_rows; ++i) {
        const sdde::QsRow &row = rows[(size_t)i];
        double v[sdde::QS_MAX_RANK];
        for (int r = 0; r < R; ++r) {
            const unsigned long long key = row.prefix[r];
            const unsigned long long b = (key >> 63) ? (key & 0x7fffffffffffffffull) : ~key;
            std::memcpy(&v[r], &b, 8);
        }
        const bool has_nan = q->nan_global.empty() ? row.n_nan != 0 : q->nan_global[(size_t)i] != 0;
        for (int j = 0; j < ncol; ++j) {
            const sdde::QuantileItem &it = q->plan.items[(size_t)j];
            result[i * ncol + j] = has_nan ? std::nan("") : sdde::combine_quantile(it, v[it.ia], v[it.ib]);
        }
    }
    sdde_qsel_free(c->qsel); c->qsel = nullptr;
    return SDDE_OK;
}

}  // namespace

extern "C" {

int sdde_b200_quantiles(sdde_ctx *c, const double *paths, int32_t mem, int64_t n_rows, int64_t n_traj, int64_t pitch,
                        const double *probs, int32_t n_prob, int32_t want_median, double *result, double *kernel_ms) {
    if (!c || !paths || !result || n_traj <= 0) return fail(SDDE_ERR_BAD_DESC, "bad arguments");
    std::lock_guard<std::mutex> lk(c->mu);
    CU_CHECK(cudaSetDevice(c->device));
    CU_CHECK(cudaEventRecord(c->ev_k0, c->compute));
    int rc = qsel_begin(c, paths, mem, n_rows, n_traj, pitch, n_traj, probs, n_prob, want_median, nullptr, 1);
    for (int pass = 0; pass < sdde::QS_PASSES && !rc; ++pass) {
        rc = qsel_local(c, pass);
        if (!rc) rc = qsel_apply(c, pass);
    }
    if (rc) return rc;
    CU_CHECK(cudaEventRecord(c->ev_k1, c->compute));
    rc = qsel_end(c, result);
    if (rc) return rc;
    if (kernel_ms) { float ms = 0.f; CU_CHECK(cudaEventElapsedTime(&ms, c->ev_k0, c->ev_k1)); *kernel_ms = ms; }
    return SDDE_OK;
}

int sdde_b200_qsel_ranks(int64_t n_total, const double *probs, int32_t n_prob, int32_t want_median, int32_t *n_rank) {
    if (!n_rank || n_prob < 0 || (n_prob > 0 && !probs)) return fail(SDDE_ERR_BAD_DESC, "bad arguments");
    sdde::QuantilePlan plan;
    const std::string perr = sdde::plan_quantiles(n_total, probs, n_prob, want_median != 0, plan);
    if (!perr.empty()) return fail(SDDE_ERR_BAD_DESC, "%s", perr.c_str());
    if ((int)plan.ranks.size() > sdde::QS_MAX_RANK) return fail(SDDE_ERR_UNSUPPORTED, "at most %d distinct order statistics per call", sdde::QS_MAX_RANK);
    *n_rank = (int32_t)plan.ranks.size();
    return SDDE_OK;
}

int sdde_b200_qsel_begin(sdde_ctx *c, const double *paths, int32_t mem, int64_t n_rows, int64_t n_local, int64_t pitch, int64_t n_total,
                         const double *probs, int32_t n_prob, int32_t want_median, uint32_t *d_hist) {
    if (!c || !d_hist) return fail(SDDE_ERR_BAD_DESC, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CU_CHECK(cudaSetDevice(c->device));
    int rc = qsel_begin(c, paths, mem, n_rows, n_local, pitch, n_total, probs, n_prob, want_median, d_hist, 0);
    if (rc) return rc;
    CU_CHECK(cudaStreamSynchronize(c->compute));
    return SDDE_OK;
}

int sdde_b200_qsel_local(sdde_ctx *c, int32_t pass) {
    if (!c || !c->qsel || pass < 0 || pass >= sdde::QS_PASSES) return fail(SDDE_ERR_BAD_DESC, "no active selection, or pass out of range");
    std::lock_guard<std::mutex> lk(c->mu);
    CU_CHECK(cudaSetDevice(c->device));
    int rc = qsel_local(c, pass);
    if (rc) return rc;
    CU_CHECK(cudaStreamSynchronize(c->compute));       // the caller's collective runs on its own stream
    return SDDE_OK;
}

int sdde_b200_qsel_apply(sdde_ctx *c, int32_t pass) {
    if (!c || !c->qsel || pass < 0 || pass >= sdde::QS_PASSES) return fail(SDDE_ERR_BAD_DESC, "no active selection, or pass out of range");
    std::lock_guard<std::mutex> lk(c->mu);
    CU_CHECK(cudaSetDevice(c->device));
    int rc = qsel_apply(c, pass);
    if (rc) return rc;
    CU_CHECK(cudaStreamSynchronize(c->compute));
    return SDDE_OK;
}

int sdde_b200_qsel_nan(sdde_ctx *c, int64_t *counts, int32_t set) {
    if (!c || !c->qsel || !counts) return fail(SDDE_ERR_BAD_DESC, "no active selection");
    std::lock_guard<std::mutex> lk(c->mu);
    CU_CHECK(cudaSetDevice(c->device));
    sdde_qsel *q = c->qsel;
    if (set) { q->nan_global.assign(counts, counts + q->n_rows); return SDDE_OK; }
    std::vector<sdde::QsRow> rows((size_t)q->n_rows);
    CU_CHECK(cudaMemcpyAsync(rows.data(), c->qs_rows.p, (size_t)q->n_rows * sizeof(sdde::QsRow), cudaMemcpyDeviceToHost, c->compute));
    CU_CHECK(cudaStreamSynchronize(c->compute));
    for (int64_t i = 0; i < q->n_rows; ++i) counts[i] = (int64_t)rows[(size_t)i].n_nan;
    return SDDE_OK;
}

int sdde_b200_qsel_end(sdde_ctx *c, double *result) {
    if (!c || !c->qsel || !result) return fail(SDDE_ERR_BAD_DESC, "no active selection");
    std::lock_guard<std::mutex> lk(c->mu);
    CU_CHECK(cudaSetDevice(c->device));
    return qsel_end(c, result);
}

}  // extern "C"
