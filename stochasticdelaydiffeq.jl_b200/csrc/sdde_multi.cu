// sdde_multi.cu -- sdde_b200_solve_multi: several GPUs driven from ONE host process (the Julia `ccall` case).
// One host thread per context solves a contiguous global-id range; the per-GPU moment triples are all-gathered over
// NCCL (communicators from ncclCommInitAll, libnccl resolved at run time) and Chan-merged in rank order on the device.
// This is the path's single collective; what EnsembleDistributed does with pmap in SciMLBase.
#include <algorithm>
#include <cstring>
#include <thread>

#include "sdde_internal.cuh"
#include "nccl_dyn.hpp"
#define fail sdde_fail

namespace {

// NCCL communicators of a device group, created once per device list (ncclCommInitAll) and kept for the process.
struct NcclGroup {
    std::vector<int> devices;
    std::vector<ncclComm_t> comms;
};
std::mutex g_nccl_mu;
std::vector<NcclGroup *> g_nccl_groups;

NcclGroup *nccl_group_for(const std::vector<int> &devices, std::string &err) {
    const sdde::NcclApi &N = sdde::nccl_api();
    if (!N.ok) { err = N.why; return nullptr; }
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    for (NcclGroup *g : g_nccl_groups)
        if (g->devices == devices) return g;
    NcclGroup *g = new NcclGroup();
    g->devices = devices;
    g->comms.assign(devices.size(), nullptr);
    ncclResult_t r = N.CommInitAll(g->comms.data(), (int)devices.size(), devices.data());
    if (r != ncclSuccess) { err = std::string("ncclCommInitAll: ") + N.GetErrorString(r); delete g; return nullptr; }
    g_nccl_groups.push_back(g);
    return g;
}

}  // namespace

extern "C" {

int sdde_b200_solve_multi(sdde_ctx *const *ctxs, int32_t n_ctx, const sdde_desc *d, double *out, double *moments, uint8_t *status,
                          sdde_timing *tm) {
    if (!ctxs || n_ctx < 1) return fail(SDDE_ERR_BAD_DESC, "no contexts");
    for (int r = 0; r < n_ctx; ++r) if (!ctxs[r]) return fail(SDDE_ERR_BAD_DESC, "context %d is NULL", r);
    int rc = sdde_check_desc(d);
    if (rc) return rc;
    if (d->noise_t) return fail(SDDE_ERR_UNSUPPORTED, "sdde_b200_solve_multi takes supplied noise on the schedule's own nodes only");
    if (d->out_mem != SDDE_MEM_HOST || (d->noise_mode != SDDE_NOISE_PHILOX && d->noise_mem != SDDE_MEM_HOST))
        return fail(SDDE_ERR_UNSUPPORTED, "sdde_b200_solve_multi takes host buffers only");
    if (d->model_id >= SDDE_MODEL_USER_BASE) return fail(SDDE_ERR_UNSUPPORTED, "user models are compiled per context: use one context per call");
    if (n_ctx == 1) return sdde_b200_solve(ctxs[0], d, out, moments, status, tm);
    // trajectory-independent sizes
    int64_t rows = 0, noise_rows = 0;
    rc = sdde_plan_sizes(ctxs[0], d, &rows, &noise_rows);
    if (rc) return rc;
    const int64_t NT = d->n_traj;
    const int64_t OP = d->out_pitch > 0 ? d->out_pitch : NT;
    const int64_t per = (NT + n_ctx - 1) / n_ctx;

    // The path's single collective: all-gather of the per-GPU moment triples over NCCL + device merge in rank order.
    NcclGroup *grp = nullptr;
    std::string nccl_why;
    if (d->want_moments && rows > 0 && ctxs[0]->opt_nccl && (int64_t)(n_ctx - 1) * per < NT /* every GPU has work */) {
        std::vector<int> devs;
        for (int r = 0; r < n_ctx; ++r) devs.push_back(ctxs[r]->device);
        bool distinct = true;
        for (int r = 0; r < n_ctx && distinct; ++r) for (int q = 0; q < r; ++q) if (devs[q] == devs[r]) distinct = false;
        if (distinct) grp = nccl_group_for(devs, nccl_why);
    }
    for (int r = 0; r < n_ctx; ++r) ctxs[r]->keep_moments_on_device = (grp != nullptr);

    std::vector<int> rcs((size_t)n_ctx, 0);
    std::vector<std::string> errs((size_t)n_ctx);
    std::vector<std::vector<double>> moms((size_t)n_ctx);
    std::vector<std::thread> th;
    for (int r = 0; r < n_ctx; ++r) {
        th.emplace_back([&, r]() {
            const int64_t lo = std::min(NT, (int64_t)r * per), hi = std::min(NT, lo + per), n = hi - lo;
            if (n == 0) return;
            sdde_desc dr = *d;
            dr.n_traj = n; dr.traj_offset = d->traj_offset + lo; dr.out_pitch = OP;
            // per-trajectory inputs are [row][stride] arrays: the slice of this GPU starts at column lo.  params/u0 keep
            // their stride; supplied noise is re-packed densely (its row pitch is n_traj by definition).
            std::vector<double> nz;
            if (d->param_stride) dr.params = d->params + lo;
            if (d->u0_stride) dr.u0 = d->u0 + lo;
            if (d->noise_mode != SDDE_NOISE_PHILOX) {
                nz.resize((size_t)std::max<int64_t>(1, noise_rows * n));
                for (int64_t k = 0; k < noise_rows; ++k) memcpy(nz.data() + k * n, d->noise + k * NT + lo, (size_t)n * 8);
                dr.noise = nz.data();
            }
            if (d->want_moments) moms[(size_t)r].assign((size_t)std::max<int64_t>(1, rows * 3), 0.0);
            rcs[(size_t)r] = sdde_b200_solve(ctxs[r], &dr, d->want_paths ? out + lo : nullptr,
                                             d->want_moments ? moms[(size_t)r].data() : nullptr,
                                             status ? status + lo : nullptr, tm ? tm + r : nullptr);
            if (rcs[(size_t)r]) errs[(size_t)r] = sdde_last_error_string();     // thread-local -> hand to the caller's thread
        });
    }
    for (auto &t : th) t.join();
    for (int r = 0; r < n_ctx; ++r) ctxs[r]->keep_moments_on_device = false;
    for (int r = 0; r < n_ctx; ++r)
        if (rcs[(size_t)r]) return fail(rcs[(size_t)r], "GPU %d: %s", r, errs[(size_t)r].c_str());
    int merge_kind = 0;
    if (d->want_moments && rows > 0 && grp) {
        const sdde::NcclApi &N = sdde::nccl_api();
        const size_t cnt = (size_t)rows * 3;
        for (int r = 0; r < n_ctx; ++r) {
            CU_CHECK(cudaSetDevice(ctxs[r]->device));
            CU_CHECK(ctxs[r]->gather.reserve(cnt * 8 * (size_t)n_ctx));
        }
        ncclResult_t nr = N.GroupStart();
        for (int r = 0; r < n_ctx && nr == ncclSuccess; ++r) {
            CU_CHECK(cudaSetDevice(ctxs[r]->device));
            nr = N.AllGather(ctxs[r]->mom.p, ctxs[r]->gather.p, cnt, ncclDouble, grp->comms[(size_t)r], ctxs[r]->compute);
        }
        const ncclResult_t ne = N.GroupEnd();
        if (nr == ncclSuccess) nr = ne;
        if (nr != ncclSuccess) return fail(SDDE_ERR_NCCL - (int)nr, "NCCL all-gather of the moment triples failed: %s", N.GetErrorString(nr));
        // every GPU now holds all partial triples; merge in rank order on GPU 0 and read the result back
        sdde_ctx *c0 = ctxs[0];
        CU_CHECK(cudaSetDevice(c0->device));
        CU_CHECK(sdde_launch_merge_moments(c0->compute, (const double *)c0->gather.p, (long long)cnt, n_ctx, rows, (double *)c0->gather.p));
        CU_CHECK(cudaMemcpyAsync(moments, c0->gather.p, cnt * 8, cudaMemcpyDeviceToHost, c0->compute));
        for (int r = 0; r < n_ctx; ++r) {
            CU_CHECK(cudaSetDevice(ctxs[r]->device));
            CU_CHECK(cudaStreamSynchronize(ctxs[r]->compute));
        }
        merge_kind = 2;
    } else if (d->want_moments) {
        bool first = true;
        for (int r = 0; r < n_ctx; ++r) {
            if (moms[(size_t)r].empty()) continue;
            if (first) { memcpy(moments, moms[(size_t)r].data(), (size_t)rows * 3 * 8); first = false; }
            else sdde_b200_merge_moments(moments, moms[(size_t)r].data(), rows);
        }
        merge_kind = 1;
    }
    if (tm) for (int r = 0; r < n_ctx; ++r) tm[r].moment_merge = merge_kind;
    return SDDE_OK;
}

}  // extern "C"
