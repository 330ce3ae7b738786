// sdde_api.cu -- core of the C ABI of libsdde_b200.so (see include/sdde_b200.h): descriptor validation, host schedule,
// wave loop over trajectories, launches of the sm_100a step kernels, moments, output of the saved paths.
// (sdde_multi.cu: several GPUs from one process; sdde_summary_api.cu: order statistics; sdde_tools.cu: test hooks.)
//
// Replaces the ensemble dispatch [EXT] SciMLBase.__solve(::AbstractEnsembleProblem, alg, ensemblealg; trajectories)
// -> per trajectory DiffEqBase.__solve (/root/reference/src/solve.jl:1-9) for the fixed-step constant-lag case.
// No CPU fallback: every compute entry point needs a CUDA device.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>

#include "rng_table.hpp"
#include "sra_tableau.hpp"
#include "sdde_internal.cuh"
#include "sdde_dispatch.cuh"

namespace {
thread_local std::string g_last_error;
}

int sdde_fail(int code, const char *fmt, ...) {
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
const std::string &sdde_last_error_string() { return g_last_error; }
#define fail sdde_fail

// ------------------------------------------------------------------------------------------------
// small kernels owned by the solve path
// ------------------------------------------------------------------------------------------------
namespace sdde {

// (count, mean, M2) of one slice of one row of `out` per block, in ONE read of the data: sums of (x - c) and (x - c)^2
// about c = the slice's first value (a draw from the same distribution, so |c - mean| is a few standard deviations at most
// and M2 = S2 - S1^2/n loses a few bits, not digits); fixed reduction order => deterministic.  grid (rows, n_split); the
// slices of a row are Chan-merged in order by merge_moments_kernel.  Device side of
// [EXT] SciMLBase.EnsembleAnalysis.timeseries_steps_meanvar.
__global__ void __launch_bounds__(256) moments_kernel(const double *__restrict__ out, long long pitch, long long off,
                                                      long long n, double *__restrict__ part) {
    __shared__ double sh1[8], sh2[8];
    const long long rows = gridDim.x, row = blockIdx.x;       // rows on grid.x (save_everystep can exceed 65535 rows)
    const double *rp = out + row * pitch + off;
    const long long per = (n + gridDim.y - 1) / gridDim.y;
    const long long lo = per * blockIdx.y, hi = lo + per < n ? lo + per : n;
    double *dst = part + ((long long)blockIdx.y * rows + row) * 3;
    if (hi <= lo) {
        if (threadIdx.x == 0) { dst[0] = 0.0; dst[1] = 0.0; dst[2] = 0.0; }
        return;
    }
    const double c = rp[lo];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double s1 = 0.0, s2 = 0.0;
    long long i = lo + threadIdx.x;
    for (; i + 3 * 256 < hi; i += 4 * 256) {              // four independent loads in flight per thread
        const double d0 = rp[i] - c, d1 = rp[i + 256] - c, d2 = rp[i + 512] - c, d3 = rp[i + 768] - c;
        s1 += (d0 + d1) + (d2 + d3);
        s2 += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    for (; i < hi; i += 256) { const double d = rp[i] - c; s1 += d; s2 += d * d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if (lane == 0) { sh1[warp] = s1; sh2[warp] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < 8; ++w) { t1 += sh1[w]; t2 += sh2[w]; }
        const double cnt = (double)(hi - lo);
        const double m2 = t2 - t1 * (t1 / cnt);
        dst[0] = cnt;
        dst[1] = c + t1 / cnt;
        dst[2] = m2 > 0.0 || m2 != m2 ? m2 : 0.0;
    }
}

// number of non-zero status bytes (trajectories that saw a non-finite state)
__global__ void __launch_bounds__(256) count_status_kernel(const unsigned char *__restrict__ st, long long n, unsigned int *__restrict__ cnt) {
    unsigned int c = 0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) c += st[i] != 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(cnt, c);
}

// Chan merge of `parts` arrays of (count, mean, M2) triples, in order: dst[i] = merge(src[0][i], src[1][i], ...).
// One thread per triple; used across trajectory waves and, after the NCCL all-gather, across GPUs.
__global__ void merge_moments_kernel(const double *src, long long part_stride, int parts, long long n, double *dst) {   // dst may be src (in-place wave merge)
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double cnt = src[3 * i], mean = src[3 * i + 1], m2 = src[3 * i + 2];
    for (int p = 1; p < parts; ++p) {
        const double *s = src + (long long)p * part_stride + 3 * i;
        const double nb = s[0], mb = s[1], sb = s[2];
        const double nn = cnt + nb;
        if (nn > 0) {
            const double dlt = mb - mean;
            mean = mean + dlt * (nb / nn);
            m2 = m2 + sb + dlt * dlt * (cnt * nb / nn);
        }
        cnt = nn;
    }
    dst[3 * i] = cnt; dst[3 * i + 1] = mean; dst[3 * i + 2] = m2;
}

// W on an arbitrary NoiseGrid -> W on the schedule's nodes (rows (k, c) of dst), linear in time; see plan_noise_grid.
__global__ void __launch_bounds__(256) noise_regrid_kernel(const double *__restrict__ src, long long spitch, long long soff,
                                                           const GridNode *__restrict__ nodes, int nc, long long nt,
                                                           double *__restrict__ dst, long long dpitch) {
    const long long row = blockIdx.x;                       // k * nc + c; rows on grid.x (grid.y is capped at 65535)
    const GridNode g = nodes[row / nc];
    const int c = (int)(row % nc);
    const double *lo = src + ((long long)g.lo * nc + c) * spitch + soff, *hi = src + ((long long)g.hi * nc + c) * spitch + soff;
    for (long long tr = (long long)blockIdx.y * 256 + threadIdx.x; tr < nt; tr += (long long)gridDim.y * 256)
        dst[row * dpitch + tr] = g.lo == g.hi ? hi[tr] : fma(g.theta, hi[tr], (1.0 - g.theta) * lo[tr]);
}

}  // namespace sdde

cudaError_t sdde_launch_merge_moments(cudaStream_t stream, const double *src, long long part_stride, int parts, long long n, double *dst) {
    sdde::merge_moments_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(src, part_stride, parts, n, dst);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// descriptor -> schedule input / kernel selection
// ------------------------------------------------------------------------------------------------
namespace {

inline bool is_sra(int alg) { return alg >= SDDE_ALG_SRA1 && alg <= SDDE_ALG_SRA_TABLEAU; }

int check_desc(const sdde_desc *d) {
    if (!d) return fail(SDDE_ERR_BAD_DESC, "descriptor is NULL");
    if (d->struct_size != (int32_t)sizeof(sdde_desc))
        return fail(SDDE_ERR_BAD_DESC, "sdde_desc.struct_size = %d, library expects %d (ABI mismatch)", d->struct_size,
                    (int)sizeof(sdde_desc));
    if (d->alg < SDDE_ALG_EM_SPLIT || d->alg > SDDE_ALG_SRA_TABLEAU)
        return fail(SDDE_ERR_UNSUPPORTED, "algorithm %d is not on this path (EM / RKMil / RKMilCommute / EulerHeun / WangLi3SMil / SRA, fixed step)", d->alg);
    if (is_sra(d->alg)) {
        sdde::SraHost tab;
        const std::string err = sdde::sra_resolve(d->alg, d->sra_tableau, d->sra_stages, tab);
        if (!err.empty()) return fail(SDDE_ERR_BAD_DESC, "%s", err.c_str());
        if (d->n_lag * SDDE_SRA_MAX_SETS > 8) return fail(SDDE_ERR_UNSUPPORTED, "the SRA kernels take at most 2 lags");
    }
    if (d->n_lag < 1 || d->n_lag > SDDE_MAX_LAG) return fail(SDDE_ERR_BAD_DESC, "n_lag = %d out of range", d->n_lag);
    if (d->n_state < 1 || d->n_state > SDDE_MAX_STATE) return fail(SDDE_ERR_BAD_DESC, "n_state = %d out of range", d->n_state);
    if (d->n_noise < 1 || d->n_noise > SDDE_MAX_NOISE) return fail(SDDE_ERR_BAD_DESC, "n_noise = %d out of range", d->n_noise);
    if (d->n_param < 0 || d->n_param > SDDE_MAX_PARAM) return fail(SDDE_ERR_BAD_DESC, "n_param = %d out of range", d->n_param);
    if (d->n_param > 0 && !d->params) return fail(SDDE_ERR_BAD_DESC, "params is NULL");
    if (!d->u0) return fail(SDDE_ERR_BAD_DESC, "u0 is NULL");
    if (d->n_saveat < 0 || (d->n_saveat > 0 && !d->saveat)) return fail(SDDE_ERR_BAD_DESC, "saveat array missing");
    if (d->n_tstops < 0 || (d->n_tstops > 0 && !d->tstops)) return fail(SDDE_ERR_BAD_DESC, "tstops array missing");
    if (d->n_d_disc < 0 || (d->n_d_disc > 0 && (!d->d_disc_t || !d->d_disc_order2)))
        return fail(SDDE_ERR_BAD_DESC, "d_discontinuities arrays missing");
    if (d->n_traj < 0) return fail(SDDE_ERR_BAD_DESC, "n_traj < 0");
    if (d->history_mode != SDDE_HISTORY_DENSE)
        return fail(SDDE_ERR_UNSUPPORTED, "only the dense history mode exists on the GPU path (SURVEY.md H1)");
    if (d->noise_mode < SDDE_NOISE_PHILOX || d->noise_mode > SDDE_NOISE_SUPPLIED_GRID)
        return fail(SDDE_ERR_BAD_DESC, "noise_mode = %d", d->noise_mode);
    if (d->noise_mode != SDDE_NOISE_PHILOX && !d->noise && d->n_traj > 0)
        return fail(SDDE_ERR_BAD_DESC, "noise_mode is SUPPLIED but noise is NULL");
    if (d->noise_t && (d->noise_mode != SDDE_NOISE_SUPPLIED_GRID || d->n_noise_t < 2))
        return fail(SDDE_ERR_BAD_DESC, "noise_t needs noise_mode = SDDE_NOISE_SUPPLIED_GRID and n_noise_t >= 2");
    if (d->param_stride != 0 && d->param_stride < d->n_traj) return fail(SDDE_ERR_BAD_DESC, "param_stride < n_traj");
    if (d->u0_stride != 0 && d->u0_stride < d->n_traj) return fail(SDDE_ERR_BAD_DESC, "u0_stride < n_traj");
    if (d->n_save_idxs < 0 || d->n_save_idxs > d->n_state) return fail(SDDE_ERR_BAD_DESC, "n_save_idxs out of range");
    for (int i = 0; i < d->n_save_idxs; ++i) {
        if (d->save_idxs[i] < 0 || d->save_idxs[i] >= d->n_state) return fail(SDDE_ERR_BAD_DESC, "save_idxs[%d] = %d out of range", i, d->save_idxs[i]);
        for (int k = 0; k < i; ++k) if (d->save_idxs[k] == d->save_idxs[i]) return fail(SDDE_ERR_BAD_DESC, "duplicate entry in save_idxs");
    }
    return SDDE_OK;
}

std::string schedule_key(const sdde_desc *d) {
    std::string k;
    auto add = [&k](const void *p, size_t n) { k.append((const char *)p, n); };
    const int32_t ints[] = {d->alg, d->n_lag, d->order_discontinuity_t0, d->constant_lags_declared, d->save_start, d->save_end,
                            d->save_everystep, d->n_saveat, d->n_tstops};
    add(ints, sizeof(ints));
    const double dbl[] = {d->t0, d->tf, d->dt};
    add(dbl, sizeof(dbl));
    add(d->lags, sizeof(double) * (size_t)d->n_lag);
    if (d->n_saveat > 0) add(d->saveat, sizeof(double) * (size_t)d->n_saveat);
    if (d->n_tstops > 0) add(d->tstops, sizeof(double) * (size_t)d->n_tstops);
    add(&d->n_d_disc, sizeof(d->n_d_disc));
    if (d->n_d_disc > 0) { add(d->d_disc_t, sizeof(double) * (size_t)d->n_d_disc); add(d->d_disc_order2, sizeof(int32_t) * (size_t)d->n_d_disc); }
    if (d->alg == SDDE_ALG_SRA_TABLEAU && d->sra_tableau && d->sra_stages >= 1 && d->sra_stages <= 3) {   // stage times and order shape the schedule
        add(&d->sra_stages, sizeof(d->sra_stages)); add(&d->sra_order2, sizeof(d->sra_order2));
        add(d->sra_tableau, sizeof(double) * (size_t)(2 * d->sra_stages));
    }
    return k;
}

int make_schedule(const sdde_desc *d, sdde::Schedule &S) {
    sdde::ScheduleInput in;
    in.t0 = d->t0; in.tf = d->tf; in.dt = d->dt;
    in.lags.assign(d->lags, d->lags + d->n_lag);
    in.alg_order2 = (d->alg == SDDE_ALG_RKMIL_ITO || d->alg == SDDE_ALG_RKMIL_STRAT || d->alg == SDDE_ALG_RKMILC_ITO ||
                     d->alg == SDDE_ALG_RKMILC_STRAT || d->alg >= SDDE_ALG_WANGLI3SMIL_A) ? 2 : 1;   // [EXT] alg_order: RKMil, RKMilCommute, WangLi3SMil 1//1; EM, EulerHeun 1//2
    in.n_stage = d->alg == SDDE_ALG_EULER_HEUN ? 2 : 1;
    if (is_sra(d->alg)) {
        // lookup set k at t_n + cset[k] dt_n: the deduplicated evaluation times of the tableau's f and g calls (unused sets repeat t_n)
        sdde::SraHost tab;
        const std::string terr = sdde::sra_resolve(d->alg, d->sra_tableau, d->sra_stages, tab);
        if (!terr.empty()) return fail(SDDE_ERR_BAD_DESC, "%s", terr.c_str());
        in.alg_order2 = (d->alg == SDDE_ALG_SRA_TABLEAU && d->sra_order2 > 0) ? d->sra_order2 : tab.order2;
        in.n_stage = SDDE_SRA_MAX_SETS;
        in.stage_c.assign(tab.cset, tab.cset + SDDE_SRA_MAX_SETS);
    }
    in.order_discontinuity_t0 = d->order_discontinuity_t0;
    in.constant_lags_declared = d->constant_lags_declared != 0;
    if (d->n_saveat > 0) in.saveat.assign(d->saveat, d->saveat + d->n_saveat);
    if (d->n_tstops > 0) in.tstops.assign(d->tstops, d->tstops + d->n_tstops);
    if (d->n_d_disc > 0) { in.disc_t.assign(d->d_disc_t, d->d_disc_t + d->n_d_disc); in.disc_order2.assign(d->d_disc_order2, d->d_disc_order2 + d->n_d_disc); }
    in.save_start = d->save_start != 0;
    in.save_end = d->save_end != 0;
    in.save_everystep = d->save_everystep != 0;
    std::string err = sdde::build_schedule(in, S);
    if (!err.empty()) return fail(SDDE_ERR_BAD_DESC, "%s", err.c_str());
    // src/solve.jl:581-585: dt > minimum(constant_lags) is an error for every algorithm but EM
    if (d->alg != SDDE_ALG_EM_SPLIT && d->alg != SDDE_ALG_EM_NOSPLIT) {
        double minlag = d->lags[0];
        for (int l = 1; l < d->n_lag; ++l) minlag = std::min(minlag, d->lags[l]);
        for (double x : S.dt)
            if (x > minlag)
                return fail(SDDE_ERR_UNSUPPORTED,
                            "dt > minimum(constant_lags). This is not allowed by the integrator. Please dt `dtmax` "
                            "less than the minimum lag or use `EM`.");
    }
    return SDDE_OK;
}

// CUDA source of the n-state linear delay system with THIS problem's matrices baked in as literals: zero entries
// disappear at compile time and the remaining coefficients become immediates instead of one constant load (LDCU) per
// matrix entry and step (tools/micro/ldcu.cu).  Same operation order as ModelLinSys, so results are unchanged
// (a dropped term is +0*x).  Compiled once per (shape, parameters) and context through the user-model machinery.
std::string linsys_jit_source(int ns, int nl, int nw, const double *p) {
    char buf[128];
    auto lit = [&](double v) { snprintf(buf, sizeof(buf), "%.17g", v); std::string t(buf); if (t.find_first_of(".eEn") == std::string::npos) t += ".0"; return t; };
    const int offB = ns * ns, offC = ns * ns * (1 + nl), offM = offC + ns, offN = offM + ns * nw;
    std::string src = "struct UserModel {\n";
    snprintf(buf, sizeof(buf), "    static constexpr int NS = %d, NL = %d, NW = %d, NP = %d;\n", ns, nl, nw, offN + ns * nw);
    src += buf;
    src += "    static constexpr bool DIAG = false;\n";
    src += "    template <class P> __device__ __forceinline__ static void f(double *du, const double *u, const double (*ul)[NS], const P &, double) {\n";
    for (int i = 0; i < ns; ++i) {
        // acc = p[i*ns]*u[0] (+ ...)  -- keep the first product even when its coefficient is zero so that the
        // accumulation starts exactly like the generic model
        std::string e = "(" + lit(p[i * ns]) + " * u[0])";
        for (int j = 1; j < ns; ++j) if (p[i * ns + j] != 0.0) e = "(" + e + " + " + lit(p[i * ns + j]) + " * u[" + std::to_string(j) + "])";
        for (int l = 0; l < nl; ++l)
            for (int j = 0; j < ns; ++j) {
                const double v = p[offB + l * ns * ns + i * ns + j];
                if (v != 0.0) e = "(" + e + " + " + lit(v) + " * ul[" + std::to_string(l) + "][" + std::to_string(j) + "])";
            }
        if (p[offC + i] != 0.0) e = "(" + e + " + " + lit(p[offC + i]) + ")";
        src += "        du[" + std::to_string(i) + "] = " + e + ";\n";
    }
    src += "    }\n";
    src += "    template <class P> __device__ __forceinline__ static void g(double *gv, const double *u, const double (*)[NS], const P &, double) {\n";
    for (int i = 0; i < ns; ++i)
        for (int j = 0; j < nw; ++j) {
            const double m = p[offM + i * nw + j], n0 = p[offN + i * nw + j];
            std::string e = (m != 0.0) ? (lit(m) + " * u[" + std::to_string(i) + "]") : std::string("0.0");
            if (n0 != 0.0) e = e + " + " + lit(n0);
            src += "        gv[" + std::to_string(i * nw + j) + "] = " + e + ";\n";
        }
    src += "    }\n};\n";
    return src;
}

int select_kernel_alg(sdde_ctx *ctx, const sdde_desc *d, int alg, sdde::KernelInfo &k);

// RKMilCommute: a noise-matrix model has its own kernels (alg 5 / 6); on a diagonal model the update is the RKMil one
// written with the iterated integral, so the RKMil kernels (alg 2 / 3) serve
// WangLi3SMil_A..F -> stage operators in execution order, 2 bits each (0 = D drift, 1 = C Ito correction, 2 = M diffusion)
static int wangli_stage_order(int alg) {
    static const int order[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};   // A: D C M, B: D M C, C: C D M, D: C M D, E: M D C, F: M C D
    const int *o = order[alg - SDDE_ALG_WANGLI3SMIL_A];
    return o[0] | (o[1] << 2) | (o[2] << 4);
}

int select_kernel(sdde_ctx *ctx, const sdde_desc *d, sdde::KernelInfo &k) {
    if (is_sra(d->alg)) return select_kernel_alg(ctx, d, sdde::K_ALG_SRA, k);
    if (d->alg >= SDDE_ALG_WANGLI3SMIL_A) {
        const int rc = select_kernel_alg(ctx, d, sdde::K_ALG_WANGLI, k);
        if (rc == SDDE_OK && !k.diag) return fail(SDDE_ERR_UNSUPPORTED, "WangLi3SMil is implemented for diagonal noise only");
        return rc;
    }
    if (d->alg == SDDE_ALG_RKMILC_ITO || d->alg == SDDE_ALG_RKMILC_STRAT) {
        sdde::KernelInfo probe;
        const int as_rkmil = d->alg - SDDE_ALG_RKMILC_ITO + SDDE_ALG_RKMIL_ITO;
        if (select_kernel_alg(ctx, d, as_rkmil, probe) == SDDE_OK && probe.fn && probe.diag) { k = probe; return SDDE_OK; }
    }
    return select_kernel_alg(ctx, d, d->alg, k);
}

int select_kernel_alg(sdde_ctx *ctx, const sdde_desc *d, int alg, sdde::KernelInfo &k) {
    const bool ptraj = d->param_stride != 0;
    const int nz = d->noise_mode;
    switch (d->model_id) {
        case SDDE_MODEL_LINEAR1: k = sdde::select_linear1(alg, nz, ptraj); break;
        case SDDE_MODEL_HAYES: k = sdde::select_hayes(alg, nz, ptraj); break;
        case SDDE_MODEL_MACKEY_GLASS:
            if (!ptraj && d->n_param == 4 && d->params && d->params[2] == 10.0) k = sdde::select_mackey_glass10(alg, nz, ptraj);
            else k = sdde::select_mackey_glass(alg, nz, ptraj);
            break;
        case SDDE_MODEL_SCALAR_ADD: k = sdde::select_scalar_add(alg, nz, ptraj); break;
        case SDDE_MODEL_SIR: k = sdde::select_sir(alg, nz, ptraj); break;
        case SDDE_MODEL_LINSYS:
            if (d->lanes_per_traj > 1) {
                if (ptraj) return fail(SDDE_ERR_UNSUPPORTED, "the cooperative-lanes kernel takes shared parameters only");
                k = sdde::select_linsys_coop(d->n_state, d->n_lag, d->n_noise, alg, nz);
                if (k.fn && k.lanes != d->lanes_per_traj) return fail(SDDE_ERR_UNSUPPORTED, "lanes_per_traj must be 0, 1 or n_state (%d)", d->n_state);
                if (!k.fn) return fail(SDDE_ERR_UNSUPPORTED, "no cooperative-lanes kernel for this model shape / algorithm");
            } else {
                k = sdde::KernelInfo();
                const int npar = d->n_state * d->n_state * (1 + d->n_lag) + d->n_state + 2 * d->n_state * d->n_noise;
                bool finite = true;
                if (!ptraj && d->params && d->n_param == npar)
                    for (int i = 0; i < npar; ++i) finite = finite && std::isfinite(d->params[i]);
                if (ctx && ctx->opt_jit_linsys && !ptraj && d->params && d->n_param == npar && finite) {
                    std::string key;
                    const int32_t shp[3] = {d->n_state, d->n_lag, d->n_noise};
                    key.append((const char *)shp, sizeof(shp));
                    key.append((const char *)d->params, sizeof(double) * (size_t)npar);
                    int id = -1;
                    for (auto &e : ctx->linsys_jit) if (e.first == key) id = e.second;
                    std::string err;
                    if (id < 0) {
                        char nm[64];
                        snprintf(nm, sizeof(nm), "linsys_%d_%d_%d_nvrtc_constants", d->n_state, d->n_lag, d->n_noise);
                        id = ctx->user.compile(linsys_jit_source(d->n_state, d->n_lag, d->n_noise, d->params).c_str(), err, nm);
                        ctx->linsys_jit.emplace_back(key, id);          // id < 0: remembered as "NVRTC unavailable", generic kernel below
                    }
                    if (id >= 0 && ctx->user.select(id, alg, nz, false, k, err)) k.n_param = d->n_param;
                    else k = sdde::KernelInfo();
                }
                if (!k.fn) k = sdde::select_linsys(d->n_state, d->n_lag, d->n_noise, alg, nz, ptraj);
            }
            break;
        default:
            if (d->model_id >= SDDE_MODEL_USER_BASE && ctx) {
                std::string err;
                if (!ctx->user.select(d->model_id, alg, nz, ptraj, k, err))
                    return fail(SDDE_ERR_UNSUPPORTED, "%s", err.c_str());
                break;
            }
            return fail(SDDE_ERR_UNSUPPORTED, "unknown model id %d", d->model_id);
    }
    if (!k.fn) {
        if ((alg == SDDE_ALG_RKMIL_ITO || alg == SDDE_ALG_RKMIL_STRAT))
            return fail(SDDE_ERR_UNSUPPORTED, "RKMil requires diagonal noise (model %d has a non-diagonal noise matrix)",
                        d->model_id);
        if (alg == sdde::K_ALG_SRA)
            return fail(SDDE_ERR_UNSUPPORTED, "SRA1 / SRA3 / SOSRA / SRA(tableau) need additive noise: built-in models hayes and scalar_add, "
                        "or a compiled user model (model %d has no SRA kernel)", d->model_id);
        if (alg == sdde::K_ALG_WANGLI)
            return fail(SDDE_ERR_UNSUPPORTED, "WangLi3SMil is implemented for diagonal noise only (model %d has a noise matrix)", d->model_id);
        return fail(SDDE_ERR_UNSUPPORTED, "no kernel for model %d shape (%d states, %d lags, %d noise) alg %d", d->model_id,
                    d->n_state, d->n_lag, d->n_noise, alg);
    }
    if (k.n_state != d->n_state || k.n_lag != d->n_lag || k.n_noise != d->n_noise)
        return fail(SDDE_ERR_BAD_DESC, "model %d expects (n_state,n_lag,n_noise) = (%d,%d,%d), descriptor has (%d,%d,%d)",
                    d->model_id, k.n_state, k.n_lag, k.n_noise, d->n_state, d->n_lag, d->n_noise);
    if (k.n_param != d->n_param)
        return fail(SDDE_ERR_BAD_DESC, "model %d expects %d parameters, descriptor has %d", d->model_id, k.n_param, d->n_param);
    return SDDE_OK;
}

void pack_tables(const sdde::Schedule &S, const sdde::KernelInfo &k, int64_t pitch, int batch, int chunk, bool interp_body,
                 std::vector<unsigned char> &steps, std::vector<sdde::SaveRec> &saves) {
    const int64_t N = S.n_steps();
    const int L = S.n_lag;
    const int64_t R = S.ring_slots;
    const int64_t slot_bytes = (int64_t)k.n_state * pitch * 8;     // one ring slot = [n_state][pitch] doubles
    steps.assign((size_t)N * k.rec_size, 0);
    std::vector<int> nsave((size_t)N, 0);
    saves.resize(S.save_t.size());
    for (size_t i = 0; i < S.save_t.size(); ++i) {
        saves[i].step = (int)S.save_step[i];
        saves[i].exact = S.save_exact[i];
        saves[i].theta = S.save_theta[i];
        if (S.save_step[i] >= 0) nsave[(size_t)S.save_step[i]]++;
    }
    std::vector<long long> prev_hi((size_t)std::max(1, L * S.n_stage), -1);     // byte offset of the previous step's upper node, -1: none
    for (int64_t n = 0; n < N; ++n) {
        unsigned char *rec = steps.data() + (size_t)n * k.rec_size;
        double hdr[3] = {S.t[(size_t)n], S.dt[(size_t)n], S.sqdt[(size_t)n]};
        memcpy(rec, hdr, 24);
        int mask = 0;
        const int NLK = L * S.n_stage;
        for (int l = 0; l < NLK; ++l) {            // lookup l = stage * L + lag, same order as Schedule::lk_*
            const size_t e = (size_t)n * NLK + l;
            const int mode = S.lk_mode[e];
            long long off_lo = 0, off_hi = 0;
            double theta = S.lk_theta[e];
            if (mode == sdde::LK_EXACT) {            // exact node: lo = hi = node, theta = 1  ->  fma(1, hi, 0*lo) == hi
                off_lo = off_hi = (S.lk_idx[e] % R) * slot_bytes; theta = 1.0; mask |= 1 << l;
            } else if (mode == sdde::LK_INTERP) {
                off_lo = ((S.lk_idx[e] - 1) % R) * slot_bytes; off_hi = (S.lk_idx[e] % R) * slot_bytes; mask |= 1 << l;
            } else if (mode == sdde::LK_U0) {
                mask |= 0x10000 << l;
            }
            // consecutive steps usually bracket with a shared node: lo(n) == hi(n-1) -> the kernel keeps it in registers
            const bool ring = mode == sdde::LK_EXACT || mode == sdde::LK_INTERP;
            if (ring && prev_hi[(size_t)l] == off_lo) mask |= 0x100 << l;
            prev_hi[(size_t)l] = ring ? off_hi : -1;
            unsigned char *lk = rec + 48 + 32 * l;
            memcpy(lk, &theta, 8);
            memcpy(lk + 8, &off_lo, 8);
            memcpy(lk + 16, &off_hi, 8);
        }
        int ns[2] = {nsave[(size_t)n], mask};
        memcpy(rec + 24, ns, 8);
        const long long off_store = ((n + 1) % R) * slot_bytes;
        memcpy(rec + 32, &off_store, 8);
        const double inv_sq = 1.0 / S.sqdt[(size_t)n];
        memcpy(rec + 40, &inv_sq, 8);
    }
    // Mark the batches (of the size the kernel will run: `batch` steps, starting at multiples of `batch` inside every
    // chunk) that qualify for a lean body of run_batch: every lookup of every step reads the ring, the stored slots are
    // consecutive ring slots without a wrap, dt is bitwise the same for all steps, at most 255 saves per step; "exact"
    // when in addition every lookup hits a node exactly and the nodes are consecutive slots.  The mark and the packed
    // save counts live in the batch's first record.
    if (batch < 1) return;                                  // option "lean" = 0: every batch takes the general body
    const int NLK = L * S.n_stage;
    auto rd64 = [](const unsigned char *p) { long long v; memcpy(&v, p, 8); return v; };
    for (int64_t c0 = 0; c0 < N; c0 += chunk) {
        const int64_t cend = std::min<int64_t>(N, c0 + chunk);
        for (int64_t n = c0; n + batch <= cend; n += batch) {
            unsigned char *r0 = steps.data() + (size_t)n * k.rec_size;
            int m0;
            memcpy(&m0, r0 + 28, 4);
            bool ok = true, all_exact = true, all_interp = true;
            unsigned long long packed = 0;
            for (int kk = 0; kk < batch && ok; ++kk) {
                const unsigned char *r = steps.data() + (size_t)(n + kk) * k.rec_size;
                ok = ok && memcmp(r + 8, r0 + 8, 16) == 0;                                   // dt, sqrt|dt|
                ok = ok && rd64(r + 32) == rd64(r0 + 32) + kk * slot_bytes;                  // stored slots consecutive
                ok = ok && nsave[(size_t)(n + kk)] <= 255;
                packed |= (unsigned long long)(nsave[(size_t)(n + kk)] & 255) << (8 * kk);
                for (int l = 0; l < NLK && ok; ++l) {
                    const int mode = S.lk_mode[(size_t)(n + kk) * NLK + l];
                    ok = ok && (mode == sdde::LK_EXACT || mode == sdde::LK_INTERP);          // every lookup reads the ring
                    const unsigned char *lk = r + 48 + 32 * l, *lk0 = r0 + 48 + 32 * l;
                    const bool consec = rd64(lk + 16) == rd64(lk0 + 16) + kk * slot_bytes;    // upper nodes on consecutive slots
                    all_exact = all_exact && mode == sdde::LK_EXACT && consec;
                    all_interp = all_interp && mode == sdde::LK_INTERP && consec && rd64(lk + 8) == rd64(lk + 16) - slot_bytes;
                }
            }
            ok = ok && batch <= 8;
            if (ok) {
                m0 |= all_exact ? SDDE_MASK_FAST_EXACT : (all_interp && batch > 1 && interp_body) ? SDDE_MASK_FAST_INTERP : SDDE_MASK_FAST_RING;
                memcpy(r0 + 28, &m0, 4);
                memcpy(r0 + 48 + 24, &packed, 8);            // lk[0].pad
            }
        }
    }
}

}  // namespace

int sdde_check_desc(const sdde_desc *d) { return check_desc(d); }

int sdde_plan_sizes(sdde_ctx *ctx, const sdde_desc *d, int64_t *rows, int64_t *noise_rows) {
    sdde::KernelInfo K;
    int rc = select_kernel(ctx, d, K);
    if (rc) return rc;
    sdde::Schedule S;
    rc = make_schedule(d, S);
    if (rc) return rc;
    *rows = (int64_t)S.save_t.size() * (d->n_save_idxs > 0 ? d->n_save_idxs : d->n_state);
    *noise_rows = (S.n_steps() + (d->noise_mode == SDDE_NOISE_SUPPLIED_GRID ? 1 : 0)) * (int64_t)K.n_chan;
    return SDDE_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int sdde_b200_version(void) { return SDDE_B200_VERSION; }

const char *sdde_b200_last_error(void) { return g_last_error.c_str(); }

int sdde_b200_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(SDDE_ERR_CUDA - (int)e, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }
    return n;
}

sdde_ctx *sdde_b200_create(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        fail(SDDE_ERR_CUDA - (int)e, "no CUDA device available (%s); libsdde_b200 has no CPU fallback",
             e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return nullptr;
    }
    if (device < 0 || device >= n) { fail(SDDE_ERR_BAD_DESC, "device %d out of range (0..%d)", device, n - 1); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { fail(SDDE_ERR_CUDA, "cudaSetDevice(%d) failed", device); return nullptr; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10) {
        fail(SDDE_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library ships sm_100a code only", device, prop.major, prop.minor);
        return nullptr;
    }
    sdde_ctx *c = new sdde_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreate(&c->ev_k0) == cudaSuccess && cudaEventCreate(&c->ev_k1) == cudaSuccess &&
              cudaEventCreate(&c->ev_c0) == cudaSuccess && cudaEventCreate(&c->ev_c1) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i)
        ok = cudaEventCreateWithFlags(&c->ev_kdone[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { fail(SDDE_ERR_CUDA, "stream/event creation failed"); sdde_b200_destroy(c); return nullptr; }
    const size_t tab_bytes = (size_t)sdde::rng_table_entries() * 16;
    if (c->rng_tab.reserve(tab_bytes) != cudaSuccess ||
        cudaMemcpy(c->rng_tab.p, sdde::rng_table_host(), tab_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        fail(SDDE_ERR_CUDA, "upload of the normal generator's table failed");
        sdde_b200_destroy(c);
        return nullptr;
    }
    return c;
}

void sdde_b200_destroy(sdde_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    c->rng_tab.release();
    c->ring.release(); c->out[0].release(); c->out[1].release(); c->steps.release(); c->saves.release();
    c->noise.release(); c->u0.release(); c->ptraj.release(); c->status.release(); c->mom.release();
    c->h_status.release(); c->h_mom.release(); c->h_count.release(); c->gather.release();
    c->bounce[0].release(); c->bounce[1].release();
    for (cudaEvent_t &e : c->ev_bounce) { if (e) cudaEventDestroy(e); e = nullptr; }
    c->pool.reset();
    c->noise_src.release(); c->noise_nodes.release(); c->mom_part.release();
    sdde_qsel_free(c->qsel); c->qsel = nullptr;
    c->qs_rows.release(); c->qs_hist.release(); c->qs_paths.release(); c->qs_cand.release();
    c->user.clear();
    if (c->compute) cudaStreamDestroy(c->compute);
    if (c->copy_in) cudaStreamDestroy(c->copy_in);
    if (c->copy_out) cudaStreamDestroy(c->copy_out);
    cudaEvent_t evs[] = {c->ev_k0, c->ev_k1, c->ev_c0, c->ev_c1, c->ev_in, c->ev_kdone[0], c->ev_kdone[1],
                         c->ev_copied[0], c->ev_copied[1]};
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    delete c;
}

int sdde_b200_set_option(sdde_ctx *c, const char *key, int64_t value) {
    if (!c || !key) return fail(SDDE_ERR_BAD_DESC, "NULL context or key");
    std::lock_guard<std::mutex> lk(c->mu);
    if (!strcmp(key, "wave_traj")) c->opt_wave_traj = value;
    else if (!strcmp(key, "batch")) c->opt_batch = value;
    else if (!strcmp(key, "nccl")) c->opt_nccl = value;
    else if (!strcmp(key, "jit_linsys")) c->opt_jit_linsys = value;
    else if (!strcmp(key, "mem_limit")) { c->opt_mem_limit = value; c->mem_key.clear(); }
    else if (!strcmp(key, "zero_copy")) { c->opt_zero_copy = value < 0 ? -1 : (value != 0); for (auto &z : c->zc) z = sdde_ctx::ZcTrial(); }   // -1 = auto (default)
    else if (!strcmp(key, "drop_caches")) { c->sched_key.clear(); c->table_key.clear(); }   // next solve replays the schedule and re-uploads its tables
    else if (!strcmp(key, "lean")) { c->opt_lean = value; c->table_key.clear(); }   // 0: general body only, 1: all lean bodies, 2: without the consecutive-interpolation body (A/B)
    else if (!strcmp(key, "balance")) c->opt_balance = value != 0;
    else if (!strcmp(key, "bounce_min_mb")) { if (value < 0) return fail(SDDE_ERR_BAD_DESC, "bounce_min_mb out of range"); c->opt_bounce_min_mb = value; }
    else if (!strcmp(key, "host_threads")) { if (value < 0 || value > 256) return fail(SDDE_ERR_BAD_DESC, "host_threads out of range"); c->opt_host_threads = value; c->pool.reset(); }
    else if (!strcmp(key, "ramp_start")) { if (value < 256 || value > ((int64_t)1 << 22)) return fail(SDDE_ERR_BAD_DESC, "ramp_start out of range"); c->opt_ramp_start = value; }
    else if (!strcmp(key, "chunk_steps")) { if (value < 1 || value > 1024) return fail(SDDE_ERR_BAD_DESC, "chunk_steps out of range"); c->opt_chunk_steps = value; }
    else return fail(SDDE_ERR_BAD_DESC, "unknown option '%s'", key);
    return SDDE_OK;
}

int sdde_b200_schedule(const sdde_desc *d, sdde_schedule *o) {
    if (!o) return fail(SDDE_ERR_BAD_DESC, "output schedule is NULL");
    memset(o, 0, sizeof(*o));
    int rc = check_desc(d);
    if (rc) return rc;
    sdde::Schedule S;
    rc = make_schedule(d, S);
    if (rc) return rc;
    auto dupd = [](const std::vector<double> &v) { double *p = (double *)malloc(sizeof(double) * std::max<size_t>(1, v.size())); if (p && !v.empty()) memcpy(p, v.data(), sizeof(double) * v.size()); return p; };
    auto dupl = [](const std::vector<int64_t> &v) { int64_t *p = (int64_t *)malloc(sizeof(int64_t) * std::max<size_t>(1, v.size())); if (p && !v.empty()) memcpy(p, v.data(), sizeof(int64_t) * v.size()); return p; };
    auto dupi = [](const std::vector<int32_t> &v) { int32_t *p = (int32_t *)malloc(sizeof(int32_t) * std::max<size_t>(1, v.size())); if (p && !v.empty()) memcpy(p, v.data(), sizeof(int32_t) * v.size()); return p; };
    o->n_steps = S.n_steps();
    o->n_save = (int64_t)S.save_t.size();
    o->n_lag = S.n_lag;
    o->n_stage = S.n_stage;
    o->ring_slots = S.ring_slots;
    o->min_gap = S.min_gap;
    o->n_tstops_hit = (int32_t)S.tstops_hit.size();
    o->t = dupd(S.t); o->dt = dupd(S.dt); o->sqdt = dupd(S.sqdt);
    o->lk_mode = dupi(S.lk_mode); o->lk_idx = dupl(S.lk_idx); o->lk_theta = dupd(S.lk_theta);
    o->save_t = dupd(S.save_t); o->save_step = dupl(S.save_step); o->save_theta = dupd(S.save_theta);
    o->save_exact = dupi(S.save_exact); o->tstops_hit = dupd(S.tstops_hit);
    if (!o->t || !o->dt || !o->sqdt || !o->lk_mode || !o->lk_idx || !o->lk_theta || !o->save_t || !o->save_step ||
        !o->save_theta || !o->save_exact || !o->tstops_hit) {
        sdde_b200_schedule_free(o);
        return fail(SDDE_ERR_NOMEM, "out of host memory");
    }
    return SDDE_OK;
}

void sdde_b200_schedule_free(sdde_schedule *s) {
    if (!s) return;
    free(s->t); free(s->dt); free(s->sqdt); free(s->lk_mode); free(s->lk_idx); free(s->lk_theta);
    free(s->save_t); free(s->save_step); free(s->save_theta); free(s->save_exact); free(s->tstops_hit);
    memset(s, 0, sizeof(*s));
}

int sdde_b200_sra_tableau(int32_t alg, double *out, int32_t *stages) {
    if (!out || !stages) return fail(SDDE_ERR_BAD_DESC, "NULL output");
    if (alg != SDDE_ALG_SRA1 && alg != SDDE_ALG_SRA3 && alg != SDDE_ALG_SOSRA) return fail(SDDE_ERR_BAD_DESC, "not a built-in SRA algorithm");
    sdde::SraHost t;
    const std::string err = sdde::sra_resolve(alg, nullptr, 0, t);
    if (!err.empty()) return fail(SDDE_ERR_BAD_DESC, "%s", err.c_str());
    const int s = t.s;
    double *p = out;
    for (int i = 0; i < s; ++i) *p++ = t.c0[i];
    for (int i = 0; i < s; ++i) *p++ = t.c1[i];
    for (int i = 0; i < s; ++i) for (int j = 0; j < s; ++j) *p++ = t.A0[i][j];
    for (int i = 0; i < s; ++i) for (int j = 0; j < s; ++j) *p++ = t.B0[i][j];
    for (int i = 0; i < s; ++i) *p++ = t.alpha[i];
    for (int i = 0; i < s; ++i) *p++ = t.beta1[i];
    for (int i = 0; i < s; ++i) *p++ = t.beta2[i];
    *stages = s;
    return SDDE_OK;
}

int sdde_b200_merge_moments(double *a, const double *b, int64_t n) {
    if (!a || !b || n < 0) return fail(SDDE_ERR_BAD_DESC, "bad arguments");
    for (int64_t i = 0; i < n; ++i) {
        const double na = a[3 * i], ma = a[3 * i + 1], sa = a[3 * i + 2];
        const double nb = b[3 * i], mb = b[3 * i + 1], sb = b[3 * i + 2];
        const double nn = na + nb;
        if (nn <= 0) { a[3 * i] = 0; a[3 * i + 1] = 0; a[3 * i + 2] = 0; continue; }
        const double dlt = mb - ma;
        a[3 * i] = nn;
        a[3 * i + 1] = ma + dlt * (nb / nn);
        a[3 * i + 2] = sa + sb + dlt * dlt * (na * nb / nn);
    }
    return SDDE_OK;
}

int sdde_b200_moments(sdde_ctx *c, const double *paths, int32_t mem, int64_t n_rows, int64_t n_traj, int64_t pitch, double *moments) {
    if (!c || !paths || !moments || n_rows <= 0 || n_traj <= 0 || pitch < n_traj) return fail(SDDE_ERR_BAD_DESC, "bad arguments");
    if (mem != SDDE_MEM_HOST && mem != SDDE_MEM_DEVICE) return fail(SDDE_ERR_BAD_DESC, "mem must be SDDE_MEM_HOST or SDDE_MEM_DEVICE");
    std::lock_guard<std::mutex> lk(c->mu);
    CU_CHECK(cudaSetDevice(c->device));
    const double *src = paths;
    if (mem == SDDE_MEM_HOST) {
        CU_CHECK(c->qs_paths.reserve((size_t)n_rows * pitch * 8));
        CU_CHECK(cudaMemcpyAsync(c->qs_paths.p, paths, (size_t)n_rows * pitch * 8, cudaMemcpyHostToDevice, c->compute));
        src = (const double *)c->qs_paths.p;
    }
    int64_t msplit = (4 * (int64_t)c->sm_count + n_rows - 1) / n_rows;
    msplit = std::max<int64_t>(1, std::min<int64_t>(msplit, (n_traj + 16383) / 16384));
    CU_CHECK(c->mom.reserve((size_t)n_rows * 3 * 8));
    CU_CHECK(c->mom_part.reserve((size_t)msplit * n_rows * 3 * 8));
    sdde::moments_kernel<<<dim3((unsigned)n_rows, (unsigned)msplit), 256, 0, c->compute>>>(src, pitch, 0, n_traj, (double *)c->mom_part.p);
    sdde::merge_moments_kernel<<<(unsigned)((n_rows + 127) / 128), 128, 0, c->compute>>>((const double *)c->mom_part.p, n_rows * 3, (int)msplit,
                                                                                       n_rows, (double *)c->mom.p);
    CU_CHECK(cudaGetLastError());
    CU_CHECK(cudaMemcpyAsync(moments, c->mom.p, (size_t)n_rows * 3 * 8, cudaMemcpyDeviceToHost, c->compute));
    CU_CHECK(cudaStreamSynchronize(c->compute));
    return SDDE_OK;
}

static int solve_once(sdde_ctx *c, const sdde_desc *d, double *out, double *moments, uint8_t *status, sdde_timing *tm);

int sdde_b200_solve(sdde_ctx *c, const sdde_desc *d, double *out, double *moments, uint8_t *status, sdde_timing *tm) {
    int rc = solve_once(c, d, out, moments, status, tm);
    if (rc == SDDE_ERR_CUDA - (int)cudaErrorMemoryAllocation && c) {
        // The wave size was planned from a cached cudaMemGetInfo (it costs ~1.5 ms per call) and other allocators share the
        // device (torch, other contexts, this context's summary workspaces): drop our reusable workspaces, forget the cached
        // budget and plan once more from a fresh reading -- the wave shrinks instead of the solve failing.
        {
            std::lock_guard<std::mutex> lk(c->mu);
            cudaSetDevice(c->device);
            cudaGetLastError();
            c->ring.release(); c->out[0].release(); c->out[1].release(); c->noise.release(); c->noise_src.release();
            c->mem_key.clear(); c->mem_free_cached = 0;
        }
        rc = solve_once(c, d, out, moments, status, tm);
    }
    return rc;
}

static int solve_once(sdde_ctx *c, const sdde_desc *d, double *out, double *moments, uint8_t *status, sdde_timing *tm) {
    const auto wall0 = std::chrono::steady_clock::now();
    static const bool dbg = getenv("SDDE_DEBUG_TIMING") != nullptr;
    auto lap = [&](const char *what) { if (dbg) fprintf(stderr, "[sdde timing] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count()); };
    if (tm) memset(tm, 0, sizeof(*tm));
    if (!c) return fail(SDDE_ERR_CUDA, "NULL context (no CUDA device? libsdde_b200 has no CPU fallback)");
    int rc = check_desc(d);
    if (rc) return rc;
    if (d->want_paths && !out) return fail(SDDE_ERR_BAD_DESC, "want_paths set but out is NULL");
    if (d->want_moments && !moments) return fail(SDDE_ERR_BAD_DESC, "want_moments set but moments is NULL");
    std::lock_guard<std::mutex> lk(c->mu);
    CU_CHECK(cudaSetDevice(c->device));
    cudaGetLastError();        // a failed earlier call must not leak its (non-sticky) error into this one

    sdde::KernelInfo K;
    rc = select_kernel(c, d, K);
    if (rc) return rc;
    const std::string skey = schedule_key(d);
    if (skey != c->sched_key) {
        c->sched_key.clear(); c->table_key.clear();
        rc = make_schedule(d, c->sched);
        if (rc) return rc;
        c->sched_key = skey;
    }
    const sdde::Schedule &S = c->sched;

    const int64_t N = S.n_steps();
    const int64_t n_save = (int64_t)S.save_t.size();
    const int NS = K.n_state, NC = K.n_chan;
    const int NSAVED = d->n_save_idxs > 0 ? d->n_save_idxs : NS;
    const int64_t rows = n_save * NSAVED;
    const int64_t NT = d->n_traj;
    const int64_t OP = d->out_pitch > 0 ? d->out_pitch : NT;      // row pitch of the caller's out array
    if (d->out_pitch != 0 && d->out_pitch < NT) return fail(SDDE_ERR_BAD_DESC, "out_pitch < n_traj");
    if (N >= (int64_t)1 << 30) return fail(SDDE_ERR_UNSUPPORTED, "more than 2^30 steps");
    if (d->out_rows > 0 && rows > d->out_rows && (d->want_paths || d->want_moments))
        return fail(SDDE_ERR_BAD_DESC, "this solve saves %lld rows (%lld save points x %d states) but the caller's buffers hold %lld "
                    "(sdde_desc.out_rows): size them from sdde_b200_schedule", (long long)rows, (long long)n_save, NSAVED, (long long)d->out_rows);
    if (tm) {
        tm->n_steps = N; tm->n_save = n_save; tm->ring_slots = S.ring_slots; tm->block_threads = K.block;
        tm->traj_per_thread = K.tpt;
        if (K.lanes > 1) snprintf(tm->kernel_name, sizeof(tm->kernel_name), "sdde_coop_kernel<%s,alg%d,noise%d,lanes%d>", K.name, d->alg, d->noise_mode, K.lanes);
        else snprintf(tm->kernel_name, sizeof(tm->kernel_name), "sdde_solve_kernel<%s,alg%d,noise%d,ptraj%d,tpt%d,B%d>", K.name,
                 d->alg, d->noise_mode, d->param_stride != 0, K.tpt, K.batch);
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, K.fn) == cudaSuccess) tm->regs_per_thread = fa.numRegs;   // also accepts a cudaKernel_t
        else cudaGetLastError();
    }
    if (moments) for (int64_t i = 0; i < rows * 3; ++i) moments[i] = 0.0;
    if (NT == 0) {
        if (tm) tm->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
        return SDDE_OK;
    }

    // ---- schedule tables -> device ----
    int64_t h2d = 0, d2h = 0;

    // ---- wave geometry ----
    const int64_t per_block = (int64_t)K.block * K.tpt / K.lanes;
    size_t free_b = 0, total_b = 0;
    {
        // the memory plan depends on the problem's SIZES only (not on the schedule's contents): cudaMemGetInfo stalls for
        // up to ~16 ms while copy engines are busy (measured), so it is asked again only when the sizes change
        std::string mkey;
        const int64_t mk[] = {NT, (int64_t)d->out_mem, (int64_t)d->want_paths, (int64_t)d->want_moments, (int64_t)d->noise_mode, (int64_t)NS,
                              (int64_t)S.ring_slots, rows, N, (int64_t)K.n_param, (int64_t)d->n_noise_t, (int64_t)(d->u0_stride != 0),
                              (int64_t)(d->param_stride != 0), (int64_t)d->noise_mem};
        mkey.append((const char *)mk, sizeof(mk));
        if (mkey != c->mem_key || c->mem_free_cached == 0) {
            CU_CHECK(cudaMemGetInfo(&free_b, &total_b));
            free_b += c->ring.cap + c->out[0].cap + c->out[1].cap + c->noise.cap;   // our own cached workspaces are reusable
            c->mem_free_cached = free_b;
            c->mem_key = mkey;
        }
        free_b = c->mem_free_cached;
        if (c->opt_mem_limit > 0) free_b = std::min<size_t>(free_b, (size_t)((double)c->opt_mem_limit / 0.80));
    }
    lap("memory budget known");
    // Host output in page-locked (mapped) memory and no moment pass over it: the kernel's own 128-bit stores cross PCIe,
    // no staging buffers, no copy engine, no first-wave latency (C2: 77 ms instead of 81 ms per 10^6-path solve).
    double *out_mapped = nullptr;
    int zc_class = 0;
    int zc_trial = -1;             // auto mode on a large output: which arm this solve runs (0 staged, 1 zero-copy), timed below
    if (d->want_paths && d->out_mem == SDDE_MEM_HOST && !d->want_moments && c->opt_zero_copy != 0 && c->opt_wave_traj <= 0 && rows > 0) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, out) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
            out_mapped = (double *)at.devicePointer;
        else
            cudaGetLastError();
        if (out_mapped && c->opt_zero_copy < 0 && (double)rows * (double)NT * 8.0 >= 64.0 * 1048576.0) {
            // auto: the first large solve of a size class runs staged, the second zero-copy, later ones the faster of the two
            zc_class = 0;
            for (double b = (double)rows * (double)NT * 8.0; b >= 2.0 && zc_class < 63; b *= 0.5) ++zc_class;
            const sdde_ctx::ZcTrial &z = c->zc[zc_class];
            zc_trial = z.n[0] == 0 ? 0 : (z.n[1] == 0 ? 1 : (z.ms_per_gb[1] < z.ms_per_gb[0] ? 1 : 0));
            if (zc_trial == 0) out_mapped = nullptr;
        }
    }
    const bool zero_copy = out_mapped != nullptr;
    if (zero_copy && tm) strncat(tm->kernel_name, " -> mapped host out", sizeof(tm->kernel_name) - strlen(tm->kernel_name) - 1);
    const bool out_dev = (d->want_paths && d->out_mem == SDDE_MEM_DEVICE) || zero_copy;
    const bool need_stage = (d->want_paths && !out_dev) || (!d->want_paths && d->want_moments);
    const bool noise_host = d->noise_mode != SDDE_NOISE_PHILOX && d->noise_mem == SDDE_MEM_HOST;
    const int64_t noise_rows = (d->noise_mode == SDDE_NOISE_PHILOX) ? 0 : (N + (d->noise_mode == SDDE_NOISE_SUPPLIED_GRID ? 1 : 0)) * NC;
    // NoiseGrid on its own time grid: W is interpolated onto the schedule's nodes on the device first (noise_regrid_kernel)
    const bool regrid = d->noise_mode == SDDE_NOISE_SUPPLIED_GRID && d->noise_t != nullptr;
    const int64_t src_rows = regrid ? d->n_noise_t * NC : noise_rows;
    std::vector<sdde::GridNode> grid_nodes;
    if (regrid) {
        const std::string gerr = sdde::plan_noise_grid(S.t, d->noise_t, d->n_noise_t, grid_nodes);
        if (!gerr.empty()) return fail(SDDE_ERR_BAD_DESC, "%s", gerr.c_str());
    }
    const double bytes_per_traj = 8.0 * ((double)S.ring_slots * NS + (need_stage ? 2.0 * rows : 0.0) + (noise_host ? (double)src_rows : 0.0) +
                                         (regrid ? (double)noise_rows : 0.0) + NS + K.n_param) + 1.0;
    int64_t wave = c->opt_wave_traj > 0 ? c->opt_wave_traj : ((int64_t)1 << 20);
    if (c->opt_wave_traj <= 0 && d->want_paths && !out_dev && NT > ((int64_t)1 << 19)) wave = (int64_t)1 << 19;  // pipeline D2H (PCIe-bound: few, wide waves)
    const int64_t mem_cap = (int64_t)((double)free_b * 0.80 / bytes_per_traj);
    // a wave is a whole number of warps (wq trajectories); the ring / staging pitch is rounded up to whole CTAs, so every
    // thread of the grid owns in-bounds columns (threads past the wave's last trajectory recompute it and store nothing)
    const int64_t wq = (int64_t)128 * K.tpt / K.lanes;
    wave = std::min(wave, std::max<int64_t>(wq, mem_cap));
    wave = std::max<int64_t>(wq, (wave / wq) * wq);
    wave = std::min(wave, ((NT + wq - 1) / wq) * wq);
    // (+ one CTA: the tail round of the balanced grid may overshoot the wave by less than one tail CTA, see the launch)
    const int64_t pitch = ((wave + per_block - 1) / per_block) * per_block + per_block;
    // Wave list.  When the saved paths leave over PCIe (host output) the first waves are small and grow geometrically,
    // so that the D2H stream starts after a few milliseconds instead of after a whole 2^19-trajectory wave
    // and then never idles (measured on C2: D2H engine busy from 4 ms to the end).
    std::vector<int64_t> wave_off, wave_cnt;
    {
        const bool ramp = c->opt_wave_traj <= 0 && d->want_paths && !out_dev && NT > wave;
        int64_t cur = ramp ? std::max<int64_t>(wq, (c->opt_ramp_start / wq) * wq) : wave;
        for (int64_t off = 0; off < NT;) {
            const int64_t cnt = std::min(std::min(cur, wave), NT - off);
            wave_off.push_back(off); wave_cnt.push_back(cnt);
            off += cnt;
            if (ramp) cur = std::min(wave, std::max<int64_t>(wq, (cur * 3 / 2 / wq) * wq));   // x1.5: a wave's kernel must end before the previous wave's copy does (kernel ~25, PCIe ~14 trajectories/us on C2)
        }
    }
    const int n_waves = (int)wave_off.size();

    int chunk = (int)std::min<int64_t>(c->opt_chunk_steps, std::max<int64_t>(1, N));
    if (chunk >= 8) chunk -= chunk % 8;   // batches (B <= 8) never straddle a chunk; step index of a batch start stays even
    // schedule records staged per refill (+ look-ahead records for the next step / next batch prefetch) must fit in shared
    // memory beside the generator's table (in-kernel normals only): the SM's 227 KB, less 1 KB per CTA the system reserves,
    // shared by the CTAs the kernel was compiled to keep resident
    const size_t tab_bytes = d->noise_mode == SDDE_NOISE_PHILOX ? (size_t)SDDE_RNG_SMEM_BYTES(K.lanes > 1 ? 1 : SDDE_RNG_COPIES) : 0;
    const size_t smem_cap = (size_t)(227 * 1024) / (size_t)std::max(1, K.min_ctas) - 2048;
    while (chunk > 8 && tab_bytes + (size_t)(chunk + 16) * K.rec_size > smem_cap) chunk = ((chunk / 2) / 8) * 8;
    const size_t smem = tab_bytes + (size_t)(chunk + 16) * K.rec_size;
    if (smem > 48 * 1024)      // above the default limit the kernel has to opt in (built-in kernels and NVRTC kernel handles alike)
        CU_CHECK(cudaFuncSetAttribute(K.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int batch = K.batch;
    if (c->opt_batch > 0) batch = (int)std::min<int64_t>(batch, c->opt_batch);
    if (batch > S.min_gap + 1 || chunk < 8) batch = 1;     // prefetched nodes must be older than the batch, else step-by-step
    if (batch < K.batch && K.batch > 1) batch = 1;          // the kernel runs its compiled depth or step by step

    const int64_t tpitch = pitch;
    // ---- schedule tables -> device (byte offsets depend on the ring's row pitch) ----
    std::vector<unsigned char> h_steps;
    std::vector<sdde::SaveRec> h_saves;
    {
        std::string tkey = skey;
        const int64_t tk[] = {tpitch, (int64_t)K.rec_size, (int64_t)K.n_state, (int64_t)batch, (int64_t)chunk, (int64_t)c->opt_lean};
        tkey.append((const char *)tk, sizeof(tk));
        if (tkey != c->table_key) {
            c->table_key.clear();
            pack_tables(S, K, tpitch, c->opt_lean ? batch : 0, chunk, c->opt_lean != 2, h_steps, h_saves);
            CU_CHECK(c->steps.reserve(std::max<size_t>(16, h_steps.size())));
            CU_CHECK(c->saves.reserve(std::max<size_t>(16, h_saves.size() * sizeof(sdde::SaveRec))));
            if (!h_steps.empty()) CU_CHECK(cudaMemcpyAsync(c->steps.p, h_steps.data(), h_steps.size(), cudaMemcpyHostToDevice, c->compute));
            if (!h_saves.empty()) CU_CHECK(cudaMemcpyAsync(c->saves.p, h_saves.data(), h_saves.size() * sizeof(sdde::SaveRec), cudaMemcpyHostToDevice, c->compute));
            h2d += (int64_t)h_steps.size() + (int64_t)(h_saves.size() * sizeof(sdde::SaveRec));
            CU_CHECK(cudaStreamSynchronize(c->compute));     // h_steps / h_saves are stack-owned
            c->table_key = tkey;
        }
    }

    CU_CHECK(c->ring.reserve((size_t)S.ring_slots * NS * pitch * 8));
    if (need_stage) {
        CU_CHECK(c->out[0].reserve((size_t)std::max<int64_t>(1, rows) * pitch * 8));
        if (n_waves > 1 && d->want_paths) CU_CHECK(c->out[1].reserve((size_t)std::max<int64_t>(1, rows) * pitch * 8));
    }
    DevBuf &noise_in = regrid ? c->noise_src : c->noise;      // where host-supplied rows are staged
    if (noise_host) CU_CHECK(noise_in.reserve((size_t)std::max<int64_t>(1, src_rows) * pitch * 8));
    if (regrid) {
        CU_CHECK(c->noise.reserve((size_t)std::max<int64_t>(1, noise_rows) * pitch * 8));
        CU_CHECK(c->noise_nodes.reserve(grid_nodes.size() * sizeof(sdde::GridNode)));
        CU_CHECK(cudaMemcpyAsync(c->noise_nodes.p, grid_nodes.data(), grid_nodes.size() * sizeof(sdde::GridNode), cudaMemcpyHostToDevice, c->compute));
        CU_CHECK(cudaStreamSynchronize(c->compute));
        h2d += (int64_t)(grid_nodes.size() * sizeof(sdde::GridNode));
    }
    if (d->u0_stride != 0) CU_CHECK(c->u0.reserve((size_t)NS * pitch * 8));
    if (d->param_stride != 0) CU_CHECK(c->ptraj.reserve((size_t)std::max(1, K.n_param) * pitch * 8));
    const bool track = status != nullptr || tm != nullptr;     // status bytes live on the device either way; D2H only on request
    if (track) CU_CHECK(c->status.reserve((size_t)NT + (size_t)pitch + 16));   // all waves + a 4-byte counter behind them
    if (status) CU_CHECK(c->h_status.reserve((size_t)NT));
    CU_CHECK(c->h_count.reserve(16));
    unsigned int *d_count = track ? (unsigned int *)((unsigned char *)c->status.p + (((size_t)NT + (size_t)pitch + 3) & ~(size_t)3)) : nullptr;
    if (d->want_moments) CU_CHECK(c->mom.reserve((size_t)std::max<int64_t>(1, rows) * 3 * 8 * n_waves));
    if (tm) { tm->waves = n_waves; tm->ring_bytes = (int64_t)S.ring_slots * NS * pitch * 8; }

    sdde::KArgs A;
    memset(&A, 0, sizeof(A));
    A.steps = (const unsigned char *)c->steps.p;
    A.saves = (const sdde::SaveRec *)c->saves.p;
    A.ring = (double *)c->ring.p;
    A.ring_pitch = tpitch;
    A.slot_bytes = (long long)NS * tpitch * 8;
    A.rng_tab = (const double2 *)c->rng_tab.p;
    A.ring_bytes = (long long)S.ring_slots * NS * pitch * 8;
    A.n_steps = (int)N; A.n_saves = (int)n_save; A.ring_slots = S.ring_slots; A.chunk_steps = chunk; A.batch = batch;
    A.noise_kind = d->noise_mode;
    A.n_saved = NSAVED;
    A.stage_order = (d->alg >= SDDE_ALG_WANGLI3SMIL_A && d->alg <= SDDE_ALG_WANGLI3SMIL_F) ? wangli_stage_order(d->alg) : 0;
    if (is_sra(d->alg)) {
        sdde::SraHost tab;
        sdde::sra_resolve(d->alg, d->sra_tableau, d->sra_stages, tab);     // validated by check_desc
        A.sra.s = tab.s;
        for (int i = 0; i < 3; ++i) {
            A.sra.fset[i] = tab.fset[i]; A.sra.gset[i] = tab.gset[i]; A.sra.c0[i] = tab.c0[i]; A.sra.c1[i] = tab.c1[i];
            A.sra.alpha[i] = tab.alpha[i]; A.sra.beta1[i] = tab.beta1[i]; A.sra.beta2[i] = tab.beta2[i];
            for (int j = 0; j < 3; ++j) { A.sra.A0[i][j] = tab.A0[i][j]; A.sra.B0[i][j] = tab.B0[i][j]; }
        }
    }
    A.prefetch = (S.min_gap >= 1) ? 1 : 0;
    for (int j = 0; j < SDDE_K_MAX_STATE; ++j) A.save_row[j] = d->n_save_idxs > 0 ? -1 : (j < NS ? j : -1);
    for (int i = 0; i < d->n_save_idxs; ++i) A.save_row[d->save_idxs[i]] = i;
    A.seed_lo = (sdde::u32)(d->seed & 0xffffffffu); A.seed_hi = (sdde::u32)(d->seed >> 32);
    for (int r = 0; r < 10; ++r) { A.rk[2 * r] = A.seed_lo + (sdde::u32)r * 0x9E3779B9u; A.rk[2 * r + 1] = A.seed_hi + (sdde::u32)r * 0xBB67AE85u; }
    for (int j = 0; j < NS; ++j) { A.hist_c0[j] = d->hist_c0[j]; A.hist_c1[j] = d->hist_c1[j]; }
    if (d->u0_stride == 0) for (int j = 0; j < NS; ++j) A.u0c[j] = d->u0[j];
    if (d->param_stride == 0) for (int i = 0; i < K.n_param; ++i) A.p[i] = d->params[i];

    double *h_mom = nullptr;
    if (d->want_moments) { CU_CHECK(c->h_mom.reserve((size_t)std::max<int64_t>(1, rows) * 3 * 8 * n_waves)); h_mom = (double *)c->h_mom.p; }
    int launches = 0;
    float kernel_ms = 0.f;
    CU_CHECK(cudaEventRecord(c->ev_c0, c->copy_out));
    bool first_kernel = true;

    // Host output in PAGEABLE memory (what a caller gets who did not allocate page-locked memory): a cudaMemcpy2DAsync into it is
    // staged by the driver on one thread and, for a fresh allocation, takes the first-touch page faults there too (835 ms for
    // the 4 GB of C2 against 78 ms into page-locked memory).  Here each wave's tile leaves in 64 MB chunks into two page-locked
    // bounce buffers; while chunk k+1 crosses PCIe the host threads scatter chunk k into the caller's rows.
    bool pageable_out = false;
    if (d->want_paths && !out_dev && rows > 0 && (double)rows * (double)NT * 8.0 >= (double)c->opt_bounce_min_mb * 1048576.0 && c->opt_host_threads != 1) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, out) == cudaSuccess) pageable_out = at.type == cudaMemoryTypeUnregistered;
        else cudaGetLastError();
    }
    const size_t kBounce = (size_t)64 << 20;
    if (pageable_out) {
        CU_CHECK(c->bounce[0].reserve(kBounce));
        CU_CHECK(c->bounce[1].reserve(kBounce));
        for (cudaEvent_t &e : c->ev_bounce) if (!e) CU_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        if (!c->pool) {
            const int hw = (int)std::thread::hardware_concurrency();
            const int nthr = c->opt_host_threads > 0 ? (int)c->opt_host_threads : std::max(2, std::min(16, hw > 0 ? hw : 8));
            c->pool.reset(new sdde::HostPool(nthr - 1));
        }
    }
    // wave w's staging tile (rows x nt doubles, row pitch `pitch`) -> out[:, off : off + nt], blocking the calling thread; the GPU
    // meanwhile runs the next wave's kernel, which was enqueued before this call
    auto drain_pageable = [&](const double *tile, int64_t off, int64_t nt, int buf) -> int {
        CU_CHECK(cudaStreamWaitEvent(c->copy_out, c->ev_kdone[buf], 0));
        const int64_t row_bytes = nt * 8;
        const int64_t rpc = std::max<int64_t>(1, std::min<int64_t>(rows, (int64_t)kBounce / row_bytes));   // rows per chunk
        const int64_t n_seg = (row_bytes + kBounce - 1) / (int64_t)kBounce;                                 // a row longer than a buffer is cut
        const int64_t seg_elems = (nt + n_seg - 1) / n_seg;
        struct Chunk { int64_t r0, nr, e0, ne; };
        std::vector<Chunk> chunks;
        for (int64_t r0 = 0; r0 < rows; r0 += rpc)
            for (int64_t sg = 0; sg < n_seg; ++sg) {
                const int64_t e0 = sg * seg_elems, ne = std::min(seg_elems, nt - e0);
                if (ne > 0) chunks.push_back(Chunk{r0, std::min(rpc, rows - r0), e0, ne});
            }
        auto issue = [&](size_t k) -> cudaError_t {
            const Chunk &ch = chunks[k];
            cudaError_t e = cudaMemcpy2DAsync(c->bounce[k & 1].p, (size_t)ch.ne * 8, tile + ch.r0 * pitch + ch.e0, (size_t)pitch * 8, (size_t)ch.ne * 8,
                                              (size_t)ch.nr, cudaMemcpyDeviceToHost, c->copy_out);
            if (e == cudaSuccess) e = cudaEventRecord(c->ev_bounce[k & 1], c->copy_out);
            return e;
        };
        if (!chunks.empty()) CU_CHECK(issue(0));
        for (size_t k = 0; k < chunks.size(); ++k) {
            if (k + 1 < chunks.size()) CU_CHECK(issue(k + 1));            // the other buffer: its previous contents were scattered at k - 1
            CU_CHECK(cudaEventSynchronize(c->ev_bounce[k & 1]));
            const Chunk ch = chunks[k];
            const double *src = (const double *)c->bounce[k & 1].p;
            const int64_t piece = (int64_t)1 << 17;                        // 1 MB pieces: enough tasks for every thread, whole pages
            const int64_t ppr = (ch.ne + piece - 1) / piece;
            std::function<void(int64_t)> fn = [&](int64_t task) {
                const int64_t r = task / ppr, p0 = (task % ppr) * piece, pn = std::min(piece, ch.ne - p0);
                memcpy(out + (ch.r0 + r) * OP + off + ch.e0 + p0, src + r * ch.ne + p0, (size_t)pn * 8);
            };
            c->pool->parallel_for(ch.nr * ppr, fn);
        }
        return SDDE_OK;
    };
    int pending_w = -1;                              // pageable path: wave whose tile still has to be drained

    std::vector<cudaEvent_t> dbg_ev;                 // SDDE_DEBUG_TIMING: per wave (kernel start, kernel end, copy start, copy end)
    auto dbg_mark = [&](cudaStream_t s) { if (dbg) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); dbg_ev.push_back(e); } };
    for (int w = 0; w < n_waves; ++w) {
        const int64_t off = wave_off[(size_t)w];
        const int64_t nt = wave_cnt[(size_t)w];
        const int buf = w & 1;
        // inputs of this wave
        if (d->u0_stride != 0 || d->param_stride != 0 || noise_host) {
            if (w > 0) CU_CHECK(cudaStreamSynchronize(c->compute));   // single input buffer: previous wave must be done
            if (d->u0_stride != 0) {
                CU_CHECK(cudaMemcpy2DAsync(c->u0.p, pitch * 8, d->u0 + off, d->u0_stride * 8, nt * 8, NS, cudaMemcpyHostToDevice, c->copy_in));
                h2d += nt * 8 * NS;
            }
            if (d->param_stride != 0 && K.n_param > 0) {
                CU_CHECK(cudaMemcpy2DAsync(c->ptraj.p, pitch * 8, d->params + off, d->param_stride * 8, nt * 8, K.n_param, cudaMemcpyHostToDevice, c->copy_in));
                h2d += nt * 8 * K.n_param;
            }
            if (noise_host) {
                CU_CHECK(cudaMemcpy2DAsync(noise_in.p, pitch * 8, d->noise + off, NT * 8, nt * 8, src_rows, cudaMemcpyHostToDevice, c->copy_in));
                h2d += nt * 8 * src_rows;
            }
            CU_CHECK(cudaEventRecord(c->ev_in, c->copy_in));
            CU_CHECK(cudaStreamWaitEvent(c->compute, c->ev_in, 0));
        }
        A.u0 = d->u0_stride != 0 ? (const double *)c->u0.p : nullptr; A.u0_pitch = pitch;
        A.ptraj = d->param_stride != 0 ? (const double *)c->ptraj.p : nullptr; A.p_pitch = pitch;
        if (d->noise_mode != SDDE_NOISE_PHILOX) {
            if (noise_host) { A.noise = (const double *)noise_in.p; A.noise_pitch = pitch; A.noise_off = 0; }
            else { A.noise = d->noise; A.noise_pitch = NT; A.noise_off = off; }
            if (regrid) {
                const unsigned gx = (unsigned)std::min<int64_t>((nt + 255) / 256, 1024);
                sdde::noise_regrid_kernel<<<dim3((unsigned)noise_rows, gx), 256, 0, c->compute>>>(
                    A.noise, A.noise_pitch, A.noise_off, (const sdde::GridNode *)c->noise_nodes.p, NC, nt, (double *)c->noise.p, pitch);
                CU_CHECK(cudaGetLastError());
                ++launches;
                A.noise = (const double *)c->noise.p; A.noise_pitch = pitch; A.noise_off = 0;
            }
        }
        A.status = track ? (unsigned char *)c->status.p + off : nullptr;
        A.n_traj = nt;
        A.traj_gid0 = d->traj_offset + off;
        if (out_dev) {
            A.out = zero_copy ? out_mapped : out; A.out_pitch = OP; A.out_off = off; A.out_bytes = (long long)rows * OP * 8;
            A.out_vec_ok = ((OP % 4) == 0 && (off % 4) == 0 && ((uintptr_t)A.out % 16) == 0) ? 1 : 0;
            if (zero_copy) d2h += nt * 8 * rows;
        } else if (need_stage) {
            const int b = (d->want_paths && n_waves > 1) ? buf : 0;
            if (d->want_paths && w >= 2) CU_CHECK(cudaStreamWaitEvent(c->compute, c->ev_copied[b], 0));  // staging buffer free again
            A.out = (double *)c->out[b].p; A.out_pitch = pitch; A.out_off = 0; A.out_vec_ok = 1; A.out_bytes = (long long)rows * pitch * 8;
        } else {
            A.out = nullptr;
        }
        // grid: whole rounds of full CTAs (one round = every SM's resident CTAs), then the remainder spread evenly over one
        // last round of smaller CTAs (whole warps each) -- see the kernel
        int64_t n_full = (nt + per_block - 1) / per_block, tail = 0, n_tail = 0;
        if (c->opt_balance && K.lanes == 1) {
            const int64_t round_ctas = (int64_t)c->sm_count * std::max(1, K.min_ctas), wq_t = (int64_t)32 * K.tpt;
            n_full = (nt / (round_ctas * per_block)) * round_ctas;
            const int64_t rem = nt - n_full * per_block;
            if (rem > 0) {
                tail = std::min<int64_t>(per_block, (((rem + round_ctas - 1) / round_ctas + wq_t - 1) / wq_t) * wq_t);
                n_tail = (rem + tail - 1) / tail;
            }
        }
        A.n_full_ctas = (int)n_full; A.tail_traj = (int)tail;
        const unsigned grid = (unsigned)(n_full + n_tail);
        if (first_kernel) { CU_CHECK(cudaEventRecord(c->ev_k0, c->compute)); first_kernel = false; }
        dbg_mark(c->compute);
        if (K.is_user) {
            std::string err;
            if (!c->user.launch(K, grid, (unsigned)K.block, smem, c->compute, &A, sizeof(A), err)) return fail(SDDE_ERR_CUDA, "%s", err.c_str());
        } else {
            void *args[] = {(void *)&A};
            CU_CHECK(cudaLaunchKernel(K.fn, dim3(grid), dim3((unsigned)K.block), args, smem, c->compute));
        }
        ++launches;
        if (d->want_moments && rows > 0) {
            double *dm = (double *)c->mom.p + (size_t)w * rows * 3;
            // enough slices per row to fill the machine, none shorter than 16384 values
            int64_t msplit = (4 * (int64_t)c->sm_count + rows - 1) / rows;
            msplit = std::max<int64_t>(1, std::min<int64_t>(msplit, (nt + 16383) / 16384));
            if (msplit == 1) {
                sdde::moments_kernel<<<dim3((unsigned)rows, 1), 256, 0, c->compute>>>(A.out, A.out_pitch, A.out_off, nt, dm);
            } else {
                CU_CHECK(c->mom_part.reserve((size_t)msplit * rows * 3 * 8));
                sdde::moments_kernel<<<dim3((unsigned)rows, (unsigned)msplit), 256, 0, c->compute>>>(A.out, A.out_pitch, A.out_off, nt,
                                                                                                   (double *)c->mom_part.p);
                sdde::merge_moments_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, c->compute>>>(
                    (const double *)c->mom_part.p, rows * 3, (int)msplit, rows, dm);
                ++launches;
            }
            CU_CHECK(cudaGetLastError());
            ++launches;
        }
        CU_CHECK(cudaEventRecord(c->ev_kdone[buf], c->compute));
        dbg_mark(c->compute);
        // outputs of this wave
        if (d->want_paths && !out_dev && rows > 0 && pageable_out) {
            // this wave's kernel is enqueued: drain the PREVIOUS wave's tile on the host while it runs
            if (pending_w >= 0) {
                rc = drain_pageable((const double *)c->out[(n_waves > 1) ? (pending_w & 1) : 0].p, wave_off[(size_t)pending_w],
                                    wave_cnt[(size_t)pending_w], pending_w & 1);
                if (rc) return rc;
            }
            pending_w = w;
            d2h += nt * 8 * rows;
        } else if (d->want_paths && !out_dev && rows > 0) {
            CU_CHECK(cudaStreamWaitEvent(c->copy_out, c->ev_kdone[buf], 0));
            dbg_mark(c->copy_out);
            CU_CHECK(cudaMemcpy2DAsync(out + off, OP * 8, A.out, pitch * 8, nt * 8, rows, cudaMemcpyDeviceToHost, c->copy_out));
            CU_CHECK(cudaEventRecord(c->ev_copied[buf], c->copy_out));
            dbg_mark(c->copy_out);
            d2h += nt * 8 * rows;
        }
    }
    if (pending_w >= 0) {
        rc = drain_pageable((const double *)c->out[(n_waves > 1) ? (pending_w & 1) : 0].p, wave_off[(size_t)pending_w], wave_cnt[(size_t)pending_w],
                            pending_w & 1);
        if (rc) return rc;
    }
    CU_CHECK(cudaEventRecord(c->ev_k1, c->compute));
    // Status bytes leave in ONE copy after the last wave: a small D2H per wave on the compute stream would queue behind the
    // wave's path copy in the D2H copy engine and stall the next wave's kernel until that copy ends (no overlap at all).
    if (track) {
        CU_CHECK(cudaMemsetAsync(d_count, 0, 4, c->compute));
        sdde::count_status_kernel<<<(unsigned)std::min<int64_t>((NT + 255) / 256, 4 * (int64_t)c->sm_count), 256, 0, c->compute>>>(
            (const unsigned char *)c->status.p, NT, d_count);
        CU_CHECK(cudaGetLastError());
        ++launches;
        CU_CHECK(cudaMemcpyAsync(c->h_count.p, d_count, 4, cudaMemcpyDeviceToHost, c->compute));
        d2h += 4;
    }
    if (status) {
        CU_CHECK(cudaMemcpyAsync(c->h_status.p, c->status.p, (size_t)NT, cudaMemcpyDeviceToHost, c->compute));
        d2h += NT;
    }
    if (d->want_moments && rows > 0 && c->keep_moments_on_device) {
        // multi-GPU path: wave merge on the device, result stays in mom[0..rows) for the NCCL all-gather
        if (n_waves > 1) {
            sdde::merge_moments_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, c->compute>>>(
                (const double *)c->mom.p, rows * 3, n_waves, rows, (double *)c->mom.p);
            CU_CHECK(cudaGetLastError());
            ++launches;
        }
    } else if (d->want_moments && rows > 0) {
        CU_CHECK(cudaMemcpyAsync(h_mom, c->mom.p, (size_t)rows * 3 * 8 * n_waves, cudaMemcpyDeviceToHost, c->compute));
        d2h += (int64_t)rows * 3 * 8 * n_waves;
    }
    CU_CHECK(cudaEventRecord(c->ev_c1, c->copy_out));
    lap("all work enqueued");
    CU_CHECK(cudaStreamSynchronize(c->compute));
    lap("compute stream synced");
    CU_CHECK(cudaStreamSynchronize(c->copy_out));
    CU_CHECK(cudaStreamSynchronize(c->copy_in));
    lap("copy streams synced");
    CU_CHECK(cudaGetLastError());
    if (dbg && !dbg_ev.empty()) {
        const int per = (d->want_paths && !out_dev && rows > 0 && !pageable_out) ? 4 : 2;
        for (size_t i = 0; i + per <= dbg_ev.size(); i += per) {
            float t[4] = {0, 0, 0, 0};
            for (int j = 0; j < per; ++j) cudaEventElapsedTime(&t[j], dbg_ev[0], dbg_ev[i + j]);
            fprintf(stderr, "[sdde timing] wave %zu (%lld traj): kernel %.2f..%.2f ms  copy %.2f..%.2f ms\n", i / per,
                    (long long)wave_cnt[i / per], t[0], t[1], t[2], t[3]);
        }
        for (cudaEvent_t e : dbg_ev) cudaEventDestroy(e);
    }
    if (zc_trial >= 0 && !first_kernel) {
        float a = 0.f, b = 0.f;
        if (cudaEventElapsedTime(&a, c->ev_k0, c->ev_k1) == cudaSuccess) {
            if (zc_trial == 0 && cudaEventElapsedTime(&b, c->ev_k0, c->ev_c1) == cudaSuccess) a = std::max(a, b);
            sdde_ctx::ZcTrial &z = c->zc[zc_class];
            const double per_gb = (double)a / ((double)rows * (double)NT * 8.0e-9);
            z.ms_per_gb[zc_trial] = z.n[zc_trial] ? std::min(z.ms_per_gb[zc_trial], per_gb) : per_gb;
            ++z.n[zc_trial];
        }
        cudaGetLastError();
    }
    if (status) memcpy(status, c->h_status.p, (size_t)NT);
    if (d->want_moments && rows > 0 && !c->keep_moments_on_device) {
        memcpy(moments, h_mom, (size_t)rows * 3 * 8);
        for (int w = 1; w < n_waves; ++w) sdde_b200_merge_moments(moments, h_mom + (size_t)w * rows * 3, rows);
    }
    if (tm) {
        float ms = 0.f;
        if (!first_kernel && cudaEventElapsedTime(&ms, c->ev_k0, c->ev_k1) == cudaSuccess) kernel_ms = ms;
        tm->kernel_ms = kernel_ms;
        if (cudaEventElapsedTime(&ms, c->ev_c0, c->ev_c1) == cudaSuccess) tm->d2h_ms = ms;
        tm->h2d_bytes = h2d; tm->d2h_bytes = d2h; tm->launches = launches;
        tm->n_unstable = (int32_t)std::min<unsigned int>(*(const unsigned int *)c->h_count.p, 0x7fffffffu);
        tm->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    }
    return SDDE_OK;
}

int sdde_b200_compile_model(sdde_ctx *c, const char *source, int32_t *model_id) {
    if (!c || !source || !model_id) return fail(SDDE_ERR_BAD_DESC, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CU_CHECK(cudaSetDevice(c->device));
    std::string err;
    int id = c->user.compile(source, err);
    if (id < 0) return fail(SDDE_ERR_COMPILE, "%s", err.c_str());
    *model_id = id;
    return SDDE_OK;
}

}  // extern "C"
