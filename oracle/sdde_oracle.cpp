// sdde_oracle.cpp -- CPU restatement (C++) of the reference's per-trajectory solve.  TEST INFRASTRUCTURE ONLY.
//
// PARITY UNPINNED: like oracle/sdde_oracle.py this follows the reference's algorithm (file:line below), but the
// reference itself (Julia) cannot be run in this environment and ships no golden vectors for this path.
//
// This is the "CPU restatement stand-in" of BASELINE.md section 2: one trajectory at a time, the way
// StochasticDelayDiffEq.jl does it, parallelised over trajectories with std::thread in static chunks (what
// [EXT] SciMLBase EnsembleThreads' Threads.@threads does).  Per trajectory it deliberately keeps the reference's
// work: its own tstop / discontinuity / saveat heaps (src/solve.jl:613-713), a growing dense history
// (sol.t / sol.u push per step, src/integrators/interface.jl:406-417), a binary search + linear interpolation per
// delayed-state lookup PER f/g EVALUATION ([EXT] HistoryFunction -> sde_interpolation, wired at src/utils.jl:246;
// f and g each call h(p, t - tau), src/functionwrapper.jl:8-9,35-36) and one normal draw per step and channel
// (src/integrators/interface.jl:449-454).  It has none of Julia's __init / deepcopy / dynamic-dispatch overhead, so
// speed-ups quoted against it are conservative.
//
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may load this library.
// Build: make -C oracle   (g++ -O2 -ffp-contract=off; no FMA so it is bit-identical to the numpy oracle).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <queue>
#include <thread>
#include <vector>

#include "../include/sdde_b200.h"

namespace {

inline double jl_eps(double x) {
    x = std::fabs(x);
    return std::nextafter(x, std::numeric_limits<double>::infinity()) - x;
}

// ---- Philox4x32-10 + Box-Muller: same stream definition as the GPU kernels (csrc/sdde_kernels.cuh) ----
inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t o[4]) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
// same definition as oracle/sdde_oracle.py::_box_muller_pair and csrc/sdde_kernels.cuh::normal_pair_from_words
inline void box_muller_pair(uint32_t wa, uint32_t wb, double &z0, double &z1) {
    const uint64_t m = ((uint64_t)(wb & 0xFFu) << 32) | wa;
    const double u = ((double)m + 0.5) * 0x1p-40;
    const double r = std::sqrt(-2.0 * std::log(u));
    const unsigned q = wb >> 30;
    const uint32_t f = (wb >> 8) & 0x3FFFFFu;
    const double x = ((double)f + 0.5) * 0x1p-22 - 0.5;
    const double s = std::sin((M_PI / 2) * x), c = std::cos((M_PI / 2) * x);
    z0 = r * (q == 0 ? c : (q == 1 ? -s : (q == 2 ? -c : s)));
    z1 = r * (q == 0 ? s : (q == 1 ? c : (q == 2 ? -s : -c)));
}
inline void normal_quad(uint64_t gid, uint64_t call, uint64_t seed, double z[4]) {
    uint32_t w[4];
    philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)call, (uint32_t)(call >> 32), (uint32_t)seed,
                  (uint32_t)(seed >> 32), w);
    box_muller_pair(w[0], w[1], z[0], z[1]);
    box_muller_pair(w[2], w[3], z[2], z[3]);
}

// ---- TIMING ONLY: a normal generator of the reference's cost class (noise_mode = SDDE_ORACLE_NOISE_ZIGGURAT) ----
// The reference draws its increments with Xoroshiro128Plus + Julia's randn (a 256-layer ziggurat on 52 random bits), /root/
// reference/src/solve.jl:313-336; the Philox + libm Box-Muller above is several times dearer per normal.  bench.py's CPU legs
// report the port's throughput with BOTH.  Doornik's ZIGNOR layout (128 layers, r = 3.442619855899, v = 9.91256303526217e-3),
// stream = xoroshiro128+ seeded per trajectory by splitmix64(seed, trajectory id).  Not used by any parity test.
#define SDDE_ORACLE_NOISE_ZIGGURAT 100
struct Zig {
    double x[129], ratio[128];
    Zig() {
        const double r = 3.442619855899, v = 9.91256303526217e-3;
        double f = std::exp(-0.5 * r * r);
        x[0] = v / f; x[1] = r; x[128] = 0.0;
        for (int i = 2; i < 128; ++i) { x[i] = std::sqrt(-2.0 * std::log(v / x[i - 1] + f)); f = std::exp(-0.5 * x[i] * x[i]); }
        for (int i = 0; i < 128; ++i) ratio[i] = x[i + 1] / x[i];
    }
};
static const Zig g_zig;
struct Xoro {
    uint64_t s0, s1;
    static uint64_t splitmix(uint64_t &z) { z += 0x9E3779B97F4A7C15ull; uint64_t r = z; r = (r ^ (r >> 30)) * 0xBF58476D1CE4E5B9ull; r = (r ^ (r >> 27)) * 0x94D049BB133111EBull; return r ^ (r >> 31); }
    Xoro(uint64_t seed, uint64_t id) { uint64_t z = seed ^ (id * 0xD1342543DE82EF95ull); s0 = splitmix(z); s1 = splitmix(z); if (!(s0 | s1)) s1 = 1; }
    inline uint64_t next() { const uint64_t a = s0; uint64_t b = s1; const uint64_t r = a + b; b ^= a; s0 = ((a << 24) | (a >> 40)) ^ b ^ (b << 16); s1 = (b << 37) | (b >> 27); return r; }
    inline double u01() { return (double)(next() >> 11) * 0x1.0p-53; }
};
inline double zig_normal(Xoro &g) {
    for (;;) {
        const uint64_t w = g.next();
        const int i = (int)(w & 127);
        const double u = 2.0 * ((double)(w >> 11) * 0x1.0p-53) - 1.0;
        if (std::fabs(u) < g_zig.ratio[i]) return u * g_zig.x[i];
        if (i == 0) {                                   // tail beyond r
            const double r = 3.442619855899;
            double xx, yy;
            do { xx = std::log(1.0 - g.u01()) / r; yy = std::log(1.0 - g.u01()); } while (-2.0 * yy < xx * xx);
            return u < 0 ? xx - r : r - xx;
        }
        const double xv = u * g_zig.x[i];
        const double f0 = std::exp(-0.5 * (g_zig.x[i] * g_zig.x[i] - xv * xv)), f1 = std::exp(-0.5 * (g_zig.x[i + 1] * g_zig.x[i + 1] - xv * xv));
        if (f1 + g.u01() * (f0 - f1) < 1.0) return xv;
    }
}

inline double ipow(double x, int n) {   // same multiplication order as oracle/sdde_oracle.py::ipow
    double result = 1.0, base = x;
    bool first = true;
    while (n > 0) {
        if (n & 1) { result = first ? base : result * base; first = false; }
        n >>= 1;
        if (n) base = base * base;
    }
    return result;
}

struct Traj;
typedef void (*LagFn)(Traj &, double t, int l, double *out);   // h(p, t - tau_l)

// ---- model zoo: f(du,u,h,p,t) / g(du,u,h,p,t) written against a history callback like the reference's closures ----
struct Model {
    int ns, nl, nw;
    bool diag;
    std::function<void(double *du, const double *u, const std::function<void(int, double *)> &h, const double *p, double t)> f, g;
};

Model make_model(const sdde_desc *d) {
    Model m;
    m.ns = d->n_state; m.nl = d->n_lag; m.nw = d->n_noise;
    switch (d->model_id) {
        case SDDE_MODEL_LINEAR1:
            m.diag = true;
            m.f = [](double *du, const double *u, const std::function<void(int, double *)> &h, const double *p, double) {
                double ul; h(0, &ul); du[0] = p[0] * u[0] + p[1] * ul; };
            m.g = [](double *du, const double *u, const std::function<void(int, double *)> &, const double *p, double) {
                du[0] = p[2] * u[0]; };
            break;
        case SDDE_MODEL_HAYES:
            m.diag = true;
            m.f = [](double *du, const double *u, const std::function<void(int, double *)> &h, const double *p, double) {
                double ul; h(0, &ul); du[0] = (p[1] * u[0] + p[2] * ul) + p[3]; };
            m.g = [](double *du, const double *u, const std::function<void(int, double *)> &h, const double *p, double) {
                double ul; h(0, &ul); du[0] = (p[4] * u[0] + p[5] * ul) + p[6]; };
            break;
        case SDDE_MODEL_MACKEY_GLASS:
            m.diag = true;
            m.f = [](double *du, const double *u, const std::function<void(int, double *)> &h, const double *p, double) {
                double x; h(0, &x); const double xn = ipow(x, (int)p[2]); du[0] = (p[0] * x) / (1.0 + xn) - p[1] * u[0]; };
            m.g = [](double *du, const double *u, const std::function<void(int, double *)> &, const double *p, double) {
                du[0] = p[3] * u[0]; };
            break;
        case SDDE_MODEL_SCALAR_ADD:
            m.diag = true;
            m.f = [](double *du, const double *, const std::function<void(int, double *)> &h, const double *p, double) {
                double ul; h(0, &ul); du[0] = p[0] * ul; };
            m.g = [](double *du, const double *u, const std::function<void(int, double *)> &, const double *p, double) {
                du[0] = p[1] + 0.0 * u[0]; };
            break;
        case SDDE_MODEL_SIR: {
            m.diag = false;
            auto rates = [](const double *u, const std::function<void(int, double *)> &h, const double *p, double &inf, double &rec) {
                const double N = (u[0] + u[1]) + u[2];
                inf = ((p[0] * p[1]) * u[1]) / N * u[0];
                double ud[3]; h(0, ud);
                const double Nd = (ud[0] + ud[1]) + ud[2];
                rec = ((p[0] * p[1]) * ud[1]) / Nd * ud[0];
            };
            m.f = [rates](double *du, const double *u, const std::function<void(int, double *)> &h, const double *p, double) {
                double inf, rec; rates(u, h, p, inf, rec); du[0] = -inf; du[1] = inf - rec; du[2] = rec; };
            m.g = [rates](double *gv, const double *u, const std::function<void(int, double *)> &h, const double *p, double) {
                double inf, rec; rates(u, h, p, inf, rec);
                const double si = std::sqrt(std::max(0.0, inf)), sr = std::sqrt(std::max(0.0, rec));
                gv[0] = -si; gv[1] = 0.0; gv[2] = si; gv[3] = -sr; gv[4] = 0.0; gv[5] = sr; };
            break;
        }
        case SDDE_MODEL_LINSYS: {
            m.diag = false;
            const int ns = d->n_state, nl = d->n_lag, nw = d->n_noise;
            m.f = [ns, nl](double *du, const double *u, const std::function<void(int, double *)> &h, const double *p, double) {
                double ul[SDDE_MAX_LAG][SDDE_MAX_STATE];
                for (int l = 0; l < nl; ++l) h(l, ul[l]);
                for (int i = 0; i < ns; ++i) {
                    double acc = p[i * ns] * u[0];
                    for (int j = 1; j < ns; ++j) acc = acc + p[i * ns + j] * u[j];
                    for (int l = 0; l < nl; ++l)
                        for (int j = 0; j < ns; ++j) acc = acc + p[ns * ns * (1 + l) + i * ns + j] * ul[l][j];
                    du[i] = acc + p[ns * ns * (1 + nl) + i];
                } };
            m.g = [ns, nl, nw](double *gv, const double *u, const std::function<void(int, double *)> &, const double *p, double) {
                const int offM = ns * ns * (1 + nl) + ns, offN = offM + ns * nw;
                for (int i = 0; i < ns; ++i)
                    for (int j = 0; j < nw; ++j) gv[i * nw + j] = p[offM + i * nw + j] * u[i] + p[offN + i * nw + j]; };
            break;
        }
        default: m.ns = 0;
    }
    return m;
}

struct Disc { double t; int order2; };
struct DiscLater { bool operator()(const Disc &a, const Disc &b) const { return a.t > b.t; } };
typedef std::priority_queue<double, std::vector<double>, std::greater<double>> TimeHeap;
typedef std::priority_queue<Disc, std::vector<Disc>, DiscLater> DiscHeap;

// One trajectory, following src/solve.jl:11-610 for the fixed-step constant-lag case.
// Returns number of steps taken.
int64_t solve_one(const sdde_desc *d, const Model &M, int64_t traj, const double *dW, int64_t n_steps_noise,
                  double *out, int64_t n_save_expected, uint8_t *status) {
    const int alg = d->alg;
    const bool sra = alg >= SDDE_ALG_SRA1 && alg <= SDDE_ALG_SRA_TABLEAU;      // two increments per channel and step: dW, dZ
    const int nc1 = M.diag ? M.ns : M.nw, nc = sra ? 2 * nc1 : nc1;
    const int ns = M.ns, nl = M.nl;
    const int64_t NT = d->n_traj;
    const double t0 = d->t0, tf = d->tf, dt = d->dt;
    // [EXT] StochasticDiffEq.alg_order: EM, EulerHeun 1//2; RKMil, RKMilCommute, WangLi3SMil 1//1; SRA family 2//1
    const int alg_order2 = sra ? (alg == SDDE_ALG_SRA_TABLEAU && d->sra_order2 > 0 ? d->sra_order2 : 4)
                               : ((alg == SDDE_ALG_RKMIL_ITO || alg == SDDE_ALG_RKMIL_STRAT || alg == SDDE_ALG_RKMILC_ITO ||
                                   alg == SDDE_ALG_RKMILC_STRAT || (alg >= SDDE_ALG_WANGLI3SMIL_A && alg <= SDDE_ALG_WANGLI3SMIL_F)) ? 2 : 1);
    double p[SDDE_MAX_PARAM];
    for (int i = 0; i < d->n_param; ++i) p[i] = d->param_stride ? d->params[i * d->param_stride + traj] : d->params[i];
    double u[SDDE_MAX_STATE], uprev[SDDE_MAX_STATE], u0[SDDE_MAX_STATE];
    for (int j = 0; j < ns; ++j) u0[j] = u[j] = d->u0_stride ? d->u0[j * d->u0_stride + traj] : d->u0[j];

    // ---- __init: queues (src/solve.jl:613-713) ----
    const int order0_2 = 2 * d->order_discontinuity_t0;
    const bool add_prop = order0_2 <= alg_order2 && d->constant_lags_declared && nl > 0;
    TimeHeap tst, sv;
    DiscHeap disc;
    if (!add_prop && d->n_tstops == 0 && d->n_d_disc == 0) tst.push(tf);
    else {
        for (int i = 0; i < d->n_tstops; ++i) if (t0 < d->tstops[i] && d->tstops[i] <= tf) tst.push(d->tstops[i]);
        for (int64_t i = 0; i < d->n_d_disc; ++i) if (t0 < d->d_disc_t[i] && d->d_disc_t[i] <= tf) tst.push(d->d_disc_t[i]);   // src/solve.jl:643-646
        if (add_prop) for (int l = 0; l < nl; ++l) if (d->lags[l] < tf - t0) tst.push(t0 + d->lags[l]);
        tst.push(tf);
    }
    for (int64_t i = 0; i < d->n_d_disc; ++i)       // src/solve.jl:690-699
        if (t0 < d->d_disc_t[i] && d->d_disc_t[i] < tf && d->d_disc_order2[i] <= alg_order2 + 2) disc.push(Disc{d->d_disc_t[i], d->d_disc_order2[i]});
    if (add_prop) for (int l = 0; l < nl; ++l) if (d->lags[l] < tf - t0) disc.push(Disc{t0 + d->lags[l], order0_2 + 1});
    for (int i = 0; i < d->n_saveat; ++i) sv.push(d->saveat[i]);

    // ---- dense history = solution arrays (src/utils.jl:143-247, src/solve.jl:465-482) ----
    std::vector<double> ht, hu;
    ht.reserve(1024); hu.reserve(1024 * ns);
    ht.push_back(t0);
    for (int j = 0; j < ns; ++j) hu.push_back(u[j]);

    int64_t ks = 0;
    bool bad = false;
    auto save = [&](const double *v) {
        if (ks < n_save_expected)
            for (int j = 0; j < ns; ++j) { out[(ks * ns + j) * NT + traj] = v[j]; bad = bad || !std::isfinite(v[j]); }
        ++ks;
    };
    if (d->save_start) save(u);

    double t = t0;
    // [EXT] HistoryFunction(p, t): prehistory | interpolation of the dense solution | constant u0
    auto hist = [&](double tl, double *o) {
        if (tl < ht.front()) { for (int j = 0; j < ns; ++j) o[j] = d->hist_c0[j] + d->hist_c1[j] * tl; return; }
        if (tl <= ht.back()) {
            const size_t i = std::lower_bound(ht.begin(), ht.end(), tl) - ht.begin();   // searchsortedfirst
            if (ht[i] == tl) { for (int j = 0; j < ns; ++j) o[j] = hu[i * ns + j]; return; }
            const double th = (tl - ht[i - 1]) / (ht[i] - ht[i - 1]);
            for (int j = 0; j < ns; ++j) o[j] = (1.0 - th) * hu[(i - 1) * ns + j] + th * hu[i * ns + j];
            return;
        }
        for (int j = 0; j < ns; ++j) o[j] = u0[j];
    };
    double tq = t0;   // time at which the current f/g evaluation asks for its delayed states
    const std::function<void(int, double *)> hcb = [&](int l, double *o) { hist(tq - d->lags[l], o); };

    int64_t n = 0;
    double zc[4] = {0, 0, 0, 0}; uint64_t zcall = ~0ull;
    Xoro xoro(d->seed, (uint64_t)(d->traj_offset + traj));
    double curW[2 * SDDE_MAX_NOISE] = {0};
    if (d->noise_mode == SDDE_NOISE_SUPPLIED_GRID) for (int c = 0; c < nc; ++c) curW[c] = dW[(int64_t)c * NT + traj];
    double du[SDDE_MAX_STATE], K[SDDE_MAX_STATE], gv[SDDE_MAX_STATE * SDDE_MAX_NOISE], g2[SDDE_MAX_STATE], ut[SDDE_MAX_STATE], dw[2 * SDDE_MAX_NOISE];
    while (!tst.empty()) {
        while (t < tst.top()) {
            // loopheader! / apply_step!: uprev <- u, dt_n <- min(dt, tstop - t), draw dW ~ N(0,|dt_n|)
            const double top = tst.top();
            const double dtn = dt != 0.0 ? std::min(dt, top - t) : top - t;
            const double sq = std::sqrt(std::fabs(dtn));
            for (int j = 0; j < ns; ++j) uprev[j] = u[j];
            for (int c = 0; c < nc; ++c) {
                if (d->noise_mode == SDDE_ORACLE_NOISE_ZIGGURAT) {
                    dw[c] = sq * zig_normal(xoro);
                } else if (d->noise_mode == SDDE_NOISE_PHILOX) {
                    const uint64_t q = (uint64_t)n * nc + c, call = q >> 2;
                    if (call != zcall) { normal_quad((uint64_t)(d->traj_offset + traj), call, d->seed, zc); zcall = call; }
                    dw[c] = sq * zc[q & 3];
                } else if (d->noise_mode == SDDE_NOISE_SUPPLIED_DW) {
                    dw[c] = n < n_steps_noise ? dW[(n * nc + c) * NT + traj] : 0.0;
                } else {
                    const double w = dW[((n + 1) * nc + c) * NT + traj];
                    dw[c] = w - curW[c]; curW[c] += dw[c];
                }
            }
            // perform_step! ([EXT] StochasticDiffEq, call site src/solve.jl:586)
            tq = t;
            M.f(du, uprev, hcb, p, t);
            for (int j = 0; j < ns; ++j) K[j] = uprev[j] + dtn * du[j];
            if (sra) {
                // [EXT] StochasticDiffEq SRAConstantCache (Roessler SRA, additive noise), tableau supplied by the caller
                // (oracle/sdde_oracle.py holds the restated SRA1 / SRA3 / SOSRA coefficients): see solve_ensemble there.
                const int S = d->sra_stages, nwv = M.diag ? 1 : M.nw, ng = M.diag ? ns : ns * M.nw;
                const double *c0 = d->sra_tableau, *c1 = c0 + S, *A0 = c1 + S, *B0 = A0 + S * S, *al = B0 + S * S, *b1 = al + S, *b2 = b1 + S;
                double chi2[SDDE_MAX_NOISE], fs[3][SDDE_MAX_STATE], gs[3][SDDE_MAX_STATE * SDDE_MAX_NOISE], H[SDDE_MAX_STATE];
                for (int c = 0; c < nc1; ++c) chi2[c] = 0.5 * (dw[c] + dw[nc1 + c] / std::sqrt(3.0));
                for (int i = 0; i < S; ++i) {
                    for (int j = 0; j < ns; ++j) H[j] = uprev[j];
                    if (i > 0) {
                        double A[SDDE_MAX_STATE], B[SDDE_MAX_STATE * SDDE_MAX_NOISE];
                        for (int j = 0; j < ns; ++j) A[j] = A0[i * S] * fs[0][j];
                        for (int q = 0; q < ng; ++q) B[q] = B0[i * S] * gs[0][q];
                        for (int jj = 1; jj < i; ++jj) {
                            for (int j = 0; j < ns; ++j) A[j] = A[j] + A0[i * S + jj] * fs[jj][j];
                            for (int q = 0; q < ng; ++q) B[q] = B[q] + B0[i * S + jj] * gs[jj][q];
                        }
                        for (int j = 0; j < ns; ++j) {
                            double noise = B[j * nwv] * chi2[M.diag ? j : 0];
                            for (int k = 1; k < nwv; ++k) noise = noise + B[j * nwv + k] * chi2[k];
                            H[j] = (uprev[j] + dtn * A[j]) + noise;
                        }
                    }
                    tq = t + c0[i] * dtn;
                    M.f(fs[i], H, hcb, p, tq);
                    tq = t + c1[i] * dtn;
                    M.g(gs[i], H, hcb, p, tq);
                }
                tq = t;
                for (int j = 0; j < ns; ++j) {
                    double at = al[0] * fs[0][j];
                    for (int i = 1; i < S; ++i) at = at + al[i] * fs[i][j];
                    double bn = 0.0, en = 0.0;
                    for (int k = 0; k < nwv; ++k) {
                        double bt = b1[0] * gs[0][j * nwv + k], e2 = b2[0] * gs[0][j * nwv + k];
                        for (int i = 1; i < S; ++i) { bt = bt + b1[i] * gs[i][j * nwv + k]; e2 = e2 + b2[i] * gs[i][j * nwv + k]; }
                        const int c = M.diag ? j : k;
                        bn = k == 0 ? bt * dw[c] : bn + bt * dw[c];
                        en = k == 0 ? e2 * chi2[c] : en + e2 * chi2[c];
                    }
                    u[j] = ((uprev[j] + dtn * at) + bn) + en;
                }
            } else if (alg == SDDE_ALG_EULER_HEUN) {
                // [EXT] EulerHeunConstantCache; second evaluation at t + dt
                const int nwv = M.diag ? 1 : M.nw;
                double g1[SDDE_MAX_STATE * SDDE_MAX_NOISE], gg2[SDDE_MAX_STATE * SDDE_MAX_NOISE], tmp[SDDE_MAX_STATE], f2[SDDE_MAX_STATE];
                M.g(g1, uprev, hcb, p, t);
                for (int i = 0; i < ns; ++i) {
                    double noise = g1[i * nwv] * dw[M.diag ? i : 0];
                    for (int j = 1; j < nwv; ++j) noise = noise + g1[i * nwv + j] * dw[j];
                    tmp[i] = (uprev[i] + du[i] * dtn) + noise;
                }
                tq = t + dtn;
                M.f(f2, tmp, hcb, p, t + dtn);
                M.g(gg2, tmp, hcb, p, t + dtn);
                tq = t;
                for (int i = 0; i < ns; ++i) {
                    double noise2 = (0.5 * (g1[i * nwv] + gg2[i * nwv])) * dw[M.diag ? i : 0];
                    for (int j = 1; j < nwv; ++j) noise2 = noise2 + (0.5 * (g1[i * nwv + j] + gg2[i * nwv + j])) * dw[j];
                    u[i] = (uprev[i] + (dtn * 0.5) * (du[i] + f2[i])) + noise2;
                }
            } else if (alg == SDDE_ALG_EM_SPLIT || alg == SDDE_ALG_EM_NOSPLIT) {
                M.g(gv, alg == SDDE_ALG_EM_SPLIT ? K : uprev, hcb, p, t);
                if (M.diag) for (int j = 0; j < ns; ++j) u[j] = K[j] + gv[j] * dw[j];
                else for (int i = 0; i < ns; ++i) {
                    double noise = gv[i * M.nw] * dw[0];
                    for (int j = 1; j < M.nw; ++j) noise = noise + gv[i * M.nw + j] * dw[j];
                    u[i] = K[i] + noise;
                }
            } else if (alg >= SDDE_ALG_WANGLI3SMIL_A && alg <= SDDE_ALG_WANGLI3SMIL_F) {
                // [EXT, LOW confidence] WangLi3SMil_A..F = the six orders of D (drift), C (Ito correction), M (Stratonovich
                // Milstein diffusion), every f / g at the step's t; see oracle/sdde_oracle.py for the operators
                static const char *orders[6] = {"DCM", "DMC", "CDM", "CMD", "MDC", "MCD"};
                const char *ord = orders[alg - SDDE_ALG_WANGLI3SMIL_A];
                double k[SDDE_MAX_STATE], k1[SDDE_MAX_STATE], tmp[SDDE_MAX_STATE];
                for (int j = 0; j < ns; ++j) u[j] = uprev[j];
                for (int s3 = 0; s3 < 3; ++s3) {
                    if (ord[s3] == 'D') {
                        M.f(du, u, hcb, p, t);
                        for (int j = 0; j < ns; ++j) u[j] = u[j] + dtn * du[j];
                    } else {
                        M.g(k, u, hcb, p, t);
                        for (int j = 0; j < ns; ++j) tmp[j] = u[j] + k[j] * sq;
                        M.g(k1, tmp, hcb, p, t);
                        for (int j = 0; j < ns; ++j) {
                            const double dgg = (k1[j] - k[j]) / sq;
                            if (ord[s3] == 'C') u[j] = u[j] - (dtn * 0.5) * dgg;
                            else u[j] = (u[j] + k[j] * dw[j]) + ((dw[j] * dw[j]) * 0.5) * dgg;
                        }
                    }
                }
            } else if ((alg == SDDE_ALG_RKMILC_ITO || alg == SDDE_ALG_RKMILC_STRAT) && !M.diag) {
                // [EXT] RKMilCommuteConstantCache on a noise matrix: J = dW dW'/2 (Ito: - |dt|/2 I); per column j
                // K_j = (Ito ? K : uprev) + sqrt(dt) G[:, j], Dg_j = (G(K_j) - G(u)) / sqrt(dt), correction += Dg_j J[:, j]
                const int nw = M.nw;
                double gt[SDDE_MAX_STATE * SDDE_MAX_NOISE], Kj[SDDE_MAX_STATE], mil[SDDE_MAX_STATE];
                M.g(gv, uprev, hcb, p, t);
                for (int i = 0; i < ns; ++i) mil[i] = 0.0;
                for (int j = 0; j < nw; ++j) {
                    for (int i = 0; i < ns; ++i) Kj[i] = (alg == SDDE_ALG_RKMILC_ITO ? K[i] : uprev[i]) + sq * gv[i * nw + j];
                    M.g(gt, Kj, hcb, p, t);
                    for (int k = 0; k < nw; ++k) {
                        double Jkj = 0.5 * (dw[k] * dw[j]);
                        if (alg == SDDE_ALG_RKMILC_ITO && k == j) Jkj = Jkj - 0.5 * std::fabs(dtn);
                        for (int i = 0; i < ns; ++i) mil[i] = mil[i] + ((gt[i * nw + k] - gv[i * nw + k]) / sq) * Jkj;
                    }
                }
                for (int i = 0; i < ns; ++i) {
                    double noise = gv[i * nw] * dw[0];
                    for (int j = 1; j < nw; ++j) noise = noise + gv[i * nw + j] * dw[j];
                    u[i] = (K[i] + noise) + mil[i];
                }
            } else {
                const bool ito = alg == SDDE_ALG_RKMIL_ITO || alg == SDDE_ALG_RKMILC_ITO;   // RKMilCommute on diagonal noise == RKMil
                M.g(gv, uprev, hcb, p, t);
                for (int j = 0; j < ns; ++j) ut[j] = (ito ? K[j] : uprev[j]) + gv[j] * sq;
                M.g(g2, ut, hcb, p, t);
                for (int j = 0; j < ns; ++j) {
                    const double ggp = (g2[j] - gv[j]) / sq;
                    const double w2 = ito ? (dw[j] * dw[j] - std::fabs(dtn)) : (dw[j] * dw[j]);
                    const double mil = ggp * w2 / 2.0;
                    u[j] = (K[j] + gv[j] * dw[j]) + mil;
                }
            }
            // loopfooter! non-adaptive (src/integrators/interface.jl:70-85)
            const double ttmp = t + dtn, tprev = t;
            t = (std::fabs(ttmp - top) < 10.0 * jl_eps(std::max(t, top))) ? top : ttmp;
            // savevalues! (src/integrators/interface.jl:357-430)
            while (!sv.empty() && sv.top() <= t) {
                const double curt = sv.top(); sv.pop();
                if (curt != t) {
                    const double th = (curt - tprev) / dtn;
                    double val[SDDE_MAX_STATE];
                    for (int j = 0; j < ns; ++j) val[j] = (1.0 - th) * uprev[j] + th * u[j];
                    save(val);
                } else save(u);
            }
            if (d->save_everystep) save(u);
            ht.push_back(t);
            for (int j = 0; j < ns; ++j) hu.push_back(u[j]);
            ++n;
            if (tst.empty()) break;
        }
        // handle_tstop! / handle_discontinuities! / add_next_discontinuities! (src/integrators/utils.jl)
        if (!tst.empty()) {
            if (t == tst.top()) tst.pop();
            if (!disc.empty()) {
                int order2 = disc.top().order2; disc.pop();
                while (!disc.empty() && disc.top().t == t) { order2 = std::min(order2, disc.top().order2); disc.pop(); }
                const double maxdt = 10.0 * jl_eps(t);
                while (!disc.empty() && std::fabs(disc.top().t - t) < maxdt) { order2 = std::min(order2, disc.top().order2); disc.pop(); }
                while (!tst.empty() && std::fabs(tst.top() - t) < maxdt) tst.pop();
                const int next2 = order2 + 1;
                if (next2 <= alg_order2 + 1 && d->constant_lags_declared)
                    for (int l = 0; l < nl; ++l)
                        if (d->lags[l] < tf - t) { disc.push(Disc{t + d->lags[l], next2}); tst.push(t + d->lags[l]); }
            }
        }
    }
    // postamble! (src/integrators/interface.jl:100-117)
    if (d->save_end) {
        // last saved time != t  <=>  the final state was not saved by saveat/save_everystep
        bool already = d->save_everystep;
        if (!already && d->n_saveat > 0) {
            double mx = d->saveat[0];
            for (int i = 1; i < d->n_saveat; ++i) mx = std::max(mx, d->saveat[i]);
            already = (mx == t);
        }
        if (!already && n == 0 && d->save_start) already = true;
        if (!already) save(u);
    }
    for (int j = 0; j < ns; ++j) bad = bad || !std::isfinite(u[j]);
    if (status) status[traj] = bad ? 1 : 0;
    return n;
}

}  // namespace

extern "C" {

// out: [n_save][n_state][n_traj] (host).  noise: as in sdde_desc (host).  Returns 0, or -2 for an unknown model.
__attribute__((visibility("default"))) int sdde_oracle_solve(const sdde_desc *d, int64_t n_save, int64_t n_steps_noise, double *out,
                                                              uint8_t *status, int n_threads, double *elapsed_s, int64_t *steps_taken) {
    Model M = make_model(d);
    if (M.ns == 0) return -2;
    if (n_threads < 1) n_threads = 1;
    const int64_t NT = d->n_traj;
    std::vector<int64_t> steps((size_t)n_threads, 0);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int w = 0; w < n_threads; ++w) {
        th.emplace_back([&, w]() {
            const int64_t lo = NT * w / n_threads, hi = NT * (w + 1) / n_threads;   // static chunks (Threads.@threads)
            int64_t s = 0;
            for (int64_t i = lo; i < hi; ++i) s = solve_one(d, M, i, d->noise, n_steps_noise, out, n_save, status);
            steps[(size_t)w] = s;
        });
    }
    for (auto &t : th) t.join();
    if (elapsed_s) *elapsed_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (steps_taken) { *steps_taken = 0; for (int64_t s : steps) *steps_taken = std::max(*steps_taken, s); }
    return 0;
}

}  // extern "C"
