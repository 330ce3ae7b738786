// schedule.cpp -- see schedule.hpp.  Compile with -ffp-contract=off (pure float64 scalar arithmetic).
#include "schedule.hpp"

#include <algorithm>
#include <cmath>
#include <functional>
#include <limits>
#include <queue>

namespace sdde {

namespace {

// Julia eps(x::Float64): distance from |x| to the next larger float.
inline double jl_eps(double x) {
    x = std::fabs(x);
    return std::nextafter(x, std::numeric_limits<double>::infinity()) - x;
}

struct Disc {
    double t;
    int order2;  // 2*order (Rational{Int} with denominators 1 or 2 only on this path)
};
struct DiscLater {
    bool operator()(const Disc &a, const Disc &b) const { return a.t > b.t; }
};
using TimeHeap = std::priority_queue<double, std::vector<double>, std::greater<double>>;
using DiscHeap = std::priority_queue<Disc, std::vector<Disc>, DiscLater>;

}  // namespace

std::string build_schedule(const ScheduleInput &in, Schedule &S) {
    const double t0 = in.t0, tf = in.tf, dt = in.dt;
    if (!(tf > t0)) return "tspan must satisfy tf > t0 (only tdir=+1 is on this path)";
    if (!std::isfinite(t0) || !std::isfinite(tf)) return "tspan must be finite";
    if (!(dt >= 0.0) || !std::isfinite(dt)) return "dt must be finite and >= 0";
    // fail before replaying a hopeless schedule (the reference would step for hours; maxiters there is 1e5 by default)
    if (dt > 0.0 && (tf - t0) / dt > (double)in.max_steps) return "schedule exceeds the maximum number of steps";
    for (double x : in.tstops) if (std::isnan(x)) return "tstops must not be NaN";
    for (double x : in.saveat) if (std::isnan(x)) return "saveat must not be NaN";
    for (double x : in.disc_t) if (std::isnan(x)) return "d_discontinuities must not be NaN";
    if (dt == 0.0 && in.tstops.empty())  // src/solve.jl:121-123
        return "Fixed timestep methods require a choice of dt or choosing the tstops";
    for (double lag : in.lags)
        if (!(lag > 0.0) || !std::isfinite(lag)) return "constant lags must be positive and finite";

    const int L = (int)in.lags.size();
    const int order0_2 = 2 * in.order_discontinuity_t0;
    // src/solve.jl:625-631
    const bool add_prop = (order0_2 <= in.alg_order2) && in.constant_lags_declared && L > 0;
    const int first_next2 = in.neutral ? order0_2 : order0_2 + 1;

    TimeHeap tst;
    DiscHeap disc;
    if (in.disc_t.size() != in.disc_order2.size()) return "d_discontinuities: times and orders differ in length";
    if (in.disc_t.empty() && !add_prop && in.tstops.empty()) {
        tst.push(tf);                                     // src/solve.jl:635-636
    } else {
        for (double x : in.tstops)
            if (t0 < x && x <= tf) tst.push(x);           // :638-641
        for (double x : in.disc_t)
            if (t0 < x && x <= tf) tst.push(x);           // :643-646
        if (add_prop) {
            const double maxlag = tf - t0;
            for (double lag : in.lags)
                if (lag < maxlag) tst.push(t0 + lag);     // :649-655
        }
        tst.push(tf);                                     // :657
    }
    for (size_t i = 0; i < in.disc_t.size(); ++i)         // :690-699: interior, order <= alg_maximum_order + 1
        if (t0 < in.disc_t[i] && in.disc_t[i] < tf && in.disc_order2[i] <= in.alg_order2 + 2)
            disc.push(Disc{in.disc_t[i], in.disc_order2[i]});
    if (add_prop) {
        const double maxlag = tf - t0;
        for (double lag : in.lags)
            if (lag < maxlag) disc.push(Disc{t0 + lag, first_next2});  // :701-710
    }
    TimeHeap sv;
    for (double x : in.saveat) sv.push(x);

    S = Schedule();
    S.n_lag = L;
    S.t.push_back(t0);
    if (in.save_start) {
        S.save_t.push_back(t0); S.save_step.push_back(-1); S.save_theta.push_back(1.0); S.save_exact.push_back(1);
    }

    double t = t0;
    int64_t n = 0;
    while (!tst.empty()) {
        while (t < tst.top()) {
            const double top = tst.top();
            // [EXT] modify_dt_for_tstops! (called at src/solve.jl:564 and from apply_step!)
            const double dtn = (dt != 0.0) ? std::min(dt, top - t) : (top - t);
            // loopfooter!, non-adaptive branch (src/integrators/interface.jl:70-80)
            const double ttmp = t + dtn;
            const double tprev = t;
            if (std::fabs(ttmp - top) < 10.0 * jl_eps(std::max(t, top))) t = top; else t = ttmp;
            S.dt.push_back(dtn);
            S.sqdt.push_back(std::sqrt(std::fabs(dtn)));
            S.t.push_back(t);
            // savevalues! (src/integrators/interface.jl:357-430)
            while (!sv.empty() && sv.top() <= t) {
                const double curt = sv.top(); sv.pop();
                S.save_t.push_back(curt); S.save_step.push_back(n);
                if (curt != t) { S.save_theta.push_back((curt - tprev) / dtn); S.save_exact.push_back(0); }
                else { S.save_theta.push_back(1.0); S.save_exact.push_back(1); }
            }
            if (in.save_everystep) {
                S.save_t.push_back(t); S.save_step.push_back(n); S.save_theta.push_back(1.0); S.save_exact.push_back(1);
            }
            ++n;
            if (n > in.max_steps) return "schedule exceeds the maximum number of steps";
            if (tst.empty()) break;
        }
        // handle_tstop! (src/integrators/utils.jl:1-22)
        if (!tst.empty()) {
            const double top = tst.top();
            if (t == top) { tst.pop(); S.tstops_hit.push_back(t); }
            else if (t > top) return "integrator stepped past a tstop (the reference raises an error here)";
            if (!disc.empty()) {
                // handle_discontinuities! (src/integrators/utils.jl:42-79)
                int order2 = disc.top().order2; disc.pop();
                while (!disc.empty() && disc.top().t == t) { order2 = std::min(order2, disc.top().order2); disc.pop(); }
                const double maxdt = 10.0 * jl_eps(t);
                while (!disc.empty() && std::fabs(disc.top().t - t) < maxdt) {
                    order2 = std::min(order2, disc.top().order2); disc.pop();
                }
                while (!tst.empty() && std::fabs(tst.top() - t) < maxdt) tst.pop();
                // add_next_discontinuities! (src/integrators/utils.jl:91-118)
                const int next2 = in.neutral ? order2 : order2 + 1;
                if (next2 <= in.alg_order2 + 1 && in.constant_lags_declared) {
                    const double maxlag = tf - t;
                    for (double lag : in.lags)
                        if (lag < maxlag) { disc.push(Disc{t + lag, next2}); tst.push(t + lag); }
                }
            }
        }
    }
    // postamble! -> [EXT] solution_endpoint_match_cur_integrator! (src/integrators/interface.jl:100-117)
    if (in.save_end && (S.save_t.empty() || S.save_t.back() != t)) {
        S.save_t.push_back(t); S.save_step.push_back(n - 1); S.save_theta.push_back(1.0); S.save_exact.push_back(1);
    }

    // History lookups: [EXT] HistoryFunction three regimes + LinearInterpolationData.
    // Stage 0 looks up t_n - tau (every f/g call of EM and RKMil); stage 1 (EulerHeun's second evaluation, at t + dt)
    // looks up (t_n + dt_n) - tau against the SAME dense history, which ends at t_n while the step is being taken.
    // SRA tableaus: lookup set k at (t_n + c_k dt_n) - tau (the evaluation times of f / g, deduplicated by the caller).
    const int64_t N = n;
    const int NSTG = in.n_stage;
    S.n_stage = NSTG;
    S.lk_mode.assign((size_t)N * NSTG * L, LK_PRE);
    S.lk_idx.assign((size_t)N * NSTG * L, -1);
    S.lk_theta.assign((size_t)N * NSTG * L, 0.0);
    int64_t need = 1;                                    // max_m (m - lowest node needed by step m)
    int64_t gap = std::numeric_limits<int32_t>::max();
    for (int64_t k = 0; k < N; ++k) {
        const double tn = S.t[(size_t)k];
        for (int sg = 0; sg < NSTG; ++sg) {
            const double cs = in.stage_c.empty() ? (double)sg : in.stage_c[(size_t)sg];
            const double tq = cs == 0.0 ? tn : (cs == 1.0 ? tn + S.dt[(size_t)k] : tn + cs * S.dt[(size_t)k]);
            for (int l = 0; l < L; ++l) {
                const double tl = tq - in.lags[(size_t)l];
                const size_t e = ((size_t)k * NSTG + sg) * L + l;
                if (tl < t0) { S.lk_theta[e] = tl; continue; }           // prehistory h(p, tl)
                if (tl > tn) { S.lk_mode[e] = LK_U0; continue; }         // never-advanced stub integrator: u0
                // searchsortedfirst(sol.t[1..k+1], tl): dense history ends at t_n
                const auto beg = S.t.begin();
                const int64_t i = std::lower_bound(beg, beg + k + 1, tl) - beg;
                if (S.t[(size_t)i] == tl) {
                    S.lk_mode[e] = LK_EXACT; S.lk_idx[e] = i; S.lk_theta[e] = 1.0;
                    need = std::max(need, k - i);
                } else {
                    const double h = S.t[(size_t)i] - S.t[(size_t)i - 1];
                    S.lk_mode[e] = LK_INTERP; S.lk_idx[e] = i;
                    S.lk_theta[e] = (tl - S.t[(size_t)i - 1]) / h;
                    need = std::max(need, k - (i - 1));
                }
                gap = std::min(gap, k - i);
            }
        }
    }
    S.ring_slots = (int)std::max<int64_t>(2, need + 1);
    S.min_gap = (int)gap;
    return std::string();
}

std::string plan_noise_grid(const std::vector<double> &t_nodes, const double *ts, int64_t n_ts, std::vector<GridNode> &out) {
    out.clear();
    if (n_ts < 2) return "a NoiseGrid needs at least two grid times";
    for (int64_t i = 1; i < n_ts; ++i)
        if (!(ts[i] > ts[i - 1])) return "NoiseGrid times must be strictly ascending";
    const double eps64 = std::numeric_limits<double>::epsilon();
    for (double t : t_nodes) {
        if (std::fabs(t - ts[n_ts - 1]) < 100.0 * eps64) t = ts[n_ts - 1];        // calculate_step!(::NoiseGrid)
        if (t > ts[n_ts - 1] || t < ts[0]) return "NoiseGrid does not cover the time span (cannot extrapolate)";
        const int64_t i = std::lower_bound(ts, ts + n_ts, t - 10.0 * eps64) - ts;    // searchsortedfirst(ts, t - 10eps)
        GridNode g;
        if (i < n_ts && ts[i] == t) { g.lo = g.hi = (int32_t)i; g.theta = 1.0; }
        else if (i > 0 && ts[i - 1] == t) { g.lo = g.hi = (int32_t)(i - 1); g.theta = 1.0; }
        else {
            if (i <= 0 || i >= n_ts) return "NoiseGrid does not cover the time span (cannot extrapolate)";
            g.lo = (int32_t)(i - 1); g.hi = (int32_t)i;
            g.theta = (t - ts[i - 1]) / (ts[i] - ts[i - 1]);
        }
        out.push_back(g);
    }
    return std::string();
}

}  // namespace sdde
