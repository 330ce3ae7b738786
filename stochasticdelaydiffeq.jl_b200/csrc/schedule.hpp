// schedule.hpp -- host-side, trajectory-independent step schedule of the fixed-step SDDE solve.
//
// Product code (part of libsdde_b200.so).  Mirrors, in float64 and without FMA contraction,
//   * tstop_saveat_disc_handling          /root/reference/src/solve.jl:613-713
//   * the solve! loop                     /root/reference/src/solve.jl:574-610
//   * loopfooter! non-adaptive time snap  /root/reference/src/integrators/interface.jl:70-85
//   * savevalues!                         /root/reference/src/integrators/interface.jl:357-430
//   * handle_tstop!/handle_discontinuities!/add_next_discontinuities!
//                                         /root/reference/src/integrators/utils.jl:1-22,42-79,91-118
//   * [EXT] DelayDiffEq HistoryFunction + StochasticDiffEq LinearInterpolationData lookup
//                                         (wired at /root/reference/src/utils.jl:246, src/solve.jl:407)
// so that every trajectory-independent quantity of one solve (t_n, dt_n, sqrt|dt_n|, the history node
// and theta of every delayed-state lookup, the step and theta of every saved point) is computed once
// on the host exactly as the reference computes it per trajectory, and the kernels only index tables.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace sdde {

enum { LK_PRE = 0, LK_EXACT = 1, LK_INTERP = 2, LK_U0 = 3 };   // LK_U0: past the end of the dense history -> constant u0

struct ScheduleInput {
    double t0 = 0, tf = 0, dt = 0;
    std::vector<double> lags;
    int alg_order2 = 1;                 // 2*alg_order: EM 1, RKMil 2, EulerHeun 1  ([EXT] StochasticDiffEq.alg_order)
    int n_stage = 1;                    // history lookup times per step: stage 0 at t_n, stage 1 at t_n + dt_n (EulerHeun)
    std::vector<double> stage_c;        // lookup set k is taken at t_n + stage_c[k] dt_n; empty = {0, 1, ...} (EulerHeun: {0, 1})
    int order_discontinuity_t0 = 0;
    bool constant_lags_declared = true;
    bool neutral = false;
    std::vector<double> saveat;         // explicit interior(+end) save times
    std::vector<double> tstops;         // user tstops
    std::vector<double> disc_t;         // user d_discontinuities: times ...
    std::vector<int> disc_order2;       // ... and 2 x order
    bool save_start = true, save_end = true, save_everystep = false;
    int64_t max_steps = 200000000;
};

struct Schedule {
    std::vector<double> t, dt, sqdt;
    int n_lag = 0, n_stage = 1;
    std::vector<int32_t> lk_mode;       // [n_steps][n_stage][n_lag]
    std::vector<int64_t> lk_idx;
    std::vector<double> lk_theta;
    std::vector<double> save_t, save_theta;
    std::vector<int64_t> save_step;
    std::vector<int32_t> save_exact;
    std::vector<double> tstops_hit;
    int ring_slots = 2;                 // capacity so that no still-needed node is overwritten
    int min_gap = 0;                    // min_n (n - newest node touched by a lookup of step n); huge if none
    int64_t n_steps() const { return (int64_t)dt.size(); }
};

// Returns empty string on success, else an error message.
std::string build_schedule(const ScheduleInput &in, Schedule &out);

// Where the schedule's nodes t_0..t_N fall on the grid of a supplied NoiseGrid(ts, W): W(t_k) = W[lo] if lo == hi, else
// (1 - theta) W[lo] + theta W[hi].  Mirrors [EXT] DiffEqNoiseProcess interpolate!(::NoiseGrid, u, p, t) (searchsortedfirst
// at t - 10 eps, exact-node shortcuts, linear interpolant) and calculate_step!'s snap of the last node (100 eps).
struct GridNode { int32_t lo, hi; double theta; };
std::string plan_noise_grid(const std::vector<double> &t_nodes, const double *ts, int64_t n_ts, std::vector<GridNode> &out);

}  // namespace sdde
