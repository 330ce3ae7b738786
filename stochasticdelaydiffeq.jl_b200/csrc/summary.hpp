// summary.hpp -- host half of the ensemble order statistics (sdde_b200_quantiles): which order statistics a set of
// probabilities needs, and how two neighbouring order statistics combine into the reported value.
//
// Product code.  Restates, in float64 without FMA contraction, what the reference side of the boundary computes with
//   [EXT] SciMLBase EnsembleAnalysis.timeseries_steps_quantile / timeseries_steps_median  (used by EnsembleSummary)
//   -> [EXT] Statistics.quantile(v, p)  (alpha = beta = 1, "type 7")  and  Statistics.median(v)
// on the componentwise vectors of one saved time point across trajectories.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace sdde {

struct QuantileItem {
    int ia = 0, ib = 0;      // indices into QuantilePlan::ranks of the two order statistics a <= b
    double gamma = 0.0;      // interpolation weight (quantile only)
    int kind = 0;            // 0 = quantile, 1 = median of an odd count (a), 2 = median of an even count (a/2 + b/2)
};

struct QuantilePlan {
    std::vector<int64_t> ranks;          // distinct 0-based order statistics to select, ascending
    std::vector<QuantileItem> items;     // one per requested output column
};

// n = number of trajectories; probs in [0, 1]; a median column is appended when want_median.
std::string plan_quantiles(int64_t n, const double *probs, int n_prob, bool want_median, QuantilePlan &plan);
double combine_quantile(const QuantileItem &it, double a, double b);

}  // namespace sdde
