// summary.cpp -- see summary.hpp.  Compile with -ffp-contract=off.
#include "summary.hpp"

#include <algorithm>
#include <cmath>

namespace sdde {

namespace {
int rank_slot(std::vector<int64_t> &ranks, int64_t r) {
    for (size_t i = 0; i < ranks.size(); ++i)
        if (ranks[i] == r) return (int)i;
    ranks.push_back(r);
    return (int)ranks.size() - 1;
}
}  // namespace

std::string plan_quantiles(int64_t n, const double *probs, int n_prob, bool want_median, QuantilePlan &plan) {
    plan = QuantilePlan();
    if (n < 1) return "quantiles need at least one trajectory";
    std::vector<int64_t> ranks;
    std::vector<QuantileItem> items;
    for (int i = 0; i < n_prob; ++i) {
        const double p = probs[i];
        if (!(p >= 0.0 && p <= 1.0)) return "quantile probabilities must lie in [0, 1]";   // Statistics.quantile ArgumentError
        QuantileItem it;
        if (n == 1) {
            it.ia = it.ib = rank_slot(ranks, 0);
        } else {
            // Statistics._quantile, alpha = beta = 1:  m = 1 - p;  aleph = n*p + m;  j = clamp(trunc(aleph), 1, n-1)
            const double m = 1.0 + p * (1.0 - 1.0 - 1.0);
            const double aleph = (double)n * p + m;
            int64_t j = (int64_t)std::trunc(aleph);
            j = std::min<int64_t>(std::max<int64_t>(j, 1), n - 1);
            double g = aleph - (double)j;
            g = std::min(std::max(g, 0.0), 1.0);
            it.gamma = g;
            it.ia = rank_slot(ranks, j - 1);       // v[j]   (1-based) -> order statistic j-1
            it.ib = rank_slot(ranks, j);           // v[j+1]
        }
        items.push_back(it);
    }
    if (want_median) {
        QuantileItem it;
        if (n % 2 == 1) { it.kind = 1; it.ia = it.ib = rank_slot(ranks, (n - 1) / 2); }
        else { it.kind = 2; it.ia = rank_slot(ranks, n / 2 - 1); it.ib = rank_slot(ranks, n / 2); }
        items.push_back(it);
    }
    // ascending ranks, items re-pointed
    std::vector<int> order(ranks.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return ranks[(size_t)a] < ranks[(size_t)b]; });
    std::vector<int> where(ranks.size());
    for (size_t i = 0; i < order.size(); ++i) { where[(size_t)order[i]] = (int)i; plan.ranks.push_back(ranks[(size_t)order[i]]); }
    for (auto &it : items) { it.ia = where[(size_t)it.ia]; it.ib = where[(size_t)it.ib]; }
    plan.items = items;
    return std::string();
}

double combine_quantile(const QuantileItem &it, double a, double b) {
    if (it.kind == 1) return a;
    if (it.kind == 2) return a / 2 + b / 2;                                   // Statistics.middle(a, b)
    if (std::isfinite(a) && std::isfinite(b)) return a + it.gamma * (b - a);  // Statistics._quantile
    return (1.0 - it.gamma) * a + it.gamma * b;
}

}  // namespace sdde
