// sra_tableau.hpp -- Roessler SRA tableaus (additive noise) of the fixed-step SRA1 / SRA3 / SOSRA algorithms and of a
// caller-supplied tableau ([EXT] StochasticDiffEq constructSRA1 / constructSRA3 / constructSOSRA, used by the reference's
// additive-noise smoke tests /root/reference/test/test_prob_sol.jl:74-87).  Host only.
#pragma once
#include <string>

namespace sdde {

struct SraHost {
    int s = 0;                                  // stages (<= 3)
    double c0[3] = {0, 0, 0}, c1[3] = {0, 0, 0};
    double A0[3][3] = {{0}}, B0[3][3] = {{0}};  // [stage i][earlier stage j], strictly lower triangular
    double alpha[3] = {0, 0, 0}, beta1[3] = {0, 0, 0}, beta2[3] = {0, 0, 0};
    int n_sets = 1;                             // distinct evaluation times t + cset[k] dt (cset[0] = 0), <= 4
    double cset[4] = {0, 0, 0, 0};              // unused sets repeat 0
    int fset[3] = {0, 0, 0}, gset[3] = {0, 0, 0};
    int order2 = 4;                             // 2 x [EXT] alg_order (discontinuity propagation depth)
};

// alg: SDDE_ALG_SRA1 / SRA3 / SOSRA, or SDDE_ALG_SRA_TABLEAU with `tab` = [c0(s) | c1(s) | A0(s x s, row i = stage i) |
// B0(s x s) | alpha(s) | beta1(s) | beta2(s)] and `stages` = s.  Returns "" or an error message.
std::string sra_resolve(int alg, const double *tab, int stages, SraHost &out);

}  // namespace sdde
