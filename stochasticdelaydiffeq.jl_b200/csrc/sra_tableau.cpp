// sra_tableau.cpp -- see sra_tableau.hpp.
#include "sra_tableau.hpp"

#include <cmath>

#include "../../include/sdde_b200.h"

namespace sdde {

namespace {

void finish(SraHost &t) {
    // deduplicate the evaluation times into lookup sets; set 0 is t itself
    t.n_sets = 1;
    for (int k = 0; k < 4; ++k) t.cset[k] = 0.0;
    auto set_of = [&t](double c) -> int {
        for (int k = 0; k < t.n_sets; ++k)
            if (t.cset[k] == c) return k;
        if (t.n_sets >= 4) return -1;
        t.cset[t.n_sets] = c;
        return t.n_sets++;
    };
    for (int i = 0; i < t.s; ++i) { t.fset[i] = set_of(t.c0[i]); t.gset[i] = set_of(t.c1[i]); }
}

}  // namespace

std::string sra_resolve(int alg, const double *tab, int stages, SraHost &t) {
    t = SraHost();
    if (alg == SDDE_ALG_SRA1) {
        // Roessler (2010) SRA1: strong order 1.5 for additive noise
        t.s = 2;
        t.c0[1] = 0.75; t.c1[0] = 1.0;
        t.A0[1][0] = 0.75; t.B0[1][0] = 1.5;
        t.alpha[0] = 1.0 / 3.0; t.alpha[1] = 2.0 / 3.0;
        t.beta1[0] = 1.0; t.beta2[0] = -1.0; t.beta2[1] = 1.0;
    } else if (alg == SDDE_ALG_SRA3) {
        // Roessler (2010) SRA3
        t.s = 3;
        t.c0[1] = 1.0; t.c0[2] = 0.5; t.c1[0] = 1.0;
        t.A0[1][0] = 1.0; t.A0[2][0] = 0.25; t.A0[2][1] = 0.25;
        t.B0[2][0] = 1.0; t.B0[2][1] = 0.5;
        t.alpha[0] = 1.0 / 6.0; t.alpha[1] = 1.0 / 6.0; t.alpha[2] = 2.0 / 3.0;
        t.beta1[0] = 1.0; t.beta2[0] = -1.0; t.beta2[1] = 1.0;
    } else if (alg == SDDE_ALG_SOSRA) {
        // Rackauckas & Nie (2018) stability-optimised SRA (coefficients restated from memory of constructSOSRA; they satisfy
        // sum alpha = 1, sum beta1 = 1, sum beta2 = 0, beta1.c1 = 1, beta2.c1 = -1, A0 e = c0 to 1e-15 -- tests/test_oracle.py)
        t.s = 3;
        t.alpha[0] = 0.2889874966892885; t.alpha[1] = 0.6859880440839937; t.alpha[2] = 0.025024459226717772;
        t.c0[1] = 0.6923962376159507; t.c0[2] = 1.0;
        t.c1[1] = 0.041248171110700504; t.c1[2] = 1.0;
        t.beta1[0] = -16.792534242221663; t.beta1[1] = 17.514995785380226; t.beta1[2] = 0.27753845684143835;
        t.beta2[0] = 0.4237535769069274; t.beta2[1] = 0.6010381474428539; t.beta2[2] = -1.0247917243497813;
        t.A0[1][0] = 0.6923962376159507; t.A0[2][0] = -3.1609142252828395; t.A0[2][1] = 4.1609142252828395;
        t.B0[1][0] = 1.3371632704399763; t.B0[2][0] = 1.442371048468624; t.B0[2][1] = 1.8632741501139225;
    } else if (alg == SDDE_ALG_SRA_TABLEAU) {
        if (!tab) return "SDDE_ALG_SRA_TABLEAU needs sdde_desc.sra_tableau";
        if (stages < 1 || stages > 3) return "an SRA tableau may have 1 to 3 stages on this path";
        const int s = stages;
        t.s = s;
        const double *p = tab;
        for (int i = 0; i < s; ++i) t.c0[i] = *p++;
        for (int i = 0; i < s; ++i) t.c1[i] = *p++;
        for (int i = 0; i < s; ++i) for (int j = 0; j < s; ++j) t.A0[i][j] = *p++;
        for (int i = 0; i < s; ++i) for (int j = 0; j < s; ++j) t.B0[i][j] = *p++;
        for (int i = 0; i < s; ++i) t.alpha[i] = *p++;
        for (int i = 0; i < s; ++i) t.beta1[i] = *p++;
        for (int i = 0; i < s; ++i) t.beta2[i] = *p++;
        for (int i = 0; i < 5 * s + 2 * s * s; ++i)
            if (!std::isfinite(tab[i])) return "non-finite entry in the SRA tableau";
        for (int i = 0; i < s; ++i)
            for (int j = i; j < s; ++j)
                if (t.A0[i][j] != 0.0 || t.B0[i][j] != 0.0) return "the SRA tableau must be explicit (A0, B0 strictly lower triangular)";
        for (int i = 0; i < s; ++i)
            if (t.c0[i] < 0.0 || t.c0[i] > 1.0 || t.c1[i] < 0.0 || t.c1[i] > 1.0) return "SRA stage times c0, c1 must lie in [0, 1]";
    } else {
        return "not an SRA algorithm";
    }
    finish(t);
    for (int i = 0; i < t.s; ++i)
        if (t.fset[i] < 0 || t.gset[i] < 0) return "the SRA tableau has more than 4 distinct stage times";
    return std::string();
}

}  // namespace sdde
