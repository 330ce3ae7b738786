// rng_table.cpp -- the lookup table of the in-kernel normal generator, computed once per process on the host.
// Layout: sdde_rng_tables.cuh.  Every entry is evaluated in long double (x87: 64-bit mantissa, aarch64: binary128) and
// rounded once to float64, i.e. correctly rounded up to the rare double-rounding case (<= 0.5005 ulp);
// tests/test_abi_host.py compares the whole table with 200-bit mpmath values.
#include <cmath>
#include <mutex>
#include <vector>

#include "rng_table.hpp"
#include "sdde_rng_tables.cuh"

namespace sdde {

namespace {
std::vector<double> g_table;
std::once_flag g_once;

void build() {
    const long double half_pi = acosl(-1.0L) / 2.0L;
    g_table.assign((size_t)SDDE_RNG_TAB_ENTRIES * 2, 0.0);
    for (int k = 0; k < SDDE_RNG_LOG_ENTRIES; ++k) {
        long double c = 1.0L + ((long double)k + 0.5L) / 512.0L;
        if (k == 511) c = 2.0L;                              // top bin anchored at u = 1: table value exactly 0
        const double inv = (double)(1.0L / c);               // the rounded reciprocal the kernel multiplies by ...
        // ... and -2 ln c' consistent with it: k < 256: -2 ln(1/inv) = 2 ln inv; k >= 256 (mantissa halved, exponent + 1):
        // 2 ln(2 inv), through log1p so that the values near u -> 1 keep full relative accuracy (exactly 0 for k = 511)
        const long double m2ln = k < 256 ? 2.0L * logl((long double)inv) : 2.0L * log1pl(2.0L * (long double)inv - 1.0L);
        g_table[2 * (size_t)k] = inv;
        g_table[2 * (size_t)k + 1] = (double)m2ln;
    }
    auto put = [&](int idx, long double ang) {
        g_table[2 * (size_t)idx] = (double)cosl(ang);
        g_table[2 * (size_t)idx + 1] = (double)sinl(ang);
    };
    for (int i = 0; i < 512; ++i) {
        const int h = i >> 8, a = i & 255;
        put(SDDE_RNG_OFF_COARSE + i, half_pi * ((long double)h + ((long double)a + 0.5L) * 0x1p-8L - 0.5L));
    }
    for (int b = 0; b < 128; ++b) put(SDDE_RNG_OFF_MID + b, half_pi * (((long double)b + 0.5L) * 0x1p-15L - 0x1p-9L));
    for (int c = 0; c < 128; ++c) put(SDDE_RNG_OFF_FINE + c, half_pi * (((long double)c + 0.5L) * 0x1p-22L - 0x1p-16L));
}
}  // namespace

const double *rng_table_host() {
    std::call_once(g_once, build);
    return g_table.data();
}
int rng_table_entries() { return SDDE_RNG_TAB_ENTRIES; }

}  // namespace sdde
