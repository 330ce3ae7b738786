// sdde_rng_tables.cuh -- layout and polynomial coefficients of the in-kernel Box-Muller transform (NVRTC-clean).
//
// The table itself is computed once per process on the host in extended precision (rng_table.cpp, checked against
// mpmath by tests/test_abi_host.py), uploaded once per context and staged into shared memory by every CTA that draws
// in-kernel normals.  Layout (double2 entries):
//   [0, 512)              LOG   k = top 9 mantissa bits of u: {1/c_k, -2 ln c'_k}, c_k = 1 + (k+1/2)/512 (c_511 = 2),
//                               c'_k = c_k (k < 256) or c_k/2
//   [512, 2560)           FINE  b = low 11 bits of the 22-bit angle fraction: {cos, sin}(beta_b),
//                               beta_b = (pi/2) ((b + 1/2) 2^-22 - 2^-12)                       (|beta_b| < 3.9e-4)
//   [2560, 2560 + 4096)   COARSE  i = (h << 11) | a, a = top 11 bits of the fraction, h = low quadrant bit:
//                               {cos, sin}(alpha_i), alpha_i = (pi/2) (h + (a + 1/2) 2^-11 - 1/2)   (half a turn)
// so that the angle of a draw, theta = (pi/2)(q + (f + 1/2) 2^-22 - 1/2) with q = 2 s + h, is alpha + beta + s pi and
// cos / sin follow from ONE angle addition (4 FP64 instructions) and a sign flip -- no polynomial, no quadrant selects.
// SDDE_RNG_HALF = 0 stages only the h = 0 half of COARSE (72 KB instead of 104 KB per CTA) and rotates quadrants in registers.
#pragma once
#ifndef SDDE_RNG_HALF
#define SDDE_RNG_HALF 1
#endif
#define SDDE_RNG_LOG_ENTRIES 512
#define SDDE_RNG_FINE_ENTRIES 2048
#define SDDE_RNG_OFF_FINE 512
#define SDDE_RNG_OFF_COARSE 2560
#define SDDE_RNG_TAB_ENTRIES_MAX (2560 + 4096)
#define SDDE_RNG_TAB_ENTRIES (2560 + (SDDE_RNG_HALF ? 4096 : 2048))
#define SDDE_RNG_TWO_LN2 1.3862943611198906
#ifdef __CUDACC__
namespace sdde {
// log1p(r) = r + r^2 (L[0] + L[1] r + L[2] r^2 + L[3] r^3), |r| <= 2^-9
static __constant__ double RNG_LOG1P_C[4] = {-0.5, 0.3333333333333333, -0.25, 0.2};
}  // namespace sdde
#endif
