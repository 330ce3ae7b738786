// sdde_rng_tables.cuh -- layout and polynomial coefficients of the in-kernel Box-Muller transform (NVRTC-clean).
//
// The table itself is computed once per process on the host in extended precision (rng_table.cpp, checked against
// mpmath by tests/test_abi_host.py) and uploaded once per context.  Layout in global memory (double2 entries):
//   [0, 512)      LOG     k = top 9 mantissa bits of u: {1/c_k, -2 ln c'_k}, c_k = 1 + (k+1/2)/512 (c_511 = 2),
//                         c'_k = c_k (k < 256) or c_k/2
//   [512, 1024)   COARSE  i = (h << 8) | a, h = low quadrant bit, a = top 8 bits of the 22-bit angle fraction:
//                         {cos, sin}(alpha_i), alpha_i = (pi/2) (h + (a + 1/2) 2^-8 - 1/2)            (half a turn)
//   [1024, 1152)  MID     b = next 7 bits:  {cos, sin}(beta_b),  beta_b  = (pi/2) ((b + 1/2) 2^-15 - 2^-9)
//   [1152, 1280)  FINE    c = last 7 bits:  {cos, sin}(gamma_c), gamma_c = (pi/2) ((c + 1/2) 2^-22 - 2^-16)
// so that the angle of a draw, theta = (pi/2)(q + (f + 1/2) 2^-22 - 1/2) with q = 2 s + h, is alpha + beta + gamma + s pi:
// cos / sin follow from TWO angle additions (8 FP64 instructions) and a sign flip -- no polynomial, no quadrant selects.
//
// Shared memory: every CTA that draws normals stages SDDE_RNG_COPIES = 8 interleaved copies of the table, entry i of copy c
// at double2 index 8 i + c, and lane L reads copy L & 7: the eight lanes of a quarter warp (one 128-byte shared-memory
// wavefront) then hit eight different 16-byte bank groups whatever their random indices are, so every 128-bit table load
// costs its minimum of 4 wavefronts.  (Measured, profiles/README.md r2b: with ONE copy the random lookups averaged ~10
// wavefronts each and the LSU data pipe, at 78 %, was the unit that bounded the step kernel.)  8 x 20 KB = 160 KB per CTA.
#pragma once
#define SDDE_RNG_LOG_ENTRIES 512
#define SDDE_RNG_OFF_COARSE 512
#define SDDE_RNG_OFF_MID 1024
#define SDDE_RNG_OFF_FINE 1152
#define SDDE_RNG_TAB_ENTRIES 1280
#define SDDE_RNG_COPIES 8
#define SDDE_RNG_TWO_LN2 1.3862943611198906
#ifdef __CUDACC__
namespace sdde {
// log1p(r) = r + r^2 (L[0] + L[1] r + L[2] r^2 + L[3] r^3), |r| <= 2^-9
static __constant__ double RNG_LOG1P_C[4] = {-0.5, 0.3333333333333333, -0.25, 0.2};
}  // namespace sdde
#endif
