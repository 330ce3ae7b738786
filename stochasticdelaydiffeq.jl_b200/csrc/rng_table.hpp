// rng_table.hpp -- host copy of the normal generator's lookup table (layout: sdde_rng_tables.cuh)
#pragma once
namespace sdde {
const double *rng_table_host();   // 2 * rng_table_entries() doubles, built on first use (thread-safe)
int rng_table_entries();
}  // namespace sdde
