// inst_mackey_glass10.cu -- kernel instantiations for Mackey-Glass with the exponent n = 10 fixed at compile time
// (BASELINE.json configs[1], configs[2]): x^10 is four multiplications and the division is branch-free.
#include "sdde_dispatch.cuh"
namespace sdde {
SDDE_DEFINE_SELECT_DIAG(select_mackey_glass10, ModelMackeyGlass10, "mackey_glass_n10")
}
