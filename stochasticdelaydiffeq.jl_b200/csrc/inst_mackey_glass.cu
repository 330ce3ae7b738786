// inst_mackey_glass.cu -- kernel instantiations for ModelMackeyGlass (runtime exponent)
#include "sdde_dispatch.cuh"
namespace sdde {
SDDE_DEFINE_SELECT_DIAG(select_mackey_glass, ModelMackeyGlass, "mackey_glass")
}
