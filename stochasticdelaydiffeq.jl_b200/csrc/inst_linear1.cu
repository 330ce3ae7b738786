// inst_linear1.cu -- kernel instantiations for ModelLinear1
#include "sdde_dispatch.cuh"
namespace sdde {
SDDE_DEFINE_SELECT_DIAG(select_linear1, ModelLinear1, "linear1")
}
