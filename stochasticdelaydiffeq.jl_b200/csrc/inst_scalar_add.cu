// inst_scalar_add.cu -- kernel instantiations for ModelScalarAdd
#include "sdde_dispatch.cuh"
namespace sdde {
SDDE_DEFINE_SELECT_DIAG_SRA(select_scalar_add, ModelScalarAdd, "scalar_add")
}
