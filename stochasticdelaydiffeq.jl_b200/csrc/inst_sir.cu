// inst_sir.cu -- kernel instantiations for ModelSIR
#include "sdde_dispatch.cuh"
namespace sdde {
SDDE_DEFINE_SELECT_EM(select_sir, ModelSIR, "sir")
}
