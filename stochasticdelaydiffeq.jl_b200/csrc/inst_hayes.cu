// inst_hayes.cu -- kernel instantiations for ModelHayes
#include "sdde_dispatch.cuh"
namespace sdde {
SDDE_DEFINE_SELECT_DIAG_SRA(select_hayes, ModelHayes, "hayes")
}
