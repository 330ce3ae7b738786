// inst_linsys.cu -- kernel instantiations for the n-state linear delay system (BASELINE.json configs[3] shape 4/2/2)
#include "sdde_dispatch.cuh"
namespace sdde {
typedef ModelLinSys<4, 2, 2> LinSys422;
typedef ModelLinSys<2, 1, 2> LinSys212;
typedef ModelLinSys<2, 1, 1> LinSys211;
SDDE_DEFINE_SELECT_EM(select_linsys422, LinSys422, "linsys_4_2_2")
SDDE_DEFINE_SELECT_EM(select_linsys212, LinSys212, "linsys_2_1_2")
SDDE_DEFINE_SELECT_EM(select_linsys211, LinSys211, "linsys_2_1_1")
KernelInfo select_linsys_coop(int ns, int nl, int nw, int alg, int noise) {
    if (!(ns == 4 && nl == 2 && nw == 2)) return KernelInfo();
#define SDDE_COOP(A, NZ) if (alg == A && noise == NZ) return make_coop_info<LinSys422, A, NZ>("linsys_4_2_2_coop4");
    SDDE_COOP(K_ALG_EM_SPLIT, K_NOISE_PHILOX) SDDE_COOP(K_ALG_EM_SPLIT, K_NOISE_DW) SDDE_COOP(K_ALG_EM_SPLIT, K_NOISE_GRID)
    SDDE_COOP(K_ALG_EM_NOSPLIT, K_NOISE_PHILOX) SDDE_COOP(K_ALG_EM_NOSPLIT, K_NOISE_DW) SDDE_COOP(K_ALG_EM_NOSPLIT, K_NOISE_GRID)
#undef SDDE_COOP
    return KernelInfo();
}
KernelInfo select_linsys(int ns, int nl, int nw, int alg, int noise, bool ptraj) {
    if (ns == 4 && nl == 2 && nw == 2) return select_linsys422(alg, noise, ptraj);
    if (ns == 2 && nl == 1 && nw == 2) return select_linsys212(alg, noise, ptraj);
    if (ns == 2 && nl == 1 && nw == 1) return select_linsys211(alg, noise, ptraj);
    return KernelInfo();
}
}
