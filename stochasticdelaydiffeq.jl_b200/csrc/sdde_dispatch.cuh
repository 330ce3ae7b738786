// sdde_dispatch.cuh -- compile-time instantiation table: (model, alg, noise, per-trajectory params) -> kernel.
#pragma once
#include "sdde_kernels.cuh"
#include "sdde_models.cuh"
#include "sdde_coop.cuh"

namespace sdde {

struct KernelInfo {
    const void *fn = nullptr;   // __global__ entry for cudaLaunchKernel
    int block = 128;            // threads per CTA
    int min_ctas = 1;           // CTAs per SM the kernel was compiled for (__launch_bounds__): shared-memory budget
    int tpt = 1;                // trajectories per thread
    int batch = 1;              // prefetch depth B compiled into the kernel
    int rec_size = 0;           // sizeof(StepRec<NL>)
    int n_state = 0, n_lag = 0, n_chan = 0, n_noise = 0, n_param = 0;
    bool diag = true;
    int n_stage = 1;            // history lookup sets per step (2: EulerHeun)
    int lanes = 1;              // lanes per trajectory (cooperative kernel: n_state)
    bool is_user = false;       // fn is a cudaKernel_t obtained from an NVRTC-built library
    const char *name = "";
};

template <class M, int ALG, int NOISE, bool PT>
inline KernelInfo make_info(const char *name) {
    KernelInfo k;
    constexpr int BB = alg_batch(ALG, Tune<M>::B);   // several lookup sets per step: shallower batches
    k.fn = (const void *)&sdde_solve_kernel<M, ALG, NOISE, PT, Tune<M>::TPT, BB>;
    k.block = Tune<M>::BLOCK;
    k.min_ctas = TuneK<M, ALG>::MINB;
    k.tpt = Tune<M>::TPT;
    k.batch = BB;
    k.n_stage = alg_stages(ALG);
    k.rec_size = (int)sizeof(StepRec<M::NL * alg_stages(ALG)>);
    k.n_state = M::NS; k.n_lag = M::NL; k.n_noise = M::NW; k.n_chan = (M::DIAG ? M::NS : M::NW) * alg_noise_mult(ALG); k.n_param = M::NP;
    k.diag = M::DIAG;
    k.name = name;
    return k;
}

template <class M, int ALG, int NOISE>
inline KernelInfo make_coop_info(const char *name) {
    KernelInfo k;
    k.fn = (const void *)&sdde_coop_kernel<M, ALG, NOISE>;
    k.block = SDDE_BLOCK_COOP;
    k.tpt = 1; k.batch = 1; k.lanes = M::NS;
    k.rec_size = (int)sizeof(StepRec<M::NL>);
    k.n_state = M::NS; k.n_lag = M::NL; k.n_noise = M::NW; k.n_chan = M::NW; k.n_param = M::NP;
    k.diag = false;
    k.name = name;
    return k;
}

#define SDDE_SEL3(M, NAME, A, NZ)                                                   \
    if (alg == A && noise == NZ) return ptraj ? make_info<M, A, NZ, true>(NAME) : make_info<M, A, NZ, false>(NAME);
#define SDDE_SEL_ALG(M, NAME, A) \
    SDDE_SEL3(M, NAME, A, K_NOISE_PHILOX) SDDE_SEL3(M, NAME, A, K_NOISE_DW) SDDE_SEL3(M, NAME, A, K_NOISE_GRID)

// all four algorithms (diagonal-noise models)
#define SDDE_DEFINE_SELECT_DIAG(FUNC, M, NAME)                  \
    KernelInfo FUNC(int alg, int noise, bool ptraj) {           \
        SDDE_SEL_ALG(M, NAME, K_ALG_EM_SPLIT)                   \
        SDDE_SEL_ALG(M, NAME, K_ALG_EM_NOSPLIT)                 \
        SDDE_SEL_ALG(M, NAME, K_ALG_RKMIL_ITO)                  \
        SDDE_SEL_ALG(M, NAME, K_ALG_RKMIL_STRAT)                \
        SDDE_SEL_ALG(M, NAME, K_ALG_EULER_HEUN)                 \
        SDDE_SEL_ALG(M, NAME, K_ALG_WANGLI)                     \
        return KernelInfo();                                    \
    }
// diagonal-noise models whose g can be additive (u-independent): the above + the SRA tableau kernel
#define SDDE_DEFINE_SELECT_DIAG_SRA(FUNC, M, NAME)              \
    KernelInfo FUNC(int alg, int noise, bool ptraj) {           \
        SDDE_SEL_ALG(M, NAME, K_ALG_EM_SPLIT)                   \
        SDDE_SEL_ALG(M, NAME, K_ALG_EM_NOSPLIT)                 \
        SDDE_SEL_ALG(M, NAME, K_ALG_RKMIL_ITO)                  \
        SDDE_SEL_ALG(M, NAME, K_ALG_RKMIL_STRAT)                \
        SDDE_SEL_ALG(M, NAME, K_ALG_EULER_HEUN)                 \
        SDDE_SEL_ALG(M, NAME, K_ALG_WANGLI)                     \
        SDDE_SEL_ALG(M, NAME, K_ALG_SRA)                        \
        return KernelInfo();                                    \
    }
// noise-matrix models: EM, EulerHeun, RKMilCommute ([EXT] RKMil itself requires diagonal noise)
#define SDDE_DEFINE_SELECT_EM(FUNC, M, NAME)                    \
    KernelInfo FUNC(int alg, int noise, bool ptraj) {           \
        SDDE_SEL_ALG(M, NAME, K_ALG_EM_SPLIT)                   \
        SDDE_SEL_ALG(M, NAME, K_ALG_EM_NOSPLIT)                 \
        SDDE_SEL_ALG(M, NAME, K_ALG_EULER_HEUN)                 \
        SDDE_SEL_ALG(M, NAME, K_ALG_RKMILC_ITO)                 \
        SDDE_SEL_ALG(M, NAME, K_ALG_RKMILC_STRAT)               \
        return KernelInfo();                                    \
    }

KernelInfo select_linear1(int alg, int noise, bool ptraj);
KernelInfo select_hayes(int alg, int noise, bool ptraj);
KernelInfo select_mackey_glass(int alg, int noise, bool ptraj);
KernelInfo select_mackey_glass10(int alg, int noise, bool ptraj);
KernelInfo select_scalar_add(int alg, int noise, bool ptraj);
KernelInfo select_sir(int alg, int noise, bool ptraj);
KernelInfo select_linsys(int ns, int nl, int nw, int alg, int noise, bool ptraj);
KernelInfo select_linsys_coop(int ns, int nl, int nw, int alg, int noise);

}  // namespace sdde
