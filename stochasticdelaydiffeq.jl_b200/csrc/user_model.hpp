// user_model.hpp -- user-defined f/g models compiled at run time with NVRTC for sm_100a.
//
// This is the generality seam of the reference: /root/reference/src/functionwrapper.jl:35-36 closes arbitrary Julia
// f(u,h,p,t)/g(u,h,p,t) over the history function.  Julia closures cannot cross a C ABI, so the host language passes
// CUDA C++ source for a `struct UserModel` (same interface as the built-in zoo in sdde_models.cuh); the step kernel
// template of sdde_kernels.cuh is instantiated around it by NVRTC and cached per context.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "sdde_dispatch.cuh"

namespace sdde {

class UserModels {
   public:
    // returns model id (>= 1000) or -1 with err set
    int compile(const char *source, std::string &err, const char *name = nullptr);
    bool select(int model_id, int alg, int noise, bool ptraj, KernelInfo &k, std::string &err);
    bool launch(const KernelInfo &k, unsigned grid, unsigned block, size_t smem, cudaStream_t st, const void *args,
                size_t args_size, std::string &err);
    void clear();

   private:
    struct Entry {
        std::vector<void *> kernels;          // cudaKernel_t per (alg, noise, ptraj)
        int ns = 0, nl = 0, nw = 0, np = 0, tpt = 1, batch = 1, rec_size = 0, block = 128, min_ctas = 1;
        bool diag = true;
        std::string name;
        std::string source;                   // wrapped program text (variants are compiled lazily from it)
    };
    std::vector<Entry> entries_;
    std::vector<void *> libs_;                // loaded cudaLibrary_t handles
};

}  // namespace sdde
