// user_model.cpp -- see user_model.hpp.
#include "user_model.hpp"

#include <nvrtc.h>

#include <cstdio>
#include <cstring>

#include "../../include/sdde_b200.h"
#include "embedded_src.inc"

namespace sdde {

namespace {

const int kAlgs = 9, kNoises = 3;
inline int variant_index(int alg, int noise, bool ptraj) { return (alg * kNoises + noise) * 2 + (ptraj ? 1 : 0); }

std::string nvrtc_log(nvrtcProgram prog) {
    size_t n = 0;
    nvrtcGetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) nvrtcGetProgramLog(prog, &log[0]);
    return log;
}

// Compile `src` with one kernel instantiation named by `expr`; returns cubin + lowered name.
bool compile_one(const std::string &src, const std::string &expr, std::vector<char> &cubin, std::string &lowered,
                 std::string &err) {
    const char *hdr_src[] = {(const char *)embedded_sdde_kernels_cuh, (const char *)embedded_sdde_models_cuh,
                             (const char *)embedded_sdde_rng_tables_cuh};
    const char *hdr_names[] = {"sdde_kernels.cuh", "sdde_models.cuh", "sdde_rng_tables.cuh"};
    nvrtcProgram prog;
    if (nvrtcCreateProgram(&prog, src.c_str(), "sdde_user_model.cu", 3, hdr_src, hdr_names) != NVRTC_SUCCESS) {
        err = "nvrtcCreateProgram failed";
        return false;
    }
    if (!expr.empty()) nvrtcAddNameExpression(prog, expr.c_str());
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
    nvrtcResult r = nvrtcCompileProgram(prog, 3, opts);
    if (r != NVRTC_SUCCESS) {
        err = std::string("NVRTC: ") + nvrtcGetErrorString(r) + "\n" + nvrtc_log(prog);
        nvrtcDestroyProgram(&prog);
        return false;
    }
    if (!expr.empty()) {
        const char *low = nullptr;
        if (nvrtcGetLoweredName(prog, expr.c_str(), &low) != NVRTC_SUCCESS || !low) {
            err = "nvrtcGetLoweredName failed for " + expr;
            nvrtcDestroyProgram(&prog);
            return false;
        }
        lowered = low;
    }
    size_t n = 0;
    nvrtcGetCUBINSize(prog, &n);
    cubin.resize(n);
    nvrtcGetCUBIN(prog, cubin.data());
    nvrtcDestroyProgram(&prog);
    return true;
}

std::string wrap_source(const char *user) {
    std::string s;
    s += "#include \"sdde_kernels.cuh\"\n#include \"sdde_models.cuh\"\nnamespace sdde {\n";
    s += user;
    s += "\n}  // namespace sdde\n";
    s += "extern \"C\" __device__ const int sdde_user_dims[10] = {sdde::UserModel::NS, sdde::UserModel::NL, sdde::UserModel::NW,\n"
         "  sdde::UserModel::NP, sdde::UserModel::DIAG ? 1 : 0, sdde::Tune<sdde::UserModel>::TPT, sdde::Tune<sdde::UserModel>::B,\n"
         "  (int)sizeof(sdde::StepRec<sdde::UserModel::NL>), sdde::Tune<sdde::UserModel>::BLOCK, sdde::Tune<sdde::UserModel>::MINB};\n"
         "extern \"C\" __global__ void sdde_user_probe(int *o) { for (int i = 0; i < 10; ++i) o[i] = sdde_user_dims[i]; }\n";
    return s;
}

}  // namespace

int UserModels::compile(const char *source, std::string &err, const char *name) {
    const std::string src = wrap_source(source);
    // 1) probe compile: syntax check + model dimensions
    std::vector<char> cubin;
    std::string lowered;
    if (!compile_one(src, "", cubin, lowered, err)) return -1;
    cudaLibrary_t lib = nullptr;
    cudaError_t e = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) { err = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e); return -1; }
    cudaKernel_t probe = nullptr;
    e = cudaLibraryGetKernel(&probe, lib, "sdde_user_probe");
    int dims[10] = {0};
    if (e == cudaSuccess) {
        int *d = nullptr;
        e = cudaMalloc(&d, sizeof(dims));
        if (e == cudaSuccess) {
            void *args[] = {(void *)&d};
            e = cudaLaunchKernel((const void *)probe, dim3(1), dim3(1), args, 0, 0);
            if (e == cudaSuccess) e = cudaMemcpy(dims, d, sizeof(dims), cudaMemcpyDeviceToHost);
            cudaFree(d);
        }
    }
    cudaLibraryUnload(lib);
    if (e != cudaSuccess) { err = std::string("user model probe failed: ") + cudaGetErrorString(e); return -1; }
    Entry en;
    en.ns = dims[0]; en.nl = dims[1]; en.nw = dims[2]; en.np = dims[3]; en.diag = dims[4] != 0;
    en.tpt = dims[5]; en.batch = dims[6]; en.rec_size = dims[7]; en.block = dims[8]; en.min_ctas = dims[9];
    if (en.ns < 1 || en.ns > SDDE_MAX_STATE || en.nl < 1 || en.nl > SDDE_MAX_LAG || en.nw < 1 || en.nw > SDDE_MAX_NOISE ||
        en.np < 0 || en.np > SDDE_MAX_PARAM) {
        err = "UserModel dimensions out of range";
        return -1;
    }
    en.kernels.assign((size_t)kAlgs * kNoises * 2, nullptr);
    en.name = name ? std::string(name) : "user" + std::to_string(entries_.size());
    en.source = src;
    entries_.push_back(en);
    return SDDE_MODEL_USER_BASE + (int)entries_.size() - 1;
}

bool UserModels::select(int model_id, int alg, int noise, bool ptraj, KernelInfo &k, std::string &err) {
    const int idx = model_id - SDDE_MODEL_USER_BASE;
    if (idx < 0 || idx >= (int)entries_.size()) { err = "unknown user model id"; return false; }
    Entry &en = entries_[(size_t)idx];
    if (!en.diag && (alg == 2 || alg == 3)) { k = KernelInfo(); return true; }   // caller reports "RKMil requires diagonal noise"
    if (en.diag && (alg == 5 || alg == 6)) { k = KernelInfo(); return true; }    // caller falls back to the RKMil kernels
    if (!en.diag && alg == 7) { k = KernelInfo(); return true; }                 // WangLi3SMil: diagonal noise only
    const int v = variant_index(alg, noise, ptraj);
    if (!en.kernels[(size_t)v]) {
        char expr[320];
        snprintf(expr, sizeof(expr),
                 "sdde::sdde_solve_kernel<sdde::UserModel, %d, %d, %s, sdde::Tune<sdde::UserModel>::TPT, "
                 "sdde::alg_batch(%d, sdde::Tune<sdde::UserModel>::B)>",      // the batch depth depends on the algorithm
                 alg, noise, ptraj ? "true" : "false", alg);
        std::vector<char> cubin;
        std::string lowered;
        if (!compile_one(en.source, expr, cubin, lowered, err)) return false;
        cudaLibrary_t lib = nullptr;
        cudaError_t e = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
        if (e != cudaSuccess) { err = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e); return false; }
        cudaKernel_t kern = nullptr;
        e = cudaLibraryGetKernel(&kern, lib, lowered.c_str());
        if (e != cudaSuccess) { err = std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(e); cudaLibraryUnload(lib); return false; }
        en.kernels[(size_t)v] = (void *)kern;
        libs_.push_back((void *)lib);
    }
    k = KernelInfo();
    k.fn = en.kernels[(size_t)v];
    k.is_user = true;
    k.block = en.block; k.min_ctas = en.min_ctas;
    k.tpt = en.tpt; k.batch = alg_batch(alg, en.batch);
    k.n_stage = alg_stages(alg);
    k.rec_size = 48 + 32 * en.nl * k.n_stage;
    k.n_state = en.ns; k.n_lag = en.nl; k.n_noise = en.nw; k.n_chan = (en.diag ? en.ns : en.nw) * alg_noise_mult(alg); k.n_param = en.np;
    k.diag = en.diag;
    k.name = en.name.c_str();
    return true;
}

bool UserModels::launch(const KernelInfo &k, unsigned grid, unsigned block, size_t smem, cudaStream_t st, const void *args,
                        size_t, std::string &err) {
    void *a[] = {const_cast<void *>(args)};
    cudaError_t e = cudaLaunchKernel(k.fn, dim3(grid), dim3(block), a, smem, st);
    if (e != cudaSuccess) { err = std::string("launch of user kernel failed: ") + cudaGetErrorString(e); return false; }
    return true;
}

void UserModels::clear() {
    for (void *l : libs_) cudaLibraryUnload((cudaLibrary_t)l);
    libs_.clear();
    entries_.clear();
}

}  // namespace sdde
