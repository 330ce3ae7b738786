// nccl_dyn.hpp -- NCCL entry points resolved at run time (dlopen of libnccl.so.2), so that libsdde_b200.so has no
// link-time dependency on NCCL and loads on machines without it.  Used for the path's single collective: the
// all-gather of the per-GPU ensemble moment triples over NVLink.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <string>

namespace sdde {

struct NcclApi {
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    std::string why;     // reason when !ok
};

// Thread-safe, resolves once.
const NcclApi &nccl_api();

}  // namespace sdde
