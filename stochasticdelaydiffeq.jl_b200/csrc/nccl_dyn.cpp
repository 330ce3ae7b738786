// nccl_dyn.cpp -- see nccl_dyn.hpp.
#include "nccl_dyn.hpp"

#include <dlfcn.h>

#include <cstdlib>

#include <mutex>

namespace sdde {

const NcclApi &nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        // One libnccl per process: the loader shares libraries by SONAME, so whichever copy is opened first is the one
        // every later user gets (e.g. a PyTorch imported afterwards, which needs ITS bundled, newer NCCL).  Order:
        // an explicit path (SDDE_B200_NCCL_LIB -- the Python layer points it at the copy PyTorch ships), a copy that
        // is already loaded, the system default.
        void *h = nullptr;
        if (const char *path = getenv("SDDE_B200_NCCL_LIB")) h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) { api.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""); return; }
        auto sym = [&](const char *n) { return dlsym(h, n); };
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.ok = api.CommInitAll && api.CommDestroy && api.AllGather && api.GroupStart && api.GroupEnd && api.GetErrorString;
        if (!api.ok) api.why = "libnccl.so.2 lacks a required symbol";
    });
    return api;
}

}  // namespace sdde
