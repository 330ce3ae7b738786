// nccl_dyn.cpp -- see nccl_dyn.hpp.
#include "nccl_dyn.hpp"

#include <dlfcn.h>

#include <mutex>

namespace sdde {

const NcclApi &nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
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
