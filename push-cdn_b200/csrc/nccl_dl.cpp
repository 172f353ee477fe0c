// nccl_dl.cpp — see nccl_dl.h
#include "nccl_dl.h"

#include <dlfcn.h>

#include <mutex>
#include <string>

namespace pcdn {

const NcclApi* nccl_api(const char** why) {
  static std::once_flag once;
  static NcclApi api;
  static bool ok = false;
  static std::string err;
  std::call_once(once, [] {
    // RTLD_NOLOAD first: reuse the copy the process already has (PyTorch ships its own libnccl.so.2)
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { const char* e = dlerror(); err = e ? e : "dlopen(libnccl.so.2) failed"; return; }
    auto sym = [&](const char* n) -> void* {
      void* p = dlsym(h, n);
      if (!p && err.empty()) err = std::string("libnccl: missing symbol ") + n;
      return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.CommCount = (decltype(api.CommCount))sym("ncclCommCount");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    ok = err.empty();
  });
  if (!ok && why) *why = err.c_str();
  return ok ? &api : nullptr;
}

}  // namespace pcdn
