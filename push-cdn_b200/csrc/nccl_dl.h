// nccl_dl.h — the few NCCL entry points the sharded engine uses, bound at run time.
//
// libpcdn_fanout.so does not link libnccl: a single-GPU broker must load on a host without NCCL.
// A sharded engine (pcdn_config.n_devices > 1 with PCDN_INGEST_NCCL) dlopens "libnccl.so.2" when it
// is created — inside a PyTorch process that resolves to the copy torch already loaded, in a plain
// C/Rust host to the system library — and fails loudly (PCDN_ENODEV) when it is missing.
// Types are declared here with NCCL's documented layout (nccl.h: ncclUniqueId is 128 opaque bytes
// passed by value; ncclComm_t is an opaque pointer; ncclUint8 = 1) so no NCCL header is needed.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace pcdn {

struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* NcclComm;

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int nranks, NcclUniqueId id, int rank) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*CommCount)(NcclComm, int*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Broadcast)(const void* send, void* recv, size_t count, int datatype, int root, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};
constexpr int kNcclUint8 = 1;

// nullptr when libnccl.so.2 cannot be loaded (*why gets the dlerror text); loaded once per process
const NcclApi* nccl_api(const char** why);

}  // namespace pcdn
