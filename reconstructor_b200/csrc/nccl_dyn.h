// nccl_dyn.h -- the handful of NCCL entry points the collective ingest needs, bound at run time (dlopen).
//
// The library has no link-time dependency on NCCL: a host that never calls pm_comm_init pays nothing, and inside a
// process that already carries an NCCL (e.g. the one PyTorch ships) dlopen("libnccl.so.2") resolves to that same
// copy, so one process never mixes two NCCL versions.  Declarations follow nccl.h (NCCL 2.x ABI).
#pragma once

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

namespace pm {

struct NcclUniqueId { char internal[128]; };       // ncclUniqueId
typedef struct ncclComm* NcclComm;                 // ncclComm_t
enum { kNcclSuccess = 0, kNcclUint8 = 1 };         // ncclResult_t / ncclDataType_t values used here

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;

  // Returns nullptr on success, else a description of what is missing.
  const char* load() {
    if (lib) return nullptr;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return "libnccl.so.2 not found (dlopen)";
    auto sym = [&](const char* s) { return dlsym(lib, s); };
    GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
    CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
    AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
    GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
    GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllGather || !GroupStart || !GroupEnd || !GetErrorString) {
      lib = nullptr;
      return "libnccl.so.2 lacks a required symbol";
    }
    return nullptr;
  }
};

inline NcclApi& nccl_api() {
  static NcclApi api;
  return api;
}

}  // namespace pm
