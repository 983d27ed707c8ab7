// nb_nccl_dyn.cuh - NCCL bound at run time (dlopen), for the library-collective flavour of the
// tally reduction ("collective" = 0; the default is the library's own peer-memory kernel,
// nb_group.cuh). Bound lazily so that libneutral_b200.so loads - and everything single-GPU
// runs - on a machine without NCCL, and so that inside a Python process that already carries
// torch's NCCL the same copy is used (the loader resolves the soname to the loaded library).
// Only the handful of entry points used are declared, with the public NCCL 2.x ABI
// (nccl.h: ncclGetUniqueId, ncclCommInitRank, ncclCommInitAll, ncclReduceScatter,
// ncclAllGather, ncclGroupStart/End, ncclCommDestroy, ncclGetErrorString).
#pragma once

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

#include <string>

namespace nb {

struct NcclComm;  // opaque
typedef NcclComm* nccl_comm_t;
struct NcclUniqueId { char internal[128]; };
enum { kNcclSuccess = 0, kNcclSum = 0, kNcclDouble = 8 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, NcclUniqueId, int) = nullptr;
  int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*ReduceScatter)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
};

// Returns the process-wide binding, or nullptr with *why filled in.
inline const NcclApi* nccl_api(std::string* why) {
  static NcclApi api;
  static bool tried = false;
  static std::string error;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      error = std::string("libnccl.so.2 could not be loaded: ") + dlerror();
    } else {
      auto bind = [&](const char* sym, void** slot) {
        *slot = dlsym(api.handle, sym);
        if (!*slot && error.empty()) error = std::string("NCCL lacks ") + sym;
      };
      bind("ncclGetUniqueId", (void**)&api.GetUniqueId);
      bind("ncclCommInitRank", (void**)&api.CommInitRank);
      bind("ncclCommInitAll", (void**)&api.CommInitAll);
      bind("ncclCommDestroy", (void**)&api.CommDestroy);
      bind("ncclGetErrorString", (void**)&api.GetErrorString);
      bind("ncclReduceScatter", (void**)&api.ReduceScatter);
      bind("ncclAllGather", (void**)&api.AllGather);
      bind("ncclGroupStart", (void**)&api.GroupStart);
      bind("ncclGroupEnd", (void**)&api.GroupEnd);
    }
  }
  if (!error.empty()) {
    if (why) *why = error;
    return nullptr;
  }
  return &api;
}

}  // namespace nb
