// ctx.h -- the part of the C-ABI handle that every translation unit of libgamc_b200.so sees: device, stream, last error,
// profiling events, launch counter.  The library is built from several .cu files (one per kernel family, compiled in
// parallel); each keeps its own state behind `ext[]` and registers a destructor for it.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <map>
#include <cuda_runtime.h>
#include "../../include/gamc_b200.h"
#include "guard.h"

constexpr int GAMC_PROF_POOL = 32768;
enum { GAMC_EXT_BATCHED = 0, GAMC_EXT_BIGD = 1, GAMC_EXT_COUNT = 4 };

struct gamc_ctx_base {
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  std::string err;
  // multi-GPU job: rank of this handle (set by gamc_comm_init; chains-sharded batched jobs need no communicator)
  int nranks = 1, rank = 0;
  // profiling: CUDA events around every launch, on the launching stream
  bool prof = false; std::vector<cudaEvent_t> ev0, ev1; std::vector<int> ev_family; int ev_used = 0;
  long long launches = 0;
  std::map<const void*, size_t> attr_set;   // kernel -> dynamic shared-memory limit currently set on this device
  // per-translation-unit state
  void* ext[GAMC_EXT_COUNT] = {nullptr, nullptr, nullptr, nullptr};
  void (*ext_free[GAMC_EXT_COUNT])(gamc_ctx_base*) = {nullptr, nullptr, nullptr, nullptr};
};

// defined in gamc_abi.cu: the handle is a gamc_ctx whose first base is gamc_ctx_base
gamc_ctx_base* gamc_base(gamc_handle h);

#define CUDA_TRY(h, call)                                                                         \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      (h)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                             \
      return GAMC_CUDA_ERROR;                                                                     \
    }                                                                                             \
  } while (0)

#define REQUIRE(h, cond, status, msg)                                                             \
  do { if (!(cond)) { (h)->err = (msg); return (status); } } while (0)

struct ProfScope {
  gamc_ctx_base* h; int idx = -1;
  ProfScope(gamc_ctx_base* h_, int family) : h(h_) {
    h->launches += 1;
    if (h->prof && h->ev_used < GAMC_PROF_POOL) {
      idx = h->ev_used++;
      h->ev_family[idx] = family;
      cudaEventRecord(h->ev0[idx], h->stream);
    }
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(h->ev1[idx], h->stream); }
};

// Raise a kernel's dynamic shared-memory limit to at least `bytes` (only ever raised: a later, larger request re-sets it).
inline gamc_status ensure_attr(gamc_ctx_base* h, const void* kern, size_t bytes) {
  auto it = h->attr_set.find(kern);
  if (it != h->attr_set.end() && it->second >= bytes) return GAMC_OK;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) { h->err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return GAMC_CUDA_ERROR; }
  h->attr_set[kern] = bytes;
  return GAMC_OK;
}
