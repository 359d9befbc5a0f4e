// guard.h -- debug build only (make guard, -DGAMC_GUARD): every device allocation of the library is surrounded by two 4 KB guard
// zones filled with a pattern; gamc_debug_guard_check() counts guard bytes that any kernel has overwritten since.  A stand-in for
// `compute-sanitizer --tool memcheck` where that tool is not available (out-of-bounds WRITES next to any buffer are caught; reads
// are not).  Peer-memory (CUDA IPC) paths are not usable in this build (an IPC handle names the allocation, not the offset pointer).
#pragma once
#ifdef GAMC_GUARD
#include <cuda_runtime.h>
#include <cstddef>

cudaError_t gamc_guard_malloc(void** p, size_t bytes);
cudaError_t gamc_guard_free(void* p);
template <typename T> inline cudaError_t gamc_guard_malloc_t(T** p, size_t bytes) { return gamc_guard_malloc(reinterpret_cast<void**>(p), bytes); }
#define cudaMalloc(p, n) gamc_guard_malloc_t((p), (n))
#define cudaFree(p) gamc_guard_free((void*)(p))
#endif
