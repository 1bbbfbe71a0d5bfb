// Shared host-side bits: launch accounting and small helpers.
#pragma once
#include <cuda_runtime.h>

namespace m3t {

extern long long g_launch_count;  // defined in capi_misc.cu
inline void count_launch(int n = 1) { __atomic_add_fetch(&g_launch_count, (long long)n, __ATOMIC_RELAXED); }

inline int launch_status() { return cudaGetLastError() == cudaSuccess ? 0 : -21; }

}  // namespace m3t
