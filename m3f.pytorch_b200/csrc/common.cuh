// Shared host-side bits: launch accounting and small helpers.
#pragma once
#include <cuda_runtime.h>

namespace m3t {

extern long long g_launch_count;  // defined in capi_misc.cu
inline void count_launch(int n = 1) { __atomic_add_fetch(&g_launch_count, (long long)n, __ATOMIC_RELAXED); }

inline int launch_status() { return cudaGetLastError() == cudaSuccess ? 0 : -21; }

// Counter-based dropout generator (splitmix64 of seed + (i+1)*golden, top 32 bits): element i is kept iff
// u_i >= thresh.  Stateless: the stand-alone pass (m3t_dropout_bf16), the TemporalBlock conv epilogue and the oracle
// (oracle/dropout.py) all derive the same mask from (seed, element index).
__host__ __device__ __forceinline__ unsigned dropout_u32(unsigned long long seed, long long i) {
  unsigned long long z = seed + (unsigned long long)(i + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (unsigned)(z >> 32);
}

inline unsigned dropout_threshold(float p) {
  const double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 4294967295u : (unsigned)t;
}

}  // namespace m3t
