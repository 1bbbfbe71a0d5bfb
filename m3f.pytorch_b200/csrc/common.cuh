// Shared host-side bits: launch accounting and small helpers.
#pragma once
#include <cuda_runtime.h>

namespace m3t {

extern long long g_launch_count;  // defined in capi_misc.cu
inline void count_launch(int n = 1) { __atomic_add_fetch(&g_launch_count, (long long)n, __ATOMIC_RELAXED); }

inline int launch_status() { return cudaGetLastError() == cudaSuccess ? 0 : -21; }

// Programmatic dependent launch.  Every kernel of the library starts with pdl_wait() (all memory operations of the
// kernels before it in the stream are complete and visible) followed by pdl_launch() (the NEXT kernel of the stream
// may be scheduled now: its CTAs take SMs as this grid's CTAs retire, run their prologue and block in their own
// pdl_wait() until this grid has completed).  What overlaps is launch latency, CTA scheduling and the prologue before
// pdl_wait() (the tcgen05 kernels put it after barrier init / TMEM allocation / descriptor prefetch).  Without the
// launch attribute both instructions are no-ops: that is the default (M3T_PDL=1 or m3t_set_pdl(1) switch it on).
// Measured on B200: +0.4 ms on the 256-clip step (30.0 -> 30.4 ms: ~400 early-resident grids take slots from the
// streaming kernels' last wave), -0.13 ms on the 32-clip step replayed as a CUDA graph (6.77 -> 6.65 ms), which is
// where engine.capture() turns it on.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// SMs the persistent (one CTA per SM) kernels size their grids for: the device's count minus m3t_set_sm_reserve(k).
// The data-parallel engine reserves k SMs while a bucket of the gradient all-reduce overlaps the backward pass, so
// that NCCL's CTAs find free SMs and a 148-CTA persistent launch does not become a two-wave one.
int usable_sms();

extern int g_pdl;   // capi_misc.cu: -1 = read M3T_PDL on first use
bool pdl_enabled();

template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

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
