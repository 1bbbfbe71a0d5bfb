// Optimiser pass over the flat parameter / gradient arena: global-norm clip (train.py:35 gradient_clip_val=1.0) +
// Adam with coupled L2 weight decay (models/model.py:388-390) in two streaming kernels, no host synchronisation
// (the clip coefficient is computed on the device from the squared norm).  HBM-bound: the step reads p, g, m, v and
// writes p, m, v (28 B per parameter); the norm pass reads g (4 B per parameter).
#include "../../include/m3t_b200.h"
#include "common.cuh"

namespace m3t {

// Block-level sum with a fixed reduction tree (deterministic for a given blockDim).
__device__ __forceinline__ float block_sum_fixed(float acc) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  return s;   // valid in thread 0
}

// Two deterministic passes instead of float atomics: data-parallel replicas must compute bit-identical clip
// coefficients from their (identical, all-reduced) gradients, or their parameters drift apart ulp by ulp.
__global__ void sumsq_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  float acc = 0.f;
  const long long n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(g4 + i);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 * 4; i < n; ++i) acc += g[i] * g[i];
  const float s = block_sum_fixed(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void sumsq_final_kernel(const float* __restrict__ partial, int nblocks, float* __restrict__ out) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  float acc = 0.f;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) acc += partial[i];
  const float s = block_sum_fixed(acc);
  if (threadIdx.x == 0) out[0] = s;
}

__global__ void adam_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                                 float wd, float bc1, float bc2_sqrt, float max_norm, float grad_scale,
                                 const float* __restrict__ gnorm_sq) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  float coef = grad_scale;
  if (max_norm > 0.f) {
    const float total = sqrtf(__ldg(gnorm_sq)) * grad_scale;
    coef *= fminf(1.f, max_norm / (total + 1e-6f));
  }
  const float step = lr / bc1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pi = p[i];
    const float gi = fmaf(wd, pi, g[i] * coef);
    const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);
    const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}

// Device-resident optimiser state (CUDA-graph replays of the training step): hyper = {lr, weight_decay, step, bc1,
// sqrt(bc2)}.  adam_tick advances the step counter and derives the bias corrections on the device, so a captured step
// holds no host scalar that changes between replays; lr / weight decay are rewritten by the host (schedulers) with
// an ordinary copy into the same buffer.
__global__ void adam_tick_kernel(float* __restrict__ hyper, float beta1, float beta2) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const float s = hyper[2] + 1.f;
  hyper[2] = s;
  hyper[3] = 1.f - powf(beta1, s);
  hyper[4] = sqrtf(1.f - powf(beta2, s));
}

__global__ void adam_clip_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                     float* __restrict__ v, long long n, const float* __restrict__ hyper, float beta1,
                                     float beta2, float eps, float max_norm, float grad_scale,
                                     const float* __restrict__ gnorm_sq) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const float lr = __ldg(hyper), wd = __ldg(hyper + 1), bc1 = __ldg(hyper + 3), bc2_sqrt = __ldg(hyper + 4);
  float coef = grad_scale;
  if (max_norm > 0.f) {
    const float total = sqrtf(__ldg(gnorm_sq)) * grad_scale;
    coef *= fminf(1.f, max_norm / (total + 1e-6f));
  }
  const float step = lr / bc1;
  const long long n4 = n / 4;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pi = p4[i], gi = g4[i], mi = m4[i], vi = v4[i];
#define M3T_ADAM1(c)                                                     \
    {                                                                    \
      const float gg = fmaf(wd, pi.c, gi.c * coef);                      \
      mi.c = fmaf(beta1, mi.c, (1.f - beta1) * gg);                      \
      vi.c = fmaf(beta2, vi.c, (1.f - beta2) * gg * gg);                 \
      pi.c = pi.c - step * mi.c / (sqrtf(vi.c) / bc2_sqrt + eps);        \
    }
    M3T_ADAM1(x) M3T_ADAM1(y) M3T_ADAM1(z) M3T_ADAM1(w)
#undef M3T_ADAM1
    p4[i] = pi; m4[i] = mi; v4[i] = vi;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 * 4; i < n; ++i) {
      const float pi = p[i];
      const float gg = fmaf(wd, pi, g[i] * coef);
      const float mi = fmaf(beta1, m[i], (1.f - beta1) * gg);
      const float vi = fmaf(beta2, v[i], (1.f - beta2) * gg * gg);
      m[i] = mi; v[i] = vi;
      p[i] = pi - step * mi / (sqrtf(vi) / bc2_sqrt + eps);
    }
}

}  // namespace m3t

using namespace m3t;

static const int kSumsqMaxBlocks = 148 * 8;

extern "C" long long m3t_sumsq_workspace_floats(void) { return kSumsqMaxBlocks; }

extern "C" int m3t_sumsq_f32(const float* g, long long n, float* out, float* workspace, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!workspace) return -1;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > kSumsqMaxBlocks) blocks = kSumsqMaxBlocks;
  if (blocks < 1) blocks = 1;
  m3t::launch_k(sumsq_partial_kernel, dim3((int)blocks), dim3(256), 0, st, g, n, workspace);
  count_launch();
  m3t::launch_k(sumsq_final_kernel, dim3(1), dim3(256), 0, st, workspace, (int)blocks, out);
  count_launch();
  return launch_status();
}

extern "C" int m3t_adam_clip_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                                  float beta2, float eps, float weight_decay, int step, float max_norm,
                                  float grad_scale, const float* gnorm_sq, void* stream) {
  if (step < 1) return -1;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  m3t::launch_k(adam_clip_kernel, dim3((int)blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, max_norm, grad_scale, gnorm_sq);
  count_launch();
  return launch_status();
}

extern "C" int m3t_adam_clip_step_dev(float* p, const float* g, float* m, float* v, long long n, float* hyper,
                                      float beta1, float beta2, float eps, float max_norm, float grad_scale,
                                      const float* gnorm_sq, void* stream) {
  if (!hyper) return -1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  m3t::launch_k(adam_tick_kernel, dim3(1), dim3(1), 0, st, hyper, beta1, beta2);
  count_launch();
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  m3t::launch_k(adam_clip_dev_kernel, dim3((int)blocks), dim3(256), 0, st, p, g, m, v, n, hyper, beta1, beta2, eps, max_norm, grad_scale,
                                                    gnorm_sq);
  count_launch();
  return launch_status();
}
