// Attention fusion mix (reference: models/att_fusion.py:21-25) as one streaming kernel per direction.
//   h_v = sigmoid(s_v), h_a = sigmoid(s_a); (w_v, w_a) = softmax(h_v, h_a); f = w_v*x_v + w_a*x_a
// One warp per (b,t) row; 16-byte vector loads/stores; the backward row reductions (sum_c df*x) are warp shuffles.
// Algorithmic bytes per row (C channels, bf16): fwd 3*2C + 8, bwd 5*2C + 16.
#include "../../include/m3t_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace m3t {

__device__ __forceinline__ void unpack8f(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = bf16lo(w[j]);
    f[2 * j + 1] = bf16hi(w[j]);
  }
}
__device__ __forceinline__ uint4 pack8f(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + __expf(-x)); }

__global__ void att_mix_fwd_kernel(const __nv_bfloat16* __restrict__ xa, const __nv_bfloat16* __restrict__ xv,
                                   const float* __restrict__ sa, const float* __restrict__ sv,
                                   __nv_bfloat16* __restrict__ f, float* __restrict__ wv_out, long long rows, int C) {
  const int lane = threadIdx.x & 31;
  const int nvec = C / 8;
  for (long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; r < rows;
       r += ((long long)gridDim.x * blockDim.x) >> 5) {
    const float hv = sigm(__ldg(sv + r)), ha = sigm(__ldg(sa + r));
    const float wv = sigm(hv - ha), wa = 1.f - wv;
    if (lane == 0 && wv_out) wv_out[r] = wv;
    const uint4* pa = reinterpret_cast<const uint4*>(xa + r * C);
    const uint4* pv = reinterpret_cast<const uint4*>(xv + r * C);
    uint4* pf = reinterpret_cast<uint4*>(f + r * C);
    for (int v = lane; v < nvec; v += 32) {
      float a[8], b[8];
      unpack8f(__ldg(pa + v), a);
      unpack8f(__ldg(pv + v), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = wv * b[j] + wa * a[j];
      pf[v] = pack8f(a);
    }
  }
}

__global__ void att_mix_bwd_kernel(const __nv_bfloat16* __restrict__ df, const __nv_bfloat16* __restrict__ xa,
                                   const __nv_bfloat16* __restrict__ xv, const float* __restrict__ sa,
                                   const float* __restrict__ sv, __nv_bfloat16* __restrict__ dxa,
                                   __nv_bfloat16* __restrict__ dxv, float* __restrict__ dsa, float* __restrict__ dsv,
                                   long long rows, int C) {
  const int lane = threadIdx.x & 31;
  const int nvec = C / 8;
  for (long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; r < rows;
       r += ((long long)gridDim.x * blockDim.x) >> 5) {
    const float hv = sigm(__ldg(sv + r)), ha = sigm(__ldg(sa + r));
    const float wv = sigm(hv - ha), wa = 1.f - wv;
    const uint4* pd = reinterpret_cast<const uint4*>(df + r * C);
    const uint4* pa = reinterpret_cast<const uint4*>(xa + r * C);
    const uint4* pv = reinterpret_cast<const uint4*>(xv + r * C);
    float accv = 0.f, acca = 0.f;
    for (int v = lane; v < nvec; v += 32) {
      float d[8], a[8], b[8], oa[8], ov[8];
      unpack8f(__ldg(pd + v), d);
      unpack8f(__ldg(pa + v), a);
      unpack8f(__ldg(pv + v), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acca += d[j] * a[j];
        accv += d[j] * b[j];
        oa[j] = wa * d[j];
        ov[j] = wv * d[j];
      }
      reinterpret_cast<uint4*>(dxa + r * C)[v] = pack8f(oa);
      reinterpret_cast<uint4*>(dxv + r * C)[v] = pack8f(ov);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      accv += __shfl_xor_sync(0xffffffffu, accv, o);
      acca += __shfl_xor_sync(0xffffffffu, acca, o);
    }
    if (lane == 0) {
      const float dhv = wv * wa * (accv - acca);
      dsv[r] = dhv * hv * (1.f - hv);
      dsa[r] = -dhv * ha * (1.f - ha);
    }
  }
}

}  // namespace m3t

using namespace m3t;

extern "C" int m3t_att_mix_fwd(const void* x_a, const void* x_v, const float* s_a, const float* s_v, void* f,
                               float* w_v, long long rows, int C, void* stream) {
  if (C % 8) return -1;
  long long blocks = (rows * 32 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  att_mix_fwd_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x_a), reinterpret_cast<const __nv_bfloat16*>(x_v), s_a, s_v,
      reinterpret_cast<__nv_bfloat16*>(f), w_v, rows, C);
  count_launch();
  return launch_status();
}

extern "C" int m3t_att_mix_bwd(const void* df, const void* x_a, const void* x_v, const float* s_a, const float* s_v,
                               void* dx_a, void* dx_v, float* ds_a, float* ds_v, long long rows, int C,
                               void* stream) {
  if (C % 8) return -1;
  long long blocks = (rows * 32 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  att_mix_bwd_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(df), reinterpret_cast<const __nv_bfloat16*>(x_a),
      reinterpret_cast<const __nv_bfloat16*>(x_v), s_a, s_v, reinterpret_cast<__nv_bfloat16*>(dx_a),
      reinterpret_cast<__nv_bfloat16*>(dx_v), ds_a, ds_v, rows, C);
  count_launch();
  return launch_status();
}
