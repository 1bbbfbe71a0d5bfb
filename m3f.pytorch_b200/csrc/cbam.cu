// CBAM gates (reference models/cbam.py:32-112) on channels-last bf16 feature maps x[F][S][C] (S = H*W):
//   ChannelGate: (avg, max over S) -> shared MLP -> sigmoid -> x * s_c[f][c]
//   SpatialGate: (max, mean over C) -> 5x5 conv (2 -> 1) -> BatchNorm2d(1) -> sigmoid -> x * s_s[f][s]
// The feature-map-sized passes live here (pooling over S / over C with arg-max bookkeeping, the two broadcast scalings
// and their reductions, the 2-channel 5x5 gate convolution with its data / weight gradients); the per-frame MLP and the
// one-channel BatchNorm act on O(F*C) / O(F*S) values on the host side (models/cbam.py).  All HBM-bound streaming
// passes: 16-byte vectors along C, one warp per pixel row for reductions over C.
#include <cuda_bf16.h>
#include <math_constants.h>

#include "../../include/m3t_b200.h"
#include "common.cuh"

namespace m3t {

__device__ __forceinline__ void cb_unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 cb_pack8(const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

// ---- pooling over S: avg[f][c], mx[f][c], arg[f][c] (first maximum, as torch.max) -------------------------------
__global__ void cbam_pool_hw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ avg,
                                    float* __restrict__ mx, int* __restrict__ arg, int F, int S, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int cgs = C / 8;
  const long long total = (long long)F * cgs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    const long long f = i / cgs;
    float sum[8], best[8];
    int bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sum[j] = 0.f; best[j] = -CUDART_INF_F; bi[j] = 0; }
    const uint4* p = reinterpret_cast<const uint4*>(x + (f * S) * C) + cg;
    for (int s = 0; s < S; ++s) {
      float v[8];
      cb_unpack8(__ldg(p + (long long)s * cgs), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sum[j] += v[j];
        if (v[j] > best[j]) { best[j] = v[j]; bi[j] = s; }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long long o = f * C + cg * 8 + j;
      avg[o] = sum[j] / (float)S;
      mx[o] = best[j];
      arg[o] = bi[j];
    }
  }
}

// dx[f][s][c] = davg[f][c] / S + (s == arg[f][c]) * dmx[f][c]
__global__ void cbam_pool_hw_bwd_kernel(const float* __restrict__ davg, const float* __restrict__ dmx,
                                        const int* __restrict__ arg, __nv_bfloat16* __restrict__ dx, int F, int S,
                                        int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int cgs = C / 8;
  const long long total = (long long)F * S * cgs;
  const float inv = 1.f / (float)S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    const long long r = i / cgs;
    const int s = (int)(r % S);
    const long long f = r / S;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long long o = f * C + cg * 8 + j;
      v[j] = davg[o] * inv + (arg[o] == s ? dmx[o] : 0.f);
    }
    reinterpret_cast<uint4*>(dx)[i] = cb_pack8(v);
  }
}

// ---- y = x * s_c[f][c] ; backward dx = dy * s_c, ds_c[f][c] = sum_s dy * x ----------------------------------------
__global__ void cbam_scale_c_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ sc,
                                    __nv_bfloat16* __restrict__ y, int F, int S, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int cgs = C / 8;
  const long long total = (long long)F * S * cgs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    const long long f = i / cgs / S;
    float v[8];
    cb_unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= __ldg(sc + f * C + cg * 8 + j);
    reinterpret_cast<uint4*>(y)[i] = cb_pack8(v);
  }
}

__global__ void cbam_scale_c_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                        const float* __restrict__ sc, __nv_bfloat16* __restrict__ dx,
                                        float* __restrict__ dsc, int F, int S, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int cgs = C / 8;
  const long long total = (long long)F * cgs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    const long long f = i / cgs;
    float s8[8], acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s8[j] = sc[f * C + cg * 8 + j]; acc[j] = 0.f; }
    const long long base = (f * S) * cgs + cg;
    for (int s = 0; s < S; ++s) {
      float d[8], v[8];
      cb_unpack8(__ldg(reinterpret_cast<const uint4*>(dy) + base + (long long)s * cgs), d);
      cb_unpack8(__ldg(reinterpret_cast<const uint4*>(x) + base + (long long)s * cgs), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] = fmaf(d[j], v[j], acc[j]);
        d[j] *= s8[j];
      }
      reinterpret_cast<uint4*>(dx)[base + (long long)s * cgs] = cb_pack8(d);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) dsc[f * C + cg * 8 + j] = acc[j];
  }
}

// ---- pooling over C, one warp per pixel: comp[f][0][s] = max_c, comp[f][1][s] = mean_c, carg[f][s] ----------------
__global__ void cbam_pool_c_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ comp,
                                   int* __restrict__ carg, int F, int S, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long rows = (long long)F * S;
  for (long long r = warp; r < rows; r += nwarps) {
    float best = -CUDART_INF_F, sum = 0.f;
    int bi = 0;
    for (int c = lane; c < C; c += 32) {
      const float v = __bfloat162float(x[r * C + c]);
      sum += v;
      if (v > best) { best = v; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }      // first maximum, as torch.max
    }
    if (lane == 0) {
      const long long f = r / S, s = r % S;
      comp[(f * 2 + 0) * S + s] = best;
      comp[(f * 2 + 1) * S + s] = sum / (float)C;
      carg[r] = bi;
    }
  }
}

// dx[f][s][c] = dcomp[f][1][s] / C + (c == carg[f][s]) * dcomp[f][0][s]
__global__ void cbam_pool_c_bwd_kernel(const float* __restrict__ dcomp, const int* __restrict__ carg,
                                       __nv_bfloat16* __restrict__ dx, int F, int S, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int cgs = C / 8;
  const long long total = (long long)F * S * cgs;
  const float inv = 1.f / (float)C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    const long long r = i / cgs;
    const long long f = r / S, s = r % S;
    const float dm = dcomp[(f * 2 + 0) * S + s], da = dcomp[(f * 2 + 1) * S + s] * inv;
    const int a = carg[r];
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = da + (cg * 8 + j == a ? dm : 0.f);
    reinterpret_cast<uint4*>(dx)[i] = cb_pack8(v);
  }
}

// ---- y = x * s_s[f][s] ; backward dx = dy * s_s, ds_s[f][s] = sum_c dy * x (one warp per pixel) --------------------
__global__ void cbam_scale_s_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ ss,
                                    __nv_bfloat16* __restrict__ y, long long rows, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int cgs = C / 8;
  const long long total = rows * cgs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float s = __ldg(ss + i / cgs);
    float v[8];
    cb_unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= s;
    reinterpret_cast<uint4*>(y)[i] = cb_pack8(v);
  }
}

__global__ void cbam_scale_s_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                        const float* __restrict__ ss, __nv_bfloat16* __restrict__ dx,
                                        float* __restrict__ dss, long long rows, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const float s = ss[r];
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float d = __bfloat162float(dy[r * C + c]);
      acc = fmaf(d, __bfloat162float(x[r * C + c]), acc);
      dx[r * C + c] = __float2bfloat16(d * s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) dss[r] = acc;
  }
}

// ---- the SpatialGate's conv: in fp32 [F][2][H][W], w [1][2][5][5], pad 2, no bias -> out fp32 [F][1][H][W] ----------
__global__ void cbam_conv5_kernel(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out,
                                  int F, int H, int W) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  __shared__ float ws[50];
  if (threadIdx.x < 50) ws[threadIdx.x] = w[threadIdx.x];
  __syncthreads();
  const long long total = (long long)F * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x0 = (int)(i % W);
    const int y0 = (int)((i / W) % H);
    const long long f = i / ((long long)H * W);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int kh = 0; kh < 5; ++kh) {
        const int yy = y0 + kh - 2;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int kw = 0; kw < 5; ++kw) {
          const int xx = x0 + kw - 2;
          if (xx < 0 || xx >= W) continue;
          acc = fmaf(ws[(c * 5 + kh) * 5 + kw], in[((f * 2 + c) * H + yy) * W + xx], acc);
        }
      }
    out[i] = acc;
  }
}

// din[f][c][y][x] = sum_{kh,kw} w[c][kh][kw] * dout[f][y - kh + 2][x - kw + 2]
__global__ void cbam_conv5_bwd_data_kernel(const float* __restrict__ dout, const float* __restrict__ w,
                                           float* __restrict__ din, int F, int H, int W) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  __shared__ float ws[50];
  if (threadIdx.x < 50) ws[threadIdx.x] = w[threadIdx.x];
  __syncthreads();
  const long long total = (long long)F * 2 * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x0 = (int)(i % W);
    const int y0 = (int)((i / W) % H);
    const int c = (int)((i / ((long long)H * W)) % 2);
    const long long f = i / ((long long)2 * H * W);
    float acc = 0.f;
#pragma unroll
    for (int kh = 0; kh < 5; ++kh) {
      const int yy = y0 - kh + 2;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 5; ++kw) {
        const int xx = x0 - kw + 2;
        if (xx < 0 || xx >= W) continue;
        acc = fmaf(ws[(c * 5 + kh) * 5 + kw], dout[(f * H + yy) * W + xx], acc);
      }
    }
    din[i] = acc;
  }
}

// dw[c][kh][kw] = sum_{f,y,x} dout[f][y][x] * in[f][c][y + kh - 2][x + kw - 2]   (block per tap, fixed-order tree)
__global__ void cbam_conv5_bwd_w_kernel(const float* __restrict__ dout, const float* __restrict__ in,
                                        float* __restrict__ dw, int F, int H, int W) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int tap = blockIdx.x;                 // 0 .. 49
  const int c = tap / 25, kh = (tap / 5) % 5, kw = tap % 5;
  const long long total = (long long)F * H * W;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < total; i += blockDim.x) {
    const int x0 = (int)(i % W);
    const int y0 = (int)((i / W) % H);
    const long long f = i / ((long long)H * W);
    const int yy = y0 + kh - 2, xx = x0 + kw - 2;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    acc = fmaf(dout[i], in[((f * 2 + c) * H + yy) * W + xx], acc);
  }
  __shared__ float red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) dw[tap] = s;
  }
}

// ---- AvgPool3d((1,2,2), stride (1,2,2)) on CL [F][H][W][C] -> [F][H/2][W/2][C] (floor; DenseNet transitions) -----------
__global__ void avgpool2x2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int F, int H,
                                  int W, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int cgs = C / 8, P = H / 2, Q = W / 2;
  const long long total = (long long)F * P * Q * cgs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    long long r = i / cgs;
    const int q = (int)(r % Q);
    r /= Q;
    const int p = (int)(r % P);
    const long long f = r / P;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int dh = 0; dh < 2; ++dh)
#pragma unroll
      for (int dw = 0; dw < 2; ++dw) {
        float v[8];
        cb_unpack8(__ldg(reinterpret_cast<const uint4*>(x + ((f * H + 2 * p + dh) * W + 2 * q + dw) * C) + cg), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= 0.25f;
    reinterpret_cast<uint4*>(y)[i] = cb_pack8(acc);
  }
}

__global__ void avgpool2x2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int F,
                                      int H, int W, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int cgs = C / 8, P = H / 2, Q = W / 2;
  const long long total = (long long)F * H * W * cgs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    long long r = i / cgs;
    const int w = (int)(r % W);
    r /= W;
    const int h = (int)(r % H);
    const long long f = r / H;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (h / 2 < P && w / 2 < Q) {       // an odd last row / column is outside every window
      cb_unpack8(__ldg(reinterpret_cast<const uint4*>(dy + ((f * P + h / 2) * Q + w / 2) * C) + cg), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= 0.25f;
    }
    reinterpret_cast<uint4*>(dx)[i] = cb_pack8(v);
  }
}

static inline int cb_blocks(long long items, int threads = 256) {
  long long b = (items + threads - 1) / threads;
  if (b > 148LL * 16) b = 148LL * 16;
  return b < 1 ? 1 : (int)b;
}

}  // namespace m3t

using namespace m3t;
#define CB_ST(s) reinterpret_cast<cudaStream_t>(s)
#define CB_BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CB_CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" int m3t_cbam_pool_hw(const void* x, float* avg, float* mx, int* arg, int F, int S, int C, void* stream) {
  if (C % 8 || F <= 0 || S <= 0) return -1;
  m3t::launch_k(cbam_pool_hw_kernel, dim3(cb_blocks((long long)F * (C / 8), 128)), dim3(128), 0, CB_ST(stream), CB_CBF(x), avg, mx, arg, F, S, C);
  count_launch();
  return launch_status();
}
extern "C" int m3t_cbam_pool_hw_bwd(const float* davg, const float* dmx, const int* arg, void* dx, int F, int S, int C,
                                    void* stream) {
  if (C % 8) return -1;
  m3t::launch_k(cbam_pool_hw_bwd_kernel, dim3(cb_blocks((long long)F * S * (C / 8))), dim3(256), 0, CB_ST(stream), davg, dmx, arg, CB_BF(dx), F,
                                                                                          S, C);
  count_launch();
  return launch_status();
}
extern "C" int m3t_cbam_scale_c(const void* x, const float* sc, void* y, int F, int S, int C, void* stream) {
  if (C % 8) return -1;
  m3t::launch_k(cbam_scale_c_kernel, dim3(cb_blocks((long long)F * S * (C / 8))), dim3(256), 0, CB_ST(stream), CB_CBF(x), sc, CB_BF(y), F, S, C);
  count_launch();
  return launch_status();
}
extern "C" int m3t_cbam_scale_c_bwd(const void* dy, const void* x, const float* sc, void* dx, float* dsc, int F, int S,
                                    int C, void* stream) {
  if (C % 8) return -1;
  m3t::launch_k(cbam_scale_c_bwd_kernel, dim3(cb_blocks((long long)F * (C / 8), 128)), dim3(128), 0, CB_ST(stream), CB_CBF(dy), CB_CBF(x), sc,
                                                                                           CB_BF(dx), dsc, F, S, C);
  count_launch();
  return launch_status();
}
extern "C" int m3t_cbam_pool_c(const void* x, float* comp, int* carg, int F, int S, int C, void* stream) {
  m3t::launch_k(cbam_pool_c_kernel, dim3(cb_blocks((long long)F * S * 32)), dim3(256), 0, CB_ST(stream), CB_CBF(x), comp, carg, F, S, C);
  count_launch();
  return launch_status();
}
extern "C" int m3t_cbam_pool_c_bwd(const float* dcomp, const int* carg, void* dx, int F, int S, int C, void* stream) {
  if (C % 8) return -1;
  m3t::launch_k(cbam_pool_c_bwd_kernel, dim3(cb_blocks((long long)F * S * (C / 8))), dim3(256), 0, CB_ST(stream), dcomp, carg, CB_BF(dx), F, S,
                                                                                         C);
  count_launch();
  return launch_status();
}
extern "C" int m3t_cbam_scale_s(const void* x, const float* ss, void* y, long long rows, int C, void* stream) {
  if (C % 8) return -1;
  m3t::launch_k(cbam_scale_s_kernel, dim3(cb_blocks(rows * (C / 8))), dim3(256), 0, CB_ST(stream), CB_CBF(x), ss, CB_BF(y), rows, C);
  count_launch();
  return launch_status();
}
extern "C" int m3t_cbam_scale_s_bwd(const void* dy, const void* x, const float* ss, void* dx, float* dss,
                                    long long rows, int C, void* stream) {
  m3t::launch_k(cbam_scale_s_bwd_kernel, dim3(cb_blocks(rows * 32)), dim3(256), 0, CB_ST(stream), CB_CBF(dy), CB_CBF(x), ss, CB_BF(dx), dss,
                                                                          rows, C);
  count_launch();
  return launch_status();
}
extern "C" int m3t_cbam_conv5(const float* in, const float* w, float* out, int F, int H, int W, void* stream) {
  m3t::launch_k(cbam_conv5_kernel, dim3(cb_blocks((long long)F * H * W)), dim3(256), 0, CB_ST(stream), in, w, out, F, H, W);
  count_launch();
  return launch_status();
}
extern "C" int m3t_cbam_conv5_bwd(const float* dout, const float* in, const float* w, float* din, float* dw, int F,
                                  int H, int W, void* stream) {
  m3t::launch_k(cbam_conv5_bwd_data_kernel, dim3(cb_blocks((long long)F * 2 * H * W)), dim3(256), 0, CB_ST(stream), dout, w, din, F, H, W);
  count_launch();
  m3t::launch_k(cbam_conv5_bwd_w_kernel, dim3(50), dim3(1024), 0, CB_ST(stream), dout, in, dw, F, H, W);
  count_launch();
  return launch_status();
}

extern "C" int m3t_avgpool2x2(const void* x, void* y, int F, int H, int W, int C, void* stream) {
  if (C % 8 || H < 2 || W < 2) return -1;
  m3t::launch_k(avgpool2x2_kernel, dim3(cb_blocks((long long)F * (H / 2) * (W / 2) * (C / 8))), dim3(256), 0, CB_ST(stream), CB_CBF(x), CB_BF(y),
                                                                                                     F, H, W, C);
  count_launch();
  return launch_status();
}
extern "C" int m3t_avgpool2x2_bwd(const void* dy, void* dx, int F, int H, int W, int C, void* stream) {
  if (C % 8 || H < 2 || W < 2) return -1;
  m3t::launch_k(avgpool2x2_bwd_kernel, dim3(cb_blocks((long long)F * H * W * (C / 8))), dim3(256), 0, CB_ST(stream), CB_CBF(dy), CB_BF(dx), F, H,
                                                                                            W, C);
  count_launch();
  return launch_status();
}
