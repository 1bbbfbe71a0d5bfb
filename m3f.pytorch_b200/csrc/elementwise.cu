// HBM-bound passes of the visual stream: input preparation, BatchNorm finalize/apply (+residual, +ReLU, +max-pool),
// their backward passes, average pooling, layout/dtype conversion and filter packing.  All activations are
// channels-last bf16; every thread moves 16-byte vectors (8 channels); grids are sized in multiples of the SM count.
#include "../../include/m3t_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace m3t {

constexpr int kEwThreads = 256;
static inline int ew_blocks(long long work_items) {
  long long b = (work_items + kEwThreads - 1) / kEwThreads;
  const long long cap = 148LL * 16;  // persistent-ish: at most 16 CTAs per SM, grid-stride beyond that
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

struct bf16x8 {
  uint4 v;
};
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = bf16lo(w[j]);
    f[2 * j + 1] = bf16hi(w[j]);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// ------------------------------------------------------------------------------------------------------------
// video -> normalised, space-to-depth(2x2), channel-padded bf16:  out[b][t][h2][w2][(ph*2+pw)*3 + c], 12..15 = 0
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void video_prep_s2d_kernel(const T* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int Tn, int H,
                                      int W, float mul, float add) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int H2 = H / 2, W2 = W / 2;
  const long long total = (long long)B * Tn * H2 * W2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w2 = (int)(i % W2);
    long long r = i / W2;
    const int h2 = (int)(r % H2);
    r /= H2;
    const int t = (int)(r % Tn);
    const int b = (int)(r / Tn);
    float f[16];
#pragma unroll
    for (int j = 12; j < 16; ++j) f[j] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int ph = 0; ph < 2; ++ph) {
        const T* p = in + ((((long long)b * 3 + c) * Tn + t) * H + (2 * h2 + ph)) * W + 2 * w2;
        float x0, x1;
        if constexpr (sizeof(T) == 4) {
          const float2 v = __ldg(reinterpret_cast<const float2*>(p));
          x0 = v.x; x1 = v.y;
        } else {
          const uchar2 v = *reinterpret_cast<const uchar2*>(p);
          x0 = (float)v.x; x1 = (float)v.y;
        }
        f[(ph * 2 + 0) * 3 + c] = fmaf(x0, mul, add);
        f[(ph * 2 + 1) * 3 + c] = fmaf(x1, mul, add);
      }
    }
    float lo[8], hi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { lo[j] = f[j]; hi[j] = f[8 + j]; }
    uint4* o = reinterpret_cast<uint4*>(out + i * 16);
    o[0] = pack8(lo);
    o[1] = pack8(hi);
  }
}

// Same normalised space-to-depth image, additionally unrolled over the 4 horizontal filter taps so that every pixel
// carries 64 channels (one 128-byte swizzle row):  out[b][t][h2][w2][jw*12 + ch] = s2d[b][t][h2][w2 + jw - 2][ch]
// for the 12 real channels ch = (ph*2+pw)*3 + c (zero outside the image), channels 48..63 = 0.  The stem conv then is a (5,4,1) filter over 64 channels: the standard CK=64 im2col path.
template <typename T>
__global__ void video_prep_s2d_w4_kernel(const T* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int Tn,
                                         int H, int W, float mul, float add) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  // One thread writes one complete 128-byte destination row (eight 16-byte stores): 4 taps x 12 channels + 16 zeros.
  // Each source pixel pair is read by the four destination pixels it is a tap of (L1-resident re-reads).
  // Measured at 4096 frames (2.26 GB): 0.69 ms = 3.3 TB/s.  Two "coalesced-store" rewrites were tried in round 2 and
  // dropped: rows staged in shared memory (two block barriers per destination row, 2.4x slower) and one thread per
  // 16-byte chunk with eight scalar gathers (2.04 ms, 3x slower) - the float2 loads + full-row stores of this form win.
  const int H2 = H / 2, W2 = W / 2;
  const long long total = (long long)B * Tn * H2 * W2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w2 = (int)(i % W2);
    long long r = i / W2;
    const int h2 = (int)(r % H2);
    r /= H2;
    const int t = (int)(r % Tn);
    const int b = (int)(r / Tn);
    float f[48];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int ph = 0; ph < 2; ++ph) {
        const T* p = in + ((((long long)b * 3 + c) * Tn + t) * H + (2 * h2 + ph)) * W;
#pragma unroll
        for (int jw = 0; jw < 4; ++jw) {
          const int ws = w2 + jw - 2;
          float x0 = 0.f, x1 = 0.f;
          if (ws >= 0 && ws < W2) {
            if constexpr (sizeof(T) == 4) {
              const float2 v = __ldg(reinterpret_cast<const float2*>(p + 2 * ws));
              x0 = fmaf(v.x, mul, add); x1 = fmaf(v.y, mul, add);
            } else {
              const uchar2 v = *reinterpret_cast<const uchar2*>(p + 2 * ws);
              x0 = fmaf((float)v.x, mul, add); x1 = fmaf((float)v.y, mul, add);
            }
          }
          f[jw * 12 + (ph * 2 + 0) * 3 + c] = x0;
          f[jw * 12 + (ph * 2 + 1) * 3 + c] = x1;
        }
      }
    }
    uint4* o = reinterpret_cast<uint4*>(out + i * 64);
#pragma unroll
    for (int g = 0; g < 6; ++g)
      o[g] = make_uint4(pack_bf16x2(f[8 * g], f[8 * g + 1]), pack_bf16x2(f[8 * g + 2], f[8 * g + 3]),
                        pack_bf16x2(f[8 * g + 4], f[8 * g + 5]), pack_bf16x2(f[8 * g + 6], f[8 * g + 7]));
    o[6] = make_uint4(0, 0, 0, 0);   // channels 48..63 are structural zeros (the stem kernels never multiply them)
    o[7] = make_uint4(0, 0, 0, 0);
  }
}

// Input pipeline on the device: the same W-unrolled space-to-depth tensor straight from DECODED
// uint8 frames [B][T][Hs][Ws][3] (HWC, channel order as decoded - the reference keeps cv2's BGR) with the reference's
// per-clip augmentations applied on the fly (models/dataset.py:46-80 load_video, :16-31 sequence_cutout):
//   crop   pixel (y,x) of the H x W clip = frame pixel (crop_y + y, crop_x + x)        (same window for every frame)
//   mirror x -> W-1-x after the crop                                                     (cv2.flip(img, 1))
//   cutout rectangle [cy1,cy2) x [cx1,cx2) of the (cropped, mirrored) clip := 127.5    (zero after normalisation)
// params int32 [B][8] = {crop_x, crop_y, flip, cy1, cy2, cx1, cx2, 0}.
__global__ void video_augment_prep_s2d_w4_kernel(const uint8_t* __restrict__ frames, const int* __restrict__ params,
                                                 __nv_bfloat16* __restrict__ out, int B, int Tn, int Hs, int Ws, int H,
                                                 int W, float mul, float add) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int H2 = H / 2, W2 = W / 2;
  const long long total = (long long)B * Tn * H2 * W2;
  const float fill = fmaf(127.5f, mul, add);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w2 = (int)(i % W2);
    long long r = i / W2;
    const int h2 = (int)(r % H2);
    r /= H2;
    const int t = (int)(r % Tn);
    const int b = (int)(r / Tn);
    const int* pp = params + b * 8;
    const int crop_x = __ldg(pp), crop_y = __ldg(pp + 1), flip = __ldg(pp + 2);
    const int cy1 = __ldg(pp + 3), cy2 = __ldg(pp + 4), cx1 = __ldg(pp + 5), cx2 = __ldg(pp + 6);
    const uint8_t* fr = frames + ((long long)b * Tn + t) * Hs * Ws * 3;
    float f[48];
#pragma unroll
    for (int ph = 0; ph < 2; ++ph) {
      const int y = 2 * h2 + ph;
      const bool ycut = y >= cy1 && y < cy2;
      const uint8_t* row = fr + (long long)(crop_y + y) * Ws * 3;
#pragma unroll
      for (int jw = 0; jw < 4; ++jw) {
        const int ws = w2 + jw - 2;
#pragma unroll
        for (int pw = 0; pw < 2; ++pw) {
          const int x = 2 * ws + pw;
          float v0 = 0.f, v1 = 0.f, v2 = 0.f;
          if (ws >= 0 && ws < W2) {
            if (ycut && x >= cx1 && x < cx2) {
              v0 = v1 = v2 = fill;
            } else {
              const uint8_t* px = row + (crop_x + (flip ? W - 1 - x : x)) * 3;
              v0 = fmaf((float)px[0], mul, add);
              v1 = fmaf((float)px[1], mul, add);
              v2 = fmaf((float)px[2], mul, add);
            }
          }
          f[jw * 12 + (ph * 2 + pw) * 3 + 0] = v0;
          f[jw * 12 + (ph * 2 + pw) * 3 + 1] = v1;
          f[jw * 12 + (ph * 2 + pw) * 3 + 2] = v2;
        }
      }
    }
    uint4* o = reinterpret_cast<uint4*>(out + i * 64);
#pragma unroll
    for (int g = 0; g < 6; ++g)
      o[g] = make_uint4(pack_bf16x2(f[8 * g], f[8 * g + 1]), pack_bf16x2(f[8 * g + 2], f[8 * g + 3]),
                        pack_bf16x2(f[8 * g + 4], f[8 * g + 5]), pack_bf16x2(f[8 * g + 6], f[8 * g + 7]));
    o[6] = make_uint4(0, 0, 0, 0);
    o[7] = make_uint4(0, 0, 0, 0);
  }
}

// ------------------------------------------------------------------------------------------------------------
// BatchNorm finalize: (sum, sumsq) -> mean, invstd, scale = gamma*invstd, shift = beta - mean*scale; running stats
// ------------------------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const float* __restrict__ stats, int C, float count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                   float* __restrict__ scale_out, float* __restrict__ shift_out) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = stats[c] / count;
  float var = stats[C + c] / count - mean * mean;
  var = fmaxf(var, 0.f);
  const float invstd = rsqrtf(var + eps);
  const float sc = gamma[c] * invstd;
  mean_out[c] = mean;
  invstd_out[c] = invstd;
  scale_out[c] = sc;
  shift_out[c] = beta[c] - mean * sc;
  if (running_mean) {
    const float unbiased = count > 1.f ? var * (count / (count - 1.f)) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
}

// eval-mode fold: scale = gamma / sqrt(running_var + eps), shift = beta - running_mean * scale (+ conv bias * scale)
__global__ void bn_fold_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ running_mean, const float* __restrict__ running_var,
                               const float* __restrict__ conv_bias, float eps, float* __restrict__ scale_out,
                               float* __restrict__ shift_out) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] * rsqrtf(running_var[c] + eps);
  const float b = conv_bias ? conv_bias[c] : 0.f;
  scale_out[c] = sc;
  shift_out[c] = beta[c] + (b - running_mean[c]) * sc;
}

// ------------------------------------------------------------------------------------------------------------
// out = act( y*scale + shift + (res*res_scale + res_shift) )      [rows][C] bf16, C % 8 == 0
// ------------------------------------------------------------------------------------------------------------
__global__ void bn_act_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                              const float* __restrict__ shift, const __nv_bfloat16* __restrict__ res,
                              const float* __restrict__ res_scale, const float* __restrict__ res_shift, int relu,
                              __nv_bfloat16* __restrict__ out, long long nvec, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i * 8) % C);
    float v[8], s[8], b[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(y) + i), v);
    load8f(scale + c, s);
    load8f(shift + c, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], s[j], b[j]);
    if (res) {
      float r[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(res) + i), r);
      if (res_scale) {
        load8f(res_scale + c, s);
        load8f(res_shift + c, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = fmaf(r[j], s[j], b[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += r[j];
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    reinterpret_cast<uint4*>(out)[i] = pack8(v);
  }
}

// Same, for the common case that every thread keeps one channel group over its whole grid-stride loop (total threads
// a multiple of C/8): the per-channel constants live in registers and two vectors are in flight per iteration.
template <bool HAS_RES, bool RELU>
__global__ void __launch_bounds__(256, 4)
bn_act_fixed_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                    const float* __restrict__ shift, const __nv_bfloat16* __restrict__ res,
                    __nv_bfloat16* __restrict__ out, long long nvec, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int c = (int)((i * 8) % C);
  float s[8], b[8];
  load8f(scale + c, s);
  load8f(shift + c, b);
  for (; i < nvec; i += 2 * stride) {
    const long long i2 = i + stride;
    const bool two = i2 < nvec;
    uint4 ya = __ldg(reinterpret_cast<const uint4*>(y) + i), yb = make_uint4(0, 0, 0, 0);
    uint4 ra = make_uint4(0, 0, 0, 0), rb = make_uint4(0, 0, 0, 0);
    if (two) yb = __ldg(reinterpret_cast<const uint4*>(y) + i2);
    if (HAS_RES) {
      ra = __ldg(reinterpret_cast<const uint4*>(res) + i);
      if (two) rb = __ldg(reinterpret_cast<const uint4*>(res) + i2);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      float v[8], r[8];
      unpack8(h ? yb : ya, v);
      if (HAS_RES) unpack8(h ? rb : ra, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float x = fmaf(v[j], s[j], b[j]);
        if (HAS_RES) x += r[j];
        v[j] = RELU ? fmaxf(x, 0.f) : x;
      }
      reinterpret_cast<uint4*>(out)[h ? i2 : i] = pack8(v);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Backward of  out = act(y*scale + shift + resid)  with train-mode BN (batch statistics).
//   dz = dout * (out > 0)                                  (relu)   |  dz = dout   (no relu)
//   pass 1: sums[0][c] += dz, sums[1][c] += dz * xhat,  xhat = (y - mean) * invstd ; optionally writes dz (bf16)
//   pass 2: dy = scale * (dz - sums0/n - xhat * sums1/n)
// `dz_in` (pass 1 input override): when the same dz feeds a second BN (downsample branch) it is read, not recomputed.
// ------------------------------------------------------------------------------------------------------------
// relu: 0 = none, 1 = mask from the stored activation (out > 0), 2 = mask recomputed from y*scale + shift > 0 (units
// without a residual input: `out` is then neither stored for backward nor read here)
template <int relu, bool DET = false>
__global__ void __launch_bounds__(256, relu == 2 ? 2 : 3)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ out,
                     const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                     const float* __restrict__ invstd, const float* __restrict__ scale,
                     const float* __restrict__ shift, __nv_bfloat16* __restrict__ dz_out, float* __restrict__ sums,
                     long long rows, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  // block handles a strip of rows for all channels: thread -> (channel group cg = tid % (C/8), row lane)
  extern __shared__ float sh[];  // [2][C]
  const int cgs = C / 8;
  const int rows_per_iter = blockDim.x / cgs;
  const int cg = threadIdx.x % cgs;
  const int rl = threadIdx.x / cgs;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  float a0[8], a1[8], mu[8], is[8], sc[8], sft[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a0[j] = a1[j] = sc[j] = sft[j] = 0.f;
  load8f(mean + cg * 8, mu);
  load8f(invstd + cg * 8, is);
  if (relu == 2) {
    load8f(scale + cg * 8, sc);
    load8f(shift + cg * 8, sft);
  }
  if (rl < rows_per_iter) {
    // two rows in flight per iteration; a1 accumulates sum dz*y and is centred once at the end
    const long long rstride = (long long)gridDim.x * rows_per_iter;
    for (long long r = (long long)blockIdx.x * rows_per_iter + rl; r < rows; r += 2 * rstride) {
      const long long i = r * cgs + cg;
      const bool two = r + rstride < rows;
      const long long i2 = two ? (r + rstride) * cgs + cg : i;
      const uint4 da = __ldg(reinterpret_cast<const uint4*>(dout) + i);
      const uint4 ya = __ldg(reinterpret_cast<const uint4*>(y) + i);
      const uint4 db = __ldg(reinterpret_cast<const uint4*>(dout) + i2);
      const uint4 yb = __ldg(reinterpret_cast<const uint4*>(y) + i2);
      uint4 oa = make_uint4(0, 0, 0, 0), ob = oa;
      if (relu == 1) {
        oa = __ldg(reinterpret_cast<const uint4*>(out) + i);
        ob = __ldg(reinterpret_cast<const uint4*>(out) + i2);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && !two) break;
        float d[8], yy[8];
        unpack8(h ? db : da, d);
        unpack8(h ? yb : ya, yy);
        if (relu == 1) {
          float o[8];
          unpack8(h ? ob : oa, o);
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] = o[j] > 0.f ? d[j] : 0.f;
        } else if (relu == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] = fmaf(yy[j], sc[j], sft[j]) > 0.f ? d[j] : 0.f;
        }
        if (dz_out) reinterpret_cast<uint4*>(dz_out)[h ? i2 : i] = pack8(d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a0[j] += d[j];
          a1[j] = fmaf(d[j], yy[j], a1[j]);
        }
      }
    }
    if constexpr (DET) {
      // deterministic: per-thread partials side by side in shared memory ([row lane][2C]), summed in row-lane order,
      // stored (not added) into this block's private slot of `sums` (slot 1 + blockIdx.x; m3t_det_reduce follows).
      // (the zero-fill of sh[0 .. 2C) above is separated from these writes by the barrier at the top)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sh[rl * 2 * C + cg * 8 + j] = a0[j];
        sh[rl * 2 * C + C + cg * 8 + j] = (a1[j] - mu[j] * a0[j]) * is[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&sh[cg * 8 + j], a0[j]);
        atomicAdd(&sh[C + cg * 8 + j], (a1[j] - mu[j] * a0[j]) * is[j]);
      }
    }
  }
  __syncthreads();
  if constexpr (DET) {
    float* slot = sums + (1 + (long long)blockIdx.x) * 2 * C;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < rows_per_iter; ++r) s += sh[r * 2 * C + i];
      slot[i] = s;
    }
  } else {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(sums + i, sh[i]);
  }
}

__global__ void bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ out,
                                    const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, const float* __restrict__ scale,
                                    const float* __restrict__ shift, const float* __restrict__ sums, float inv_count,
                                    int relu, __nv_bfloat16* __restrict__ dy, long long nvec, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i * 8) % C);
    float d[8], yy[8], mu[8], is[8], sc[8], s0[8], s1[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dout) + i), d);
    unpack8(__ldg(reinterpret_cast<const uint4*>(y) + i), yy);
    load8f(scale + c, sc);
    if (relu == 1) {
      float o[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(out) + i), o);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = o[j] > 0.f ? d[j] : 0.f;
    } else if (relu == 2) {
      float sh[8];
      load8f(shift + c, sh);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = fmaf(yy[j], sc[j], sh[j]) > 0.f ? d[j] : 0.f;
    }
    load8f(mean + c, mu);
    load8f(invstd + c, is);
    load8f(sums + c, s0);
    load8f(sums + C + c, s1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (yy[j] - mu[j]) * is[j];
      d[j] = sc[j] * (d[j] - s0[j] * inv_count - xh * s1[j] * inv_count);
    }
    reinterpret_cast<uint4*>(dy)[i] = pack8(d);
  }
}

// Fixed-channel-group variant (total threads a multiple of C/8): dy = kA*dz + kB*y + kC with the three per-channel
// coefficients in registers, two vectors in flight per iteration.
template <int RELU>
__global__ void __launch_bounds__(256, 4)
bn_bwd_apply_fixed_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ out,
                          const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                          const float* __restrict__ invstd, const float* __restrict__ scale,
                          const float* __restrict__ shift, const float* __restrict__ sums, float inv_count,
                          __nv_bfloat16* __restrict__ dy, long long nvec, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int c = (int)((i * 8) % C);
  float kA[8], kB[8], kC[8], sh[8];
  {
    float mu[8], is[8], s0[8], s1[8];
    load8f(scale + c, kA);
    load8f(mean + c, mu);
    load8f(invstd + c, is);
    load8f(sums + c, s0);
    load8f(sums + C + c, s1);
    if (RELU == 2) load8f(shift + c, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      kB[j] = -kA[j] * is[j] * s1[j] * inv_count;
      kC[j] = -kA[j] * s0[j] * inv_count - kB[j] * mu[j];
    }
  }
  for (; i < nvec; i += 2 * stride) {
    const long long i2 = i + stride;
    const bool two = i2 < nvec;
    const uint4 z4 = make_uint4(0, 0, 0, 0);
    const uint4 da = __ldg(reinterpret_cast<const uint4*>(dout) + i);
    const uint4 ya = __ldg(reinterpret_cast<const uint4*>(y) + i);
    const uint4 db = two ? __ldg(reinterpret_cast<const uint4*>(dout) + i2) : z4;
    const uint4 yb = two ? __ldg(reinterpret_cast<const uint4*>(y) + i2) : z4;
    uint4 oa = z4, ob = z4;
    if (RELU == 1) {
      oa = __ldg(reinterpret_cast<const uint4*>(out) + i);
      if (two) ob = __ldg(reinterpret_cast<const uint4*>(out) + i2);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      float d[8], yy[8];
      unpack8(h ? db : da, d);
      unpack8(h ? yb : ya, yy);
      if (RELU == 1) {
        float o[8];
        unpack8(h ? ob : oa, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = o[j] > 0.f ? d[j] : 0.f;
      } else if (RELU == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = fmaf(yy[j], kA[j], sh[j]) > 0.f ? d[j] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = fmaf(kA[j], d[j], fmaf(kB[j], yy[j], kC[j]));
      reinterpret_cast<uint4*>(dy)[h ? i2 : i] = pack8(d);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Stem tail: out = maxpool_{3x3,s2,p1}( relu(y*scale + shift) ) over (H,W) of [F][H][W][C]; idx = argmax (0..8)
// ------------------------------------------------------------------------------------------------------------
template <int K, int S, int PAD>
__global__ void bn_relu_maxpool_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                                       const float* __restrict__ shift, __nv_bfloat16* __restrict__ out,
                                       uint8_t* __restrict__ idx, __nv_bfloat16* __restrict__ ymax, int F, int H, int W,
                                       int C, int cg_shift) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  // index math in 32 bits (the host checks the element counts fit); C/8 is a power of two on this path
  const int P = (H + 2 * PAD - K) / S + 1, Q = (W + 2 * PAD - K) / S + 1;
  const unsigned cgs = (unsigned)C / 8;
  const unsigned total = (unsigned)F * P * Q * cgs;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int cg = (int)(i & (cgs - 1));
    unsigned r = i >> cg_shift;
    const int q = (int)(r % (unsigned)Q);
    r /= (unsigned)Q;
    const int p = (int)(r % (unsigned)P);
    const int f = (int)(r / (unsigned)P);
    float s[8], b[8], best[8], yb[8];
    int bi[8];
    load8f(scale + cg * 8, s);
    load8f(shift + cg * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0; yb[j] = 0.f; }
#pragma unroll
    for (int kh = 0; kh < K; ++kh) {
      const int h = S * p - PAD + kh;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int kw = 0; kw < K; ++kw) {
        const int w = S * q - PAD + kw;
        if (w < 0 || w >= W) continue;
        float v[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(y + (((long long)f * H + h) * W + w) * C) + cg), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = fmaxf(fmaf(v[j], s[j], b[j]), 0.f);
          if (a > best[j]) { best[j] = a; bi[j] = kh * K + kw; yb[j] = v[j]; }
        }
      }
    }
    reinterpret_cast<uint4*>(out)[i] = pack8(best);
    if (ymax) reinterpret_cast<uint4*>(ymax)[i] = pack8(yb);     // raw conv output at the arg-max (BN backward sums)
    if (idx) {
      uint2 pk;
      pk.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
      pk.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
      reinterpret_cast<uint2*>(idx)[i] = pk;
    }
  }
}

// Backward of the stem tail.  dz[f][h][w][c] = relu'(a) * sum_{windows (p,q) containing (h,w) with argmax == (h,w)}
// dout[f][p][q][c];  MODE 0: accumulate (sum dz, sum dz*xhat) ;  MODE 1: write dy = scale*(dz - s0/n - xhat*s1/n).
template <int MODE, int K, int S, int PAD>
__global__ void __launch_bounds__(256, 2) maxpool_bn_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const uint8_t* __restrict__ idx,
                                      const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                                      const float* __restrict__ invstd, const float* __restrict__ scale,
                                      const float* __restrict__ shift, float* __restrict__ sums, float inv_count,
                                      __nv_bfloat16* __restrict__ dy, int F, int H, int W, int C, int cg_shift) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int P = (H + 2 * PAD - K) / S + 1, Q = (W + 2 * PAD - K) / S + 1;
  const int cgs = C / 8;
  extern __shared__ float sh[];  // MODE 0: [2][C]
  if (MODE == 0) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
  }
  const int cg = threadIdx.x % cgs;        // blockDim.x % cgs == 0 -> a thread keeps its channel group
  float s[8], b[8], mu[8], is[8], s0[8], s1[8], a0[8], a1[8];
  load8f(scale + cg * 8, s);
  load8f(shift + cg * 8, b);
  load8f(mean + cg * 8, mu);
  load8f(invstd + cg * 8, is);
  if (MODE == 1) {
    load8f(sums + cg * 8, s0);
    load8f(sums + C + cg * 8, s1);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) a0[j] = a1[j] = 0.f;
  const unsigned total = (unsigned)F * H * W * cgs;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    unsigned r = i >> cg_shift;
    const int w = (int)(r % (unsigned)W);
    r /= (unsigned)W;
    const int h = (int)(r % (unsigned)H);
    const int f = (int)(r / (unsigned)H);
    float yy[8], dz[8];
    const uint4 yraw = __ldg(reinterpret_cast<const uint4*>(y) + i);
#pragma unroll
    for (int j = 0; j < 8; ++j) dz[j] = 0.f;
    // windows p with S*p-PAD <= h <= S*p-PAD+K-1  ->  p in [ceil((h+PAD-K+1)/S), floor((h+PAD)/S)]; at most two per
    // dimension for the supported pools (K <= 3, S == 2 or K == S).  All (idx, dout) loads are issued before use.
    const int hn = h + PAD - K + 1, wn = w + PAD - K + 1;
    const int p_lo = hn > 0 ? (hn + S - 1) / S : 0, p_hi = min((h + PAD) / S, P - 1);
    const int q_lo = wn > 0 ? (wn + S - 1) / S : 0, q_hi = min((w + PAD) / S, Q - 1);
    uint2 pk[4];
    uint4 dv[4];
    int me[4];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
      for (int bq = 0; bq < 2; ++bq) {
        const int p = p_lo + a, q = q_lo + bq;
        const bool ok = p <= p_hi && q <= q_hi;
        const int pp = ok ? p : p_lo, qq = ok ? q : q_lo;
        const unsigned o = ((((unsigned)f * P + min(pp, P - 1)) * Q + min(qq, Q - 1)) << cg_shift) + cg;
        me[a * 2 + bq] = ok ? (h - (S * p - PAD)) * K + (w - (S * q - PAD)) : -1;
        pk[a * 2 + bq] = __ldg(reinterpret_cast<const uint2*>(idx) + o);
        dv[a * 2 + bq] = __ldg(reinterpret_cast<const uint4*>(dout) + o);
      }
    }
    unpack8(yraw, yy);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float d[8];
      unpack8(dv[k], d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int sel = ((j < 4 ? pk[k].x : pk[k].y) >> (8 * (j & 3))) & 0xFF;
        if (sel == me[k]) dz[j] += d[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float act = fmaf(yy[j], s[j], b[j]);
      if (!(act > 0.f)) dz[j] = 0.f;
    }
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a0[j] += dz[j];
        a1[j] += dz[j] * (yy[j] - mu[j]) * is[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (yy[j] - mu[j]) * is[j];
        dz[j] = s[j] * (dz[j] - s0[j] * inv_count - xh * s1[j] * inv_count);
      }
      reinterpret_cast<uint4*>(dy)[i] = pack8(dz);
    }
  }
  if (MODE == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sh[cg * 8 + j], a0[j]);
      atomicAdd(&sh[C + cg * 8 + j], a1[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(sums + i, sh[i]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// 3x3 / stride 2 / pad 1 specialisations for even H, W (the ResNet stem tail, 56x56 -> 28x28).  The generic
// kernels above are instruction-bound (~320 instructions per 8 channels of one input position: index divisions and
// 4 windows x 8 byte-compares each).  Here a thread owns a 2x2 input block x 8 channels: the block touches exactly
// the four windows (a..a+1, b..b+1), each (idx, dout) pair is loaded once, and which argmax code of which window
// names which of the four positions is a compile-time constant (9 compare sets instead of 16).
// ------------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256, 2)
maxpool3s2_bn_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const uint8_t* __restrict__ idx,
                         const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const float* __restrict__ scale,
                         const float* __restrict__ shift, float* __restrict__ sums, float inv_count,
                         __nv_bfloat16* __restrict__ dy, int F, int H, int W, int C, int cg_shift) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int P = H / 2, Q = W / 2;
  const int cgs = C / 8;
  extern __shared__ float sh[];  // MODE 0: [2][C]
  if (MODE == 0) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
  }
  const int cg = threadIdx.x % cgs;
  // MODE 0: a0 = sum dz, a1 = sum dz*y (centred at the end).  MODE 1: dy = kA*dz + kB*y + kC.
  float s[8], b[8], a0[8], a1[8];
  load8f(scale + cg * 8, s);
  load8f(shift + cg * 8, b);
  if (MODE == 1) {
    float mu[8], is[8], s0[8], s1[8];
    load8f(mean + cg * 8, mu);
    load8f(invstd + cg * 8, is);
    load8f(sums + cg * 8, s0);
    load8f(sums + C + cg * 8, s1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float kb = -s[j] * is[j] * s1[j] * inv_count;     // coefficient of (y - mu)
      a0[j] = kb;
      a1[j] = -s[j] * s0[j] * inv_count - kb * mu[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) a0[j] = a1[j] = 0.f;
  }
  const unsigned total = (unsigned)F * P * Q * cgs;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    unsigned r = i >> cg_shift;
    const int bq = (int)(r % (unsigned)Q);
    r /= (unsigned)Q;
    const int ap = (int)(r % (unsigned)P);
    const int f = (int)(r / (unsigned)P);
    // windows (ap,bq) (ap,bq+1) (ap+1,bq) (ap+1,bq+1); missing ones get an argmax code that matches nothing
    const bool okp = ap + 1 < P, okq = bq + 1 < Q;
    const unsigned o00 = i;
    const unsigned o01 = okq ? i + cgs : i;
    const unsigned o10 = okp ? i + (unsigned)Q * cgs : i;
    const unsigned o11 = okp ? o01 + (unsigned)Q * cgs : o01;
    uint2 pk[4];
    uint4 dv[4], yr[4];
    pk[0] = __ldg(reinterpret_cast<const uint2*>(idx) + o00);
    pk[1] = __ldg(reinterpret_cast<const uint2*>(idx) + o01);
    pk[2] = __ldg(reinterpret_cast<const uint2*>(idx) + o10);
    pk[3] = __ldg(reinterpret_cast<const uint2*>(idx) + o11);
    dv[0] = __ldg(reinterpret_cast<const uint4*>(dout) + o00);
    dv[1] = __ldg(reinterpret_cast<const uint4*>(dout) + o01);
    dv[2] = __ldg(reinterpret_cast<const uint4*>(dout) + o10);
    dv[3] = __ldg(reinterpret_cast<const uint4*>(dout) + o11);
    const unsigned y00 = ((((unsigned)f * H + 2 * ap) * W + 2 * bq) << cg_shift) + cg;
    const unsigned yoff[4] = {y00, y00 + cgs, y00 + (unsigned)W * cgs, y00 + (unsigned)W * cgs + cgs};
#pragma unroll
    for (int k = 0; k < 4; ++k) yr[k] = __ldg(reinterpret_cast<const uint4*>(y) + yoff[k]);
    if (!okq) { pk[1] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu); }
    if (!okp) { pk[2] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu); }
    if (!okp || !okq) { pk[3] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu); }
    float d[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k) unpack8(dv[k], d[k]);
    // argmax code kh*3+kw of window k that names position (ph,pw) of the block; -1: the window does not cover it
    constexpr int code[4][4] = {{4, -1, -1, -1}, {5, 3, -1, -1}, {7, -1, 1, -1}, {8, 6, 2, 0}};
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) {
      float dz[8], yy[8];
      unpack8(yr[pos], yy);
#pragma unroll
      for (int j = 0; j < 8; ++j) dz[j] = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (code[pos][k] < 0) continue;
        const uint32_t m4 = (uint32_t)code[pos][k] * 0x01010101u;
        const uint32_t x0 = pk[k].x ^ m4, x1 = pk[k].y ^ m4;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t xx = j < 4 ? x0 : x1;
          if ((xx & (0xFFu << (8 * (j & 3)))) == 0u) dz[j] += d[k][j];
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float act = fmaf(yy[j], s[j], b[j]);
        if (!(act > 0.f)) dz[j] = 0.f;
      }
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a0[j] += dz[j];
          a1[j] = fmaf(dz[j], yy[j], a1[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) dz[j] = fmaf(s[j], dz[j], fmaf(a0[j], yy[j], a1[j]));
        reinterpret_cast<uint4*>(dy)[yoff[pos]] = pack8(dz);
      }
    }
  }
  if (MODE == 0) {
    float mu[8], is[8];
    load8f(mean + cg * 8, mu);
    load8f(invstd + cg * 8, is);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sh[cg * 8 + j], a0[j]);
      atomicAdd(&sh[C + cg * 8 + j], (a1[j] - mu[j] * a0[j]) * is[j]);   // sum dz*xhat
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(sums + i, sh[i]);
  }
}

// Forward, same specialisation: all nine window loads are issued before the compare chain.
__global__ void __launch_bounds__(256, 2)
bn_relu_maxpool3s2_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                          const float* __restrict__ shift, __nv_bfloat16* __restrict__ out, uint8_t* __restrict__ idx,
                          __nv_bfloat16* __restrict__ ymax, int F, int H, int W, int C, int cg_shift) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int P = H / 2, Q = W / 2;
  const unsigned cgs = (unsigned)C / 8;
  const unsigned total = (unsigned)F * P * Q * cgs;
  const int cg = threadIdx.x % cgs;
  float s[8], b[8];
  load8f(scale + cg * 8, s);
  load8f(shift + cg * 8, b);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    unsigned r = i >> cg_shift;
    const int q = (int)(r % (unsigned)Q);
    r /= (unsigned)Q;
    const int p = (int)(r % (unsigned)P);
    const int f = (int)(r / (unsigned)P);
    // rows 2p-1..2p+1, cols 2q-1..2q+1: only the -1 row / column can fall outside (H, W even)
    const unsigned c00 = ((((unsigned)f * H + 2 * p) * W + 2 * q) << cg_shift) + cg;   // window centre (kh=kw=1)
    const unsigned rs = (unsigned)W * cgs;
    uint4 v[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const bool ok = (kh > 0 || p > 0) && (kw > 0 || q > 0);
        const unsigned o = c00 + (unsigned)(kh - 1) * rs + (unsigned)(kw - 1) * cgs;
        v[kh * 3 + kw] = __ldg(reinterpret_cast<const uint4*>(y) + (ok ? o : c00));
      }
    }
    float best[8], yb[8];
    int bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0; yb[j] = 0.f; }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const bool ok = (k / 3 > 0 || p > 0) && (k % 3 > 0 || q > 0);
      float x[8];
      unpack8(v[k], x);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = ok ? fmaxf(fmaf(x[j], s[j], b[j]), 0.f) : -INFINITY;
        if (a > best[j]) { best[j] = a; bi[j] = k; yb[j] = x[j]; }
      }
    }
    reinterpret_cast<uint4*>(out)[i] = pack8(best);
    if (ymax) reinterpret_cast<uint4*>(ymax)[i] = pack8(yb);     // raw conv output at the arg-max (BN backward sums)
    if (idx) {
      uint2 pk;
      pk.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
      pk.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
      reinterpret_cast<uint2*>(idx)[i] = pk;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Global average pool over HW of [F][HW][C] bf16 -> [F][C] (bf16 and/or fp32); and its backward.
// ------------------------------------------------------------------------------------------------------------
__global__ void avgpool_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out_bf16,
                               float* __restrict__ out_f32, int F, int HW, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int cgs = C / 8;
  const long long total = (long long)F * cgs;
  const float inv = 1.f / HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    const long long f = i / cgs;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int s = 0; s < HW; ++s) {
      float v[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + (f * HW + s) * C) + cg), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    if (out_bf16) reinterpret_cast<uint4*>(out_bf16)[i] = pack8(acc);
    if (out_f32) {
      float4* o = reinterpret_cast<float4*>(out_f32 + i * 8);
      o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}
// dx[f][s][c] = dout[f][c] / HW   (dout fp32 or bf16 -> dx bf16)
template <typename T>
__global__ void avgpool_bwd_kernel(const T* __restrict__ dout, __nv_bfloat16* __restrict__ dx, int F, int HW, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int cgs = C / 8;
  const long long total = (long long)F * HW * cgs;
  const float inv = 1.f / HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    const long long f = i / cgs / HW;
    float v[8];
    if constexpr (sizeof(T) == 4) load8f(reinterpret_cast<const float*>(dout) + f * C + cg * 8, v);
    else unpack8(__ldg(reinterpret_cast<const uint4*>(dout + f * C) + cg), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= inv;
    reinterpret_cast<uint4*>(dx)[i] = pack8(v);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Layout / dtype conversion.  [N][C][S] fp32  <->  [N][S][C] bf16 through a 32x32 smem transpose tile.
// ------------------------------------------------------------------------------------------------------------
__global__ void ncs_f32_to_nsc_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int C, int S,
                                           int Cpad) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 32, s0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, s = s0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && s < S) ? in[((long long)n * C + c) * S + s] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int s = s0 + j, c = c0 + threadIdx.x;
    if (s < S && c < Cpad) out[((long long)n * S + s) * Cpad + c] = __float2bfloat16(tile[threadIdx.x][j]);
  }
}
template <typename TIN>
__global__ void nsc_to_ncs_f32_kernel(const TIN* __restrict__ in, float* __restrict__ out, int C, int S, int Cpad) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 32, s0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int s = s0 + j, c = c0 + threadIdx.x;
    float v = 0.f;
    if (s < S && c < C) {
      if constexpr (sizeof(TIN) == 4) v = in[((long long)n * S + s) * Cpad + c];
      else v = __bfloat162float(in[((long long)n * S + s) * Cpad + c]);
    }
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, s = s0 + threadIdx.x;
    if (c < C && s < S) out[((long long)n * C + c) * S + s] = tile[threadIdx.x][j];
  }
}

// rows x cols fp32 (row stride ld_in) -> bf16 (row stride ld_out >= cols, pad columns zero-filled)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, long long ld_in, __nv_bfloat16* __restrict__ out,
                                     long long ld_out, long long rows, int cols) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const long long total = rows * ld_out;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ld_out;
    const int c = (int)(i - r * ld_out);
    out[i] = __float2bfloat16(c < cols ? in[r * ld_in + c] : 0.f);
  }
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, long long ld_in, float* __restrict__ out,
                                     long long ld_out, long long rows, int cols) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    out[r * ld_out + c] = __bfloat162float(in[r * ld_in + c]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Filter packing:  w fp32 [Cout][Cin][taps]  ->  fprop pack bf16 [Cout][taps][Cin]
//                                              dgrad pack bf16 [Cin][taps (flipped)][Cout]     (stride-1 dgrad)
// and the inverse for gradients: dw_packed fp32 [Cout][taps][Cin] -> dw fp32 [Cout][Cin][taps].
// ------------------------------------------------------------------------------------------------------------
__global__ void pack_filter_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                   __nv_bfloat16* __restrict__ wd, int Cout, int Cin, int taps) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const long long total = (long long)Cout * Cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    long long r = i / taps;
    const int ci = (int)(r % Cin);
    const int co = (int)(r / Cin);
    const __nv_bfloat16 v = __float2bfloat16(w[i]);
    if (wf) wf[((long long)co * taps + t) * Cin + ci] = v;
    if (wd) wd[((long long)ci * taps + (taps - 1 - t)) * Cout + co] = v;
  }
}
// All convolution filters of a model in ONE launch (table-driven): for entry e and element (co, ci, t) of its fp32
// master [Cout][Cin][taps] write the fprop pack, the flipped dgrad pack and - for stride-2 convolutions whose data
// gradient runs as parity sub-convolutions - the sub-filter of the output parity that tap t belongs to.
// Tiling of the batched pack: an entry is cut into tiles of 32 output channels x TCI input channels x all taps
// (TCI = min(64, 288 / taps), at least 1), a plain fp32 copy (has_parity < 0) into tiles of 2048 values; entry.start is
// the index of its first tile and the launch has one block per tile (m3t_b200.h restates the formula for callers).
constexpr int kPackTCO = 32;
constexpr int kPackCopy = 2048;
__host__ __device__ inline int pack_tci(int taps) {
  int t = 288 / taps;
  return t > 64 ? 64 : (t < 1 ? 1 : t);
}
__global__ void pack_filters_batched_kernel(const m3t_pack_entry* __restrict__ tab, int n, long long total_tiles) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  // One block per tile: the fp32 master tile is read once, coalesced (rows of TCI * taps contiguous floats), rounded
  // to bf16 into shared memory, and written out twice from there: the FPROP pack [co][t][ci] (consecutive threads =
  // consecutive input channels) and the flipped DGRAD pack [ci][tf][co] with the parity sub-filter of its tap
  // (consecutive threads = consecutive output channels).  The earlier versions gathered 4-byte values straight from
  // global memory in destination order: every lane its own 32-byte sector, ~2 GB of L2 -> SM traffic for 123 MB of
  // weights, 0.43-0.50 ms per step.
  __shared__ __nv_bfloat16 tile[kPackTCO * 292];
  const long long tile_id = blockIdx.x;
  if (tile_id >= total_tiles) return;
  int a = 0, b = n - 1;
  while (a < b) {                       // last entry with start <= tile_id (block-uniform)
    const int mid = (a + b + 1) >> 1;
    if (tab[mid].start <= tile_id) a = mid; else b = mid - 1;
  }
  const m3t_pack_entry& e = tab[a];
  const unsigned local = (unsigned)(tile_id - e.start);
  const unsigned Cout = (unsigned)e.Cout, Cin = (unsigned)e.Cin, taps = (unsigned)e.taps;
  const float* __restrict__ src = reinterpret_cast<const float*>(e.src);
  const int hp = e.has_parity;
  if (hp < 0) {                         // plain fp32 copy (stacked GRU bias vectors)
    const size_t count = (size_t)Cout * Cin * taps;
    float* dst = reinterpret_cast<float*>(e.wf);
    for (unsigned k = threadIdx.x; k < (unsigned)kPackCopy; k += blockDim.x) {
      const size_t i = (size_t)local * kPackCopy + k;
      if (i < count) dst[i] = __ldg(src + i);
    }
    return;
  }
  const unsigned TCI = (unsigned)pack_tci((int)taps);
  const unsigned ntci = (Cin + TCI - 1) / TCI;
  const unsigned co0 = (local / ntci) * kPackTCO, ci0 = (local % ntci) * TCI;
  const unsigned nco = min((unsigned)kPackTCO, Cout - co0), nci = min(TCI, Cin - ci0);
  const unsigned RP = nci * taps;                   // floats per source row of the tile
  unsigned pitch = (RP + 1) & ~1u;                  // bf16 row pitch with an odd number of 4-byte words:
  if ((pitch & 3u) == 0) pitch += 2;                // the column reads of the dgrad pass are conflict-free
  for (unsigned idx = threadIdx.x; idx < nco * RP; idx += blockDim.x) {
    const unsigned co = idx / RP, rem = idx - co * RP;
    tile[co * pitch + rem] = __float2bfloat16(__ldg(src + ((size_t)(co0 + co) * Cin + ci0) * taps + rem));
  }
  __syncthreads();
  if (e.wf) {
    __nv_bfloat16* __restrict__ wf = reinterpret_cast<__nv_bfloat16*>(e.wf);
    for (unsigned idx = threadIdx.x; idx < nco * RP; idx += blockDim.x) {
      const unsigned ci = idx % nci, r = idx / nci;
      const unsigned t = r % taps, co = r / taps;
      wf[((size_t)(co0 + co) * taps + t) * Cin + ci0 + ci] = tile[co * pitch + ci * taps + t];
    }
  }
  if (e.wd || hp > 0) {
    __nv_bfloat16* __restrict__ wd = reinterpret_cast<__nv_bfloat16*>(e.wd);
    for (unsigned idx = threadIdx.x; idx < nco * RP; idx += blockDim.x) {
      const unsigned co = idx % nco, r = idx / nco;
      const unsigned tf = r % taps, ci = r / taps;      // tf: flipped tap index of the dgrad pack
      const unsigned t = taps - 1 - tf;
      const __nv_bfloat16 v = tile[co * pitch + ci * taps + t];
      if (wd) wd[((size_t)(ci0 + ci) * taps + tf) * Cout + co0 + co] = v;
      if (hp > 0) {
        const int pp = e.par_of_tap[t];
        if (pp >= 0)
          reinterpret_cast<__nv_bfloat16*>(e.par[pp])[((size_t)(ci0 + ci) * e.ntaps_par[pp] + e.pos_of_tap[t]) * Cout +
                                                      co0 + co] = v;
      }
    }
  }
}

__global__ void unpack_filter_grad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Cout, int Cin,
                                          int taps) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const long long total = (long long)Cout * Cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    long long r = i / taps;
    const int ci = (int)(r % Cin);
    const int co = (int)(r / Cin);
    dw[i] = dwp[((long long)co * taps + t) * Cin + ci];
  }
}

// Stride-2 dgrad helper: dy_up[n][2p][2q][c] = dy[n][p][q][c], zero elsewhere (Hup x Wup spatial extent).
__global__ void zero_insert2_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ up, int N, int P,
                                    int Q, int Hup, int Wup, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int cgs = C / 8;
  const long long total = (long long)N * Hup * Wup * cgs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgs);
    long long r = i / cgs;
    const int w = (int)(r % Wup);
    r /= Wup;
    const int h = (int)(r % Hup);
    const int n = (int)(r / Hup);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (!(h & 1) && !(w & 1) && (h >> 1) < P && (w >> 1) < Q)
      v = __ldg(reinterpret_cast<const uint4*>(dy + (((long long)n * P + (h >> 1)) * Q + (w >> 1)) * C) + cg);
    reinterpret_cast<uint4*>(up)[i] = v;
  }
}

// out = a + b (bf16, vectors of 8) — gradient fan-in at residual forks
__global__ void add_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                __nv_bfloat16* __restrict__ out, long long nvec) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += 4 * stride) {
    uint4 va[4], vb[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {     // four vectors in flight
      const long long ih = i + h * stride;
      if (ih < nvec) {
        va[h] = __ldg(reinterpret_cast<const uint4*>(a) + ih);
        vb[h] = __ldg(reinterpret_cast<const uint4*>(b) + ih);
      }
    }
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const long long ih = i + h * stride;
      if (ih < nvec) {
        float x[8], y[8];
        unpack8(va[h], x);
        unpack8(vb[h], y);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] += y[j];
        reinterpret_cast<uint4*>(out)[ih] = pack8(x);
      }
    }
  }
}

// Inverted dropout with the counter-based generator of common.cuh (dropout_u32): the backward pass re-derives the
// mask from (seed, i) instead of reading one.
__global__ void dropout_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long nvec,
                                    unsigned thresh, float scale, unsigned long long seed) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = dropout_u32(seed, i * 8 + j) >= thresh ? v[j] * scale : 0.f;
    reinterpret_cast<uint4*>(y)[i] = pack8(v);
  }
}

// out_bf16[r][k] = idx[k] >= 0 ? w[r*row_stride + idx[k]] : 0      (filter re-layout, e.g. the space-to-depth stem)
__global__ void gather_pack_kernel(const float* __restrict__ w, const int* __restrict__ idx,
                                   __nv_bfloat16* __restrict__ out, int rows, long long row_stride, int K) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const long long total = (long long)rows * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const long long r = i / K;
    const int s = idx[k];
    out[i] = __float2bfloat16(s >= 0 ? w[r * row_stride + s] : 0.f);
  }
}
// dw[r*row_stride + idx[k]] = dwp[r][k] for idx[k] >= 0 (each target written once; the caller zero-fills dw)
__global__ void scatter_unpack_kernel(const float* __restrict__ dwp, const int* __restrict__ idx,
                                      float* __restrict__ dw, int rows, long long row_stride, int K) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const long long total = (long long)rows * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const long long r = i / K;
    const int s = idx[k];
    if (s >= 0) dw[r * row_stride + s] = dwp[i];
  }
}

// column sums of a [rows][ld] bf16 matrix (bias gradients): out[c] = sum_r x[r][c], c < cols
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long ld, long long rows, int cols,
                                   float* __restrict__ out) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  // block: 32 columns x 8 row lanes
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < cols)
    for (long long r = blockIdx.y * 8 + threadIdx.y; r < rows; r += (long long)gridDim.y * 8)
      acc += __bfloat162float(x[r * ld + c]);
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += part[j][threadIdx.x];
    atomicAdd(out + c, s);
  }
}

// dz = dy * (out > 0)   (ReLU backward on bf16 vectors of 8)
__global__ void relu_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ out,
                                __nv_bfloat16* __restrict__ dz, long long nvec) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float d[8], o[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy) + i), d);
    unpack8(__ldg(reinterpret_cast<const uint4*>(out) + i), o);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = o[j] > 0.f ? d[j] : 0.f;
    reinterpret_cast<uint4*>(dz)[i] = pack8(d);
  }
}

// Backward of the fused TemporalBlock conv epilogue  y = [relu](drop(relu(a)) + res):
//   dsum = dy * [y > 0]            (only when a residual joined; also the residual's gradient)
//   da   = dsum * scale * [t > 0]  (t = drop(relu(a)): positive exactly where the ReLU passed AND dropout kept it)
__global__ void tcn_epi_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                   const __nv_bfloat16* __restrict__ t, __nv_bfloat16* __restrict__ dsum,
                                   __nv_bfloat16* __restrict__ da, float scale, long long nvec) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float d[8], tv[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy) + i), d);
    unpack8(__ldg(reinterpret_cast<const uint4*>(t) + i), tv);
    if (y) {
      float yv[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(y) + i), yv);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = yv[j] > 0.f ? d[j] : 0.f;
      reinterpret_cast<uint4*>(dsum)[i] = pack8(d);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = tv[j] > 0.f ? d[j] * scale : 0.f;
    reinterpret_cast<uint4*>(da)[i] = pack8(d);
  }
}

// 3x3 / pad 1 patches of a single-channel image as GEMM rows: out[(n*H + h)*W + w][kh*3 + kw] (9 taps, zero padded
// to 16 columns) in bf16.  Feeds the 1-channel stem of the audio ResNet composition (BASELINE config 2).
__global__ void patch3x3_c1_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int H, int W) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const long long total = (long long)N * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const long long r = i / W;
    const int h = (int)(r % H);
    const long long n = r / H;
    float f[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int hh = h + kh - 1, ww = w + kw - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) f[kh * 3 + kw] = __ldg(x + (n * H + hh) * W + ww);
      }
    float lo[8], hi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { lo[j] = f[j]; hi[j] = f[8 + j]; }
    uint4* o = reinterpret_cast<uint4*>(out + i * 16);
    o[0] = pack8(lo);
    o[1] = pack8(hi);
  }
}

// All bf16 copies one bidirectional GRU layer needs, in ONE launch (they are rebuilt after every optimizer step):
//   wih  [6H][Ipad]   = [W_ih ; W_ih_reverse] (columns >= I zero)      input projection, both directions
//   whh  [2][3H][H]   = W_hh per direction                               forward recurrence
//   whht [2][H][3H]   = W_hh^T per direction                             BPTT recurrence
//   bias [2][6H] fp32 = [b_ih ; b_ih_reverse], [b_hh ; b_hh_reverse] (optional)
__global__ void gru_pack_weights_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_ih_r,
                                        const float* __restrict__ w_hh, const float* __restrict__ w_hh_r,
                                        __nv_bfloat16* __restrict__ wih, __nv_bfloat16* __restrict__ whh,
                                        __nv_bfloat16* __restrict__ whht, int I, int Ipad, int H,
                                        const float* __restrict__ b_ih, const float* __restrict__ b_ih_r,
                                        const float* __restrict__ b_hh, const float* __restrict__ b_hh_r,
                                        float* __restrict__ bias) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const long long n_ih = 6LL * H * Ipad, n_hh = 6LL * H * H;
  const long long total = n_ih + n_hh;
  if (bias) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 12 * H; i += gridDim.x * blockDim.x) {
      const int which = i / (6 * H), r = i - which * 6 * H;
      const float* src = which == 0 ? (r < 3 * H ? b_ih : b_ih_r) : (r < 3 * H ? b_hh : b_hh_r);
      bias[i] = src[r < 3 * H ? r : r - 3 * H];
    }
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    if (i < n_ih) {
      const int c = (int)(i % Ipad);
      const long long r = i / Ipad;
      float v = 0.f;
      if (c < I) v = r < 3 * H ? w_ih[r * I + c] : w_ih_r[(r - 3 * H) * I + c];
      wih[i] = __float2bfloat16(v);
    } else {
      const long long j = i - n_ih;                 // index into [2][3H][H]
      const int k = (int)(j % H);
      const long long r = j / H;
      const int d = (int)(r / (3 * H));
      const int g = (int)(r - (long long)d * 3 * H);
      const __nv_bfloat16 v = __float2bfloat16(d == 0 ? w_hh[(long long)g * H + k] : w_hh_r[(long long)g * H + k]);
      whh[j] = v;
      if (whht) whht[((long long)d * H + k) * 3 * H + g] = v;
    }
  }
}

}  // namespace m3t

using namespace m3t;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" int m3t_video_prep_s2d(const void* video, int is_u8, void* out, int B, int T, int H, int W, float mul,
                                  float add, void* stream) {
  if ((H | W) & 1) return -1;
  const long long items = (long long)B * T * (H / 2) * (W / 2);
  if (is_u8)
    m3t::launch_k(video_prep_s2d_kernel<uint8_t>, dim3(ew_blocks(items)), dim3(kEwThreads), 0, ST(stream), 
        reinterpret_cast<const uint8_t*>(video), BF(out), B, T, H, W, mul, add);
  else
    m3t::launch_k(video_prep_s2d_kernel<float>, dim3(ew_blocks(items)), dim3(kEwThreads), 0, ST(stream), 
        reinterpret_cast<const float*>(video), BF(out), B, T, H, W, mul, add);
  count_launch();
  return launch_status();
}

extern "C" int m3t_video_prep_s2d_w4(const void* video, int is_u8, void* out, int B, int T, int H, int W, float mul,
                                     float add, void* stream) {
  if ((H | W) & 1) return -1;
  const long long items = (long long)B * T * (H / 2) * (W / 2);
  if (is_u8)
    m3t::launch_k(video_prep_s2d_w4_kernel<uint8_t>, dim3(ew_blocks(items)), dim3(kEwThreads), 0, ST(stream), 
        reinterpret_cast<const uint8_t*>(video), BF(out), B, T, H, W, mul, add);
  else
    m3t::launch_k(video_prep_s2d_w4_kernel<float>, dim3(ew_blocks(items)), dim3(kEwThreads), 0, ST(stream), 
        reinterpret_cast<const float*>(video), BF(out), B, T, H, W, mul, add);
  count_launch();
  return launch_status();
}

extern "C" int m3t_video_augment_prep_s2d_w4(const void* frames_u8, const int* params, void* out, int B, int T, int Hs,
                                             int Ws, int H, int W, float mul, float add, void* stream) {
  if ((H | W) & 1 || H > Hs || W > Ws || B <= 0 || T <= 0) return -1;
  const long long items = (long long)B * T * (H / 2) * (W / 2);
  m3t::launch_k(video_augment_prep_s2d_w4_kernel, dim3(ew_blocks(items)), dim3(kEwThreads), 0, ST(stream), 
      reinterpret_cast<const uint8_t*>(frames_u8), params, BF(out), B, T, Hs, Ws, H, W, mul, add);
  count_launch();
  return launch_status();
}

extern "C" int m3t_bn_finalize(const float* stats, int C, double count, const float* gamma, const float* beta,
                               float eps, float momentum, float* running_mean, float* running_var, float* mean,
                               float* invstd, float* scale, float* shift, void* stream) {
  m3t::launch_k(bn_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, ST(stream), stats, C, (float)count, gamma, beta, eps, momentum,
                                                              running_mean, running_var, mean, invstd, scale, shift);
  count_launch();
  return launch_status();
}

extern "C" int m3t_bn_fold(int C, const float* gamma, const float* beta, const float* running_mean,
                           const float* running_var, const float* conv_bias, float eps, float* scale, float* shift,
                           void* stream) {
  m3t::launch_k(bn_fold_kernel, dim3((C + 127) / 128), dim3(128), 0, ST(stream), C, gamma, beta, running_mean, running_var, conv_bias, eps,
                                                          scale, shift);
  count_launch();
  return launch_status();
}

extern "C" int m3t_bn_act(const void* y, const float* scale, const float* shift, const void* res,
                          const float* res_scale, const float* res_shift, int relu, void* out, long long rows, int C,
                          void* stream) {
  if (C % 8) return -1;
  const long long nvec = rows * C / 8;
  if (kEwThreads % (C / 8) == 0 && !res_scale) {
    const int blocks = ew_blocks((nvec + 1) / 2);
#define ACT_FIXED(HR, RL)                                                                                      \
  m3t::launch_k(bn_act_fixed_kernel<HR, RL>, dim3(blocks), dim3(kEwThreads), 0, ST(stream), CBF(y), scale, shift, CBF(res), BF(out), nvec, C)
    if (res && relu) ACT_FIXED(true, true);
    else if (res) ACT_FIXED(true, false);
    else if (relu) ACT_FIXED(false, true);
    else ACT_FIXED(false, false);
#undef ACT_FIXED
    count_launch();
    return launch_status();
  }
  m3t::launch_k(bn_act_kernel, dim3(ew_blocks(nvec)), dim3(kEwThreads), 0, ST(stream), CBF(y), scale, shift, CBF(res), res_scale, res_shift,
                                                                relu, BF(out), nvec, C);
  count_launch();
  return launch_status();
}

extern "C" int m3t_bn_bwd_reduce(const void* dout, const void* out, const void* y, const float* mean,
                                 const float* invstd, const float* scale, const float* shift, int relu, void* dz_out,
                                 float* sums, long long rows, int C, void* stream) {
  // any C % 8 == 0 up to 8 * kEwThreads: when C/8 does not divide the block, the last kEwThreads % (C/8) threads idle
  // (DenseNet's 96 / 160 / 200 ... channel tensors); the trunk's power-of-two widths use every thread
  if (C % 8 || C / 8 > kEwThreads) return -1;
  const bool det = (relu & 256) != 0;
  relu &= 255;
  const int cgs = C / 8;
  const int rows_per_iter = kEwThreads / cgs;
  long long blocks = (rows + rows_per_iter - 1) / rows_per_iter;
  // exactly one resident wave (launch bounds: 3 CTAs per SM, 2 for the mask-from-y variant): every CTA gets the same
  // share of the grid-stride loop, no partial last wave
  const long long cap = 148LL * (relu == 2 ? 2 : 3);
  if (blocks > cap) blocks = cap;
  const size_t sm = 2 * C * sizeof(float);
  if (det) {       // bit 8 of `relu`: `sums` is the first of 1 + m3t_det_stats_slots() zero-filled copies
    const size_t smd = (size_t)rows_per_iter * 2 * C * sizeof(float);
#define M3T_RED_DET(R)                                                                                              \
  m3t::launch_k(bn_bwd_reduce_kernel<R, true>, dim3((int)blocks), dim3(kEwThreads), smd, ST(stream), CBF(dout), CBF(out), CBF(y), mean,    \
                                                                             invstd, scale, shift, BF(dz_out), sums, \
                                                                             rows, C)
    if (relu == 0) M3T_RED_DET(0);
    else if (relu == 1) M3T_RED_DET(1);
    else M3T_RED_DET(2);
#undef M3T_RED_DET
    count_launch();
    return launch_status();
  }
  if (relu == 0)
    m3t::launch_k(bn_bwd_reduce_kernel<0>, dim3((int)blocks), dim3(kEwThreads), sm, ST(stream), CBF(dout), CBF(out), CBF(y), mean, invstd,
                                                                        scale, shift, BF(dz_out), sums, rows, C);
  else if (relu == 1)
    m3t::launch_k(bn_bwd_reduce_kernel<1>, dim3((int)blocks), dim3(kEwThreads), sm, ST(stream), CBF(dout), CBF(out), CBF(y), mean, invstd,
                                                                        scale, shift, BF(dz_out), sums, rows, C);
  else
    m3t::launch_k(bn_bwd_reduce_kernel<2>, dim3((int)blocks), dim3(kEwThreads), sm, ST(stream), CBF(dout), CBF(out), CBF(y), mean, invstd,
                                                                        scale, shift, BF(dz_out), sums, rows, C);
  count_launch();
  return launch_status();
}

extern "C" int m3t_bn_bwd_apply(const void* dout, const void* out, const void* y, const float* mean,
                                const float* invstd, const float* scale, const float* shift, const float* sums,
                                double count, int relu, void* dy, long long rows, int C, void* stream) {
  if (C % 8) return -1;
  const long long nvec = rows * C / 8;
  if (kEwThreads % (C / 8) == 0) {
    const int blocks = ew_blocks((nvec + 1) / 2);
    const float ic = (float)(1.0 / count);
#define APPLY_FIXED(R)                                                                                              \
  m3t::launch_k(bn_bwd_apply_fixed_kernel<R>, dim3(blocks), dim3(kEwThreads), 0, ST(stream), CBF(dout), CBF(out), CBF(y), mean, invstd,    \
                                                                     scale, shift, sums, ic, BF(dy), nvec, C)
    if (relu == 0) APPLY_FIXED(0);
    else if (relu == 1) APPLY_FIXED(1);
    else APPLY_FIXED(2);
#undef APPLY_FIXED
    count_launch();
    return launch_status();
  }
  m3t::launch_k(bn_bwd_apply_kernel, dim3(ew_blocks(nvec)), dim3(kEwThreads), 0, ST(stream), CBF(dout), CBF(out), CBF(y), mean, invstd, scale,
                                                                      shift, sums, (float)(1.0 / count), relu, BF(dy),
                                                                      nvec, C);
  count_launch();
  return launch_status();
}

static int ilog2_exact(int v) {
  int s = 0;
  while ((1 << s) < v) ++s;
  return (1 << s) == v ? s : -1;
}

static int bn_relu_maxpool_impl(const void* y, const float* scale, const float* shift, void* out, void* idx,
                                void* ymax, int F, int H, int W, int C, int K, int S, int PAD, void* stream);

extern "C" int m3t_bn_relu_maxpool(const void* y, const float* scale, const float* shift, void* out, void* idx, int F,
                                   int H, int W, int C, int K, int S, int PAD, void* stream) {
  return bn_relu_maxpool_impl(y, scale, shift, out, idx, nullptr, F, H, W, C, K, S, PAD, stream);
}

extern "C" int m3t_bn_relu_maxpool_ymax(const void* y, const float* scale, const float* shift, void* out, void* idx,
                                        void* ymax, int F, int H, int W, int C, int K, int S, int PAD, void* stream) {
  return bn_relu_maxpool_impl(y, scale, shift, out, idx, ymax, F, H, W, C, K, S, PAD, stream);
}

static int bn_relu_maxpool_impl(const void* y, const float* scale, const float* shift, void* out, void* idx,
                                void* ymax, int F, int H, int W, int C, int K, int S, int PAD, void* stream) {
  const int sh = C % 8 ? -1 : ilog2_exact(C / 8);
  if (sh < 0) return -1;
  const int P = (H + 2 * PAD - K) / S + 1, Q = (W + 2 * PAD - K) / S + 1;
  const long long items = (long long)F * P * Q * (C / 8);
  if ((long long)F * H * W * (C / 8) >= (1LL << 31)) return -6;
  const int blocks = ew_blocks(items);
  if (K == 3 && S == 2 && PAD == 1 && H % 2 == 0 && W % 2 == 0 && kEwThreads % (C / 8) == 0)
    m3t::launch_k(bn_relu_maxpool3s2_kernel, dim3(blocks), dim3(kEwThreads), 0, ST(stream), CBF(y), scale, shift, BF(out),
                                                                   reinterpret_cast<uint8_t*>(idx), BF(ymax), F, H, W, C,
                                                                   sh);
  else if (K == 3 && S == 2 && PAD == 1)
    m3t::launch_k(bn_relu_maxpool_kernel<3, 2, 1>, dim3(blocks), dim3(kEwThreads), 0, ST(stream), 
        CBF(y), scale, shift, BF(out), reinterpret_cast<uint8_t*>(idx), BF(ymax), F, H, W, C, sh);
  else if (K == 2 && S == 2 && PAD == 0)
    m3t::launch_k(bn_relu_maxpool_kernel<2, 2, 0>, dim3(blocks), dim3(kEwThreads), 0, ST(stream), 
        CBF(y), scale, shift, BF(out), reinterpret_cast<uint8_t*>(idx), BF(ymax), F, H, W, C, sh);
  else
    return -1;
  count_launch();
  return launch_status();
}

template <int K, int S, int PAD>
static void launch_pool_bwd(int mode, const void* dout, const void* idx, const void* y, const float* mean,
                            const float* invstd, const float* scale, const float* shift, float* sums, double count,
                            void* dy, int F, int H, int W, int C, int sh, int blocks, cudaStream_t st) {
  if (mode == 0)
    m3t::launch_k(maxpool_bn_bwd_kernel<0, K, S, PAD>, dim3(blocks), dim3(kEwThreads), 2 * C * sizeof(float), st, 
        CBF(dout), reinterpret_cast<const uint8_t*>(idx), CBF(y), mean, invstd, scale, shift, sums, 0.f, nullptr, F, H,
        W, C, sh);
  else
    m3t::launch_k(maxpool_bn_bwd_kernel<1, K, S, PAD>, dim3(blocks), dim3(kEwThreads), 0, st, 
        CBF(dout), reinterpret_cast<const uint8_t*>(idx), CBF(y), mean, invstd, scale, shift, sums,
        (float)(1.0 / count), BF(dy), F, H, W, C, sh);
}

extern "C" int m3t_maxpool_bn_bwd(int mode, const void* dout, const void* idx, const void* y, const float* mean,
                                  const float* invstd, const float* scale, const float* shift, float* sums,
                                  double count, void* dy, int F, int H, int W, int C, int K, int S, int PAD,
                                  void* stream) {
  const int sh = C % 8 ? -1 : ilog2_exact(C / 8);
  if (sh < 0 || kEwThreads % (C / 8)) return -1;
  const long long items = (long long)F * H * W * (C / 8);
  if (items >= (1LL << 31)) return -6;
  const int blocks = ew_blocks(items);
  if (K == 3 && S == 2 && PAD == 1 && H % 2 == 0 && W % 2 == 0) {
    const int blocks4 = ew_blocks(items / 4);
    if (mode == 0)
      m3t::launch_k(maxpool3s2_bn_bwd_kernel<0>, dim3(blocks4), dim3(kEwThreads), 2 * C * sizeof(float), ST(stream), 
          CBF(dout), reinterpret_cast<const uint8_t*>(idx), CBF(y), mean, invstd, scale, shift, sums, 0.f, nullptr, F,
          H, W, C, sh);
    else
      m3t::launch_k(maxpool3s2_bn_bwd_kernel<1>, dim3(blocks4), dim3(kEwThreads), 0, ST(stream), 
          CBF(dout), reinterpret_cast<const uint8_t*>(idx), CBF(y), mean, invstd, scale, shift, sums,
          (float)(1.0 / count), BF(dy), F, H, W, C, sh);
  } else if (K == 3 && S == 2 && PAD == 1)
    launch_pool_bwd<3, 2, 1>(mode, dout, idx, y, mean, invstd, scale, shift, sums, count, dy, F, H, W, C, sh, blocks,
                             ST(stream));
  else if (K == 2 && S == 2 && PAD == 0)
    launch_pool_bwd<2, 2, 0>(mode, dout, idx, y, mean, invstd, scale, shift, sums, count, dy, F, H, W, C, sh, blocks,
                             ST(stream));
  else
    return -1;
  count_launch();
  return launch_status();
}

extern "C" int m3t_avgpool(const void* x, void* out_bf16, float* out_f32, int F, int HW, int C, void* stream) {
  if (C % 8) return -1;
  m3t::launch_k(avgpool_kernel, dim3(ew_blocks((long long)F * C / 8)), dim3(kEwThreads), 0, ST(stream), CBF(x), BF(out_bf16), out_f32, F, HW,
                                                                                 C);
  count_launch();
  return launch_status();
}

extern "C" int m3t_avgpool_bwd(const void* dout, int dout_f32, void* dx, int F, int HW, int C, void* stream) {
  if (C % 8) return -1;
  const long long items = (long long)F * HW * C / 8;
  if (dout_f32)
    m3t::launch_k(avgpool_bwd_kernel<float>, dim3(ew_blocks(items)), dim3(kEwThreads), 0, ST(stream), reinterpret_cast<const float*>(dout),
                                                                               BF(dx), F, HW, C);
  else
    m3t::launch_k(avgpool_bwd_kernel<__nv_bfloat16>, dim3(ew_blocks(items)), dim3(kEwThreads), 0, ST(stream), CBF(dout), BF(dx), F, HW, C);
  count_launch();
  return launch_status();
}

extern "C" int m3t_ncs_f32_to_nsc_bf16(const float* in, void* out, int N, int C, int S, int Cpad, void* stream) {
  dim3 grid((S + 31) / 32, (Cpad + 31) / 32, N), block(32, 8);
  m3t::launch_k(ncs_f32_to_nsc_bf16_kernel, dim3(grid), dim3(block), 0, ST(stream), in, BF(out), C, S, Cpad);
  count_launch();
  return launch_status();
}

extern "C" int m3t_nsc_to_ncs_f32(const void* in, int in_f32, float* out, int N, int C, int S, int Cpad,
                                  void* stream) {
  dim3 grid((S + 31) / 32, (C + 31) / 32, N), block(32, 8);
  if (in_f32)
    m3t::launch_k(nsc_to_ncs_f32_kernel<float>, dim3(grid), dim3(block), 0, ST(stream), reinterpret_cast<const float*>(in), out, C, S, Cpad);
  else
    m3t::launch_k(nsc_to_ncs_f32_kernel<__nv_bfloat16>, dim3(grid), dim3(block), 0, ST(stream), CBF(in), out, C, S, Cpad);
  count_launch();
  return launch_status();
}

extern "C" int m3t_cast_f32_bf16(const float* in, long long ld_in, void* out, long long ld_out, long long rows,
                                 int cols, void* stream) {
  m3t::launch_k(cast_f32_bf16_kernel, dim3(ew_blocks(rows * ld_out)), dim3(kEwThreads), 0, ST(stream), in, ld_in, BF(out), ld_out, rows,
                                                                                cols);
  count_launch();
  return launch_status();
}

extern "C" int m3t_cast_bf16_f32(const void* in, long long ld_in, float* out, long long ld_out, long long rows,
                                 int cols, void* stream) {
  m3t::launch_k(cast_bf16_f32_kernel, dim3(ew_blocks(rows * cols)), dim3(kEwThreads), 0, ST(stream), CBF(in), ld_in, out, ld_out, rows, cols);
  count_launch();
  return launch_status();
}

extern "C" int m3t_gru_pack_weights(const float* w_ih, const float* w_ih_r, const float* w_hh, const float* w_hh_r,
                                    void* wih, void* whh, void* whht, int I, int Ipad, int H, const float* b_ih,
                                    const float* b_ih_r, const float* b_hh, const float* b_hh_r, float* bias,
                                    void* stream) {
  if (I <= 0 || H <= 0 || Ipad < I) return -1;
  if (bias && !(b_ih && b_ih_r && b_hh && b_hh_r)) return -1;
  const long long total = 6LL * H * Ipad + 6LL * H * H;
  m3t::launch_k(gru_pack_weights_kernel, dim3(ew_blocks(total)), dim3(kEwThreads), 0, ST(stream), w_ih, w_ih_r, w_hh, w_hh_r, BF(wih), BF(whh),
                                                                          BF(whht), I, Ipad, H, b_ih, b_ih_r, b_hh,
                                                                          b_hh_r, bias);
  count_launch();
  return launch_status();
}

extern "C" int m3t_pack_filter(const float* w, void* w_fprop, void* w_dgrad, int Cout, int Cin, int taps,
                               void* stream) {
  m3t::launch_k(pack_filter_kernel, dim3(ew_blocks((long long)Cout * Cin * taps)), dim3(kEwThreads), 0, ST(stream), w, BF(w_fprop),
                                                                                           BF(w_dgrad), Cout, Cin,
                                                                                           taps);
  count_launch();
  return launch_status();
}

extern "C" int m3t_pack_filters_batched(const m3t_pack_entry* table_dev, int n, long long total_tiles, void* stream) {
  if (n <= 0 || n > 128 || total_tiles <= 0 || total_tiles > 0x7fffffffLL) return -1;
  m3t::launch_k(pack_filters_batched_kernel, dim3((unsigned)total_tiles), dim3(kEwThreads), 0, ST(stream), table_dev, n,
                total_tiles);
  count_launch();
  return launch_status();
}

// Number of tiles (= blocks) an entry of m3t_pack_filters_batched occupies; callers lay out entry.start with it.
extern "C" long long m3t_pack_entry_tiles(int Cout, int Cin, int taps, int has_parity) {
  if (Cout <= 0 || Cin <= 0 || taps <= 0) return -1;
  if (has_parity < 0) return ((long long)Cout * Cin * taps + kPackCopy - 1) / kPackCopy;
  const int tci = pack_tci(taps);
  return (long long)((Cout + kPackTCO - 1) / kPackTCO) * ((Cin + tci - 1) / tci);
}

extern "C" int m3t_unpack_filter_grad(const float* dw_packed, float* dw, int Cout, int Cin, int taps, void* stream) {
  m3t::launch_k(unpack_filter_grad_kernel, dim3(ew_blocks((long long)Cout * Cin * taps)), dim3(kEwThreads), 0, ST(stream), dw_packed, dw,
                                                                                                  Cout, Cin, taps);
  count_launch();
  return launch_status();
}

extern "C" int m3t_zero_insert2(const void* dy, void* up, int N, int P, int Q, int Hup, int Wup, int C,
                                void* stream) {
  if (C % 8) return -1;
  m3t::launch_k(zero_insert2_kernel, dim3(ew_blocks((long long)N * Hup * Wup * C / 8)), dim3(kEwThreads), 0, ST(stream), CBF(dy), BF(up), N,
                                                                                                 P, Q, Hup, Wup, C);
  count_launch();
  return launch_status();
}

extern "C" int m3t_add_bf16(const void* a, const void* b, void* out, long long n, void* stream) {
  if (n % 8) return -1;
  m3t::launch_k(add_bf16_kernel, dim3(ew_blocks((n / 8 + 3) / 4)), dim3(kEwThreads), 0, ST(stream), CBF(a), CBF(b), BF(out), n / 8);
  count_launch();
  return launch_status();
}

extern "C" int m3t_gather_pack_bf16(const float* w, const int* idx, void* out, int rows, long long row_stride, int K,
                                    void* stream) {
  m3t::launch_k(gather_pack_kernel, dim3(ew_blocks((long long)rows * K)), dim3(kEwThreads), 0, ST(stream), w, idx, BF(out), rows, row_stride,
                                                                                  K);
  count_launch();
  return launch_status();
}

extern "C" int m3t_scatter_unpack_f32(const float* dwp, const int* idx, float* dw, int rows, long long row_stride,
                                      int K, void* stream) {
  m3t::launch_k(scatter_unpack_kernel, dim3(ew_blocks((long long)rows * K)), dim3(kEwThreads), 0, ST(stream), dwp, idx, dw, rows, row_stride,
                                                                                     K);
  count_launch();
  return launch_status();
}

// buf = (1 + nslots) consecutive copies of a `len`-float vector: copy 0 += copy 1 + copy 2 + ... in index order.
__global__ void det_reduce_kernel(float* __restrict__ buf, long long len, int nslots) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < len;
       i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 1; k <= nslots; ++k) s += buf[(long long)k * len + i];
    buf[i] += s;
  }
}

extern "C" int m3t_det_stats_slots(void) { return 148 * 4; }
extern "C" int m3t_det_cta_slots(void) { return 148; }

extern "C" int m3t_det_reduce(float* buf, long long len, int nslots, void* stream) {
  if (!buf || len <= 0 || nslots <= 0) return -1;
  m3t::launch_k(det_reduce_kernel, dim3(ew_blocks(len)), dim3(kEwThreads), 0, ST(stream), buf, len, nslots);
  count_launch();
  return launch_status();
}

extern "C" int m3t_colsum_bf16(const void* x, long long ld, long long rows, int cols, float* out, void* stream) {
  const bool det = (cols & (1 << 30)) != 0;      // bit 30: deterministic (one row-block per column strip, no atomics race)
  cols &= ~(1 << 30);
  long long gy = (rows + 8 * 64 - 1) / (8 * 64);
  if (gy > 148) gy = 148;
  if (gy < 1 || det) gy = 1;
  dim3 grid((cols + 31) / 32, (unsigned)gy), block(32, 8);
  m3t::launch_k(colsum_bf16_kernel, dim3(grid), dim3(block), 0, ST(stream), CBF(x), ld, rows, cols, out);
  count_launch();
  return launch_status();
}

extern "C" int m3t_relu_bwd_bf16(const void* dy, const void* out, void* dz, long long n, void* stream) {
  if (n % 8) return -1;
  m3t::launch_k(relu_bwd_kernel, dim3(ew_blocks(n / 8)), dim3(kEwThreads), 0, ST(stream), CBF(dy), CBF(out), BF(dz), n / 8);
  count_launch();
  return launch_status();
}

extern "C" int m3t_tcn_epilogue_bwd_bf16(const void* dy, const void* y, const void* t, void* dsum, void* da,
                                         float scale, long long n, void* stream) {
  if (n % 8 || !dy || !t || !da || ((y != nullptr) != (dsum != nullptr))) return -1;
  m3t::launch_k(tcn_epi_bwd_kernel, dim3(ew_blocks(n / 8)), dim3(kEwThreads), 0, ST(stream), CBF(dy), CBF(y), CBF(t), BF(dsum), BF(da), scale,
                                                                      n / 8);
  count_launch();
  return launch_status();
}

extern "C" int m3t_dropout_bf16(const void* x, void* y, long long n, float p, unsigned long long seed, void* stream) {
  if (n % 8 || !(p >= 0.f) || !(p < 1.f)) return -1;
  const unsigned thresh = dropout_threshold(p);
  m3t::launch_k(dropout_bf16_kernel, dim3(ew_blocks(n / 8)), dim3(kEwThreads), 0, ST(stream), CBF(x), BF(y), n / 8, thresh, 1.f / (1.f - p),
                                                                       seed);
  count_launch();
  return launch_status();
}

extern "C" int m3t_patch3x3_c1(const float* x, void* out, int N, int H, int W, void* stream) {
  m3t::launch_k(patch3x3_c1_kernel, dim3(ew_blocks((long long)N * H * W)), dim3(kEwThreads), 0, ST(stream), x, BF(out), N, H, W);
  count_launch();
  return launch_status();
}
