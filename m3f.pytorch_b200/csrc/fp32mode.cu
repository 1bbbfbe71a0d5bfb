// fp32-parity inference mode (north-star tolerance: max error <= 1e-4 on the per-frame V/A predictions).
//
// Activations stay float32 in HBM; the tensor-core kernels are the SAME bf16 tcgen05 kernels, fed with split operands:
// every float32 value x is written as hi = bf16(x) and lo = bf16(x - hi) (x = hi + lo up to 2^-17 |x|), and a product
// sum_k a_k b_k is evaluated as  sum a_hi b_hi + a_hi b_lo + a_lo b_hi  (the dropped lo*lo term is 2^-16 relative) by
// CONCATENATING the three terms along the contraction:  A' = [a_hi | a_hi | a_lo],  B' = [b_hi | b_lo | b_hi]  per
// 64-wide (or whole-row) channel block.  bf16 x bf16 products are exact in fp32 and accumulate in the fp32 TMEM
// accumulator, so one ordinary bf16 GEMM / implicit-GEMM launch over 3x the channels yields the fp32-grade result -
// no new MMA kernel.  This file holds the passes around it, all float32:
//   split3 (+ residual + ReLU)      activation  f32 [rows][C]        -> bf16 [rows][3C] (+ the f32 activation)
//   pack_split3                     weights     f32 [N][G][C]        -> bf16 [N][G][3C]
//   video_prep_s2d_w4_split3        stem input  (B,3,T,H,W) f32/u8   -> bf16 (B,T,H/2,W/2,192)
//   maxpool3s2_f32, avgpool_f32, att_mix_f32, gru_fwd_f32 (FFMA recurrence, one launch per time step)
// Forward / eval only; the training path is the bf16 one.
#include "../../include/m3t_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace m3t {

constexpr int kFpThreads = 256;
static inline int fp_blocks(long long n) {
  long long b = (n + kFpThreads - 1) / kFpThreads;
  if (b > 148LL * 16) b = 148LL * 16;
  return (int)(b < 1 ? 1 : b);
}

__device__ __forceinline__ void split_hi_lo(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// three pieces: x = p0 + p1 + p2 up to 2^-25 |x| (8 + 8 + 8 mantissa bits)
__device__ __forceinline__ void split_3(float x, __nv_bfloat16 (&p)[3]) {
  p[0] = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(p[0]);
  p[1] = __float2bfloat16_rn(r1);
  p[2] = __float2bfloat16_rn(r1 - __bfloat162float(p[1]));
}
// Term tables, smallest products first and the main term a0 b0 LAST (the consumers run the contraction
// channel-block-major): the tensor core truncates the fp32 accumulator after every MMA, and that error is proportional
// to the accumulator's magnitude, so the corrections are summed while it is still small.
//   nt = 3: a.b ~= a0 b1 + a1 b0 + a0 b0 (error 2^-16): activation blocks [a0 a1 a0], weight blocks [b1 b0 b0]
//   nt = 6: adds a0 b2 + a1 b1 + a2 b0 in front (error 2^-24): [a0 a1 a2 a0 a1 a0] x [b2 b1 b0 b1 b0 b0]
__device__ __forceinline__ int act_piece(int nt, int k) {
  constexpr int t3[3] = {0, 1, 0}, t6[6] = {0, 1, 2, 0, 1, 0};
  return nt == 3 ? t3[k] : t6[k];
}
__device__ __forceinline__ int wgt_piece(int nt, int k) {
  constexpr int t3[3] = {1, 0, 0}, t6[6] = {2, 1, 0, 1, 0, 0};
  return nt == 3 ? t3[k] : t6[k];
}

// y = x (+ res) (relu);  out_f32 = y (optional);  out3[row] = [hi(C) | hi(C) | lo(C)]
__global__ void split3_kernel(const float* __restrict__ x, const float* __restrict__ res, int relu,
                              float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out3, long long rows, int C,
                              int nt) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const long long total = rows * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    float y = x[i];
    if (res) y += res[i];
    if (relu) y = fmaxf(y, 0.f);
    if (out_f32) out_f32[i] = y;
    __nv_bfloat16 pc[3];
    split_3(y, pc);
    __nv_bfloat16* o = out3 + r * nt * C;
    for (int k = 0; k < nt; ++k) o[k * C + c] = pc[act_piece(nt, k)];
  }
}

// w f32 [N][G][C] (tap_minor == 0) or [N][C][G] (tap_minor == 1, the nn.Conv layout with G = taps)
//   -> bf16 [N][G][3C] = [hi | lo | hi] per group
__global__ void pack_split3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, long long N, int G,
                                   int C, int tap_minor, int cpad, int nt) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const long long total = N * G * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int g = (int)(r % G);
    const long long n = r / G;
    const float v = tap_minor ? w[(n * C + c) * G + g] : w[i];
    __nv_bfloat16 pc[3];
    split_3(v, pc);
    __nv_bfloat16* o = out + (n * G + g) * cpad;    // cpad >= nt*C; the tail (if any) was zero-filled by the caller
    for (int k = 0; k < nt; ++k) o[k * C + c] = pc[wgt_piece(nt, k)];
  }
}

// video (B,3,T,H,W) -> 2x2 space-to-depth with split output for the first VGG-M conv: bf16 (B,T,H/2,W/2,64) =
// [hi 16 | hi 16 | lo 16 | 0 16] per pixel, channel (ph*2+pw)*3 + c of each 16 (12 real).
template <typename T>
__global__ void video_prep_s2d_split3_kernel(const T* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int Tn,
                                             int H, int W, float mul, float add, int nt, int cpad) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int H2 = H / 2, W2 = W / 2;
  const long long total = (long long)B * Tn * H2 * W2;
  const __nv_bfloat16 z = __float2bfloat16(0.f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w2 = (int)(i % W2);
    long long r = i / W2;
    const int h2 = (int)(r % H2);
    r /= H2;
    const int t = (int)(r % Tn);
    const int b = (int)(r / Tn);
    __nv_bfloat16* o = out + i * cpad;
    for (int k = 0; k < cpad; ++k) o[k] = z;
#pragma unroll
    for (int ph = 0; ph < 2; ++ph)
#pragma unroll
      for (int pw = 0; pw < 2; ++pw)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const T* p = in + ((((long long)b * 3 + c) * Tn + t) * H + (2 * h2 + ph)) * W + 2 * w2 + pw;
          const float v = fmaf((float)(*p), mul, add);
          __nv_bfloat16 pc[3];
          split_3(v, pc);
          const int ch = (ph * 2 + pw) * 3 + c;
          for (int k = 0; k < nt; ++k) o[k * 16 + ch] = pc[act_piece(nt, k)];
        }
  }
}

// 2x2 / stride 2 max-pool over (H,W) of [F][H][W][C] float32 (nn.MaxPool3d((1,2,2),(1,2,2)), floor mode)
__global__ void maxpool2x2_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int F, int H, int W, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int P = H / 2, Q = W / 2;
  const long long total = (long long)F * P * Q * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int q = (int)(r % Q);
    r /= Q;
    const int p = (int)(r % P);
    const int f = (int)(r / P);
    const float* b = x + (((long long)f * H + 2 * p) * W + 2 * q) * C + c;
    out[i] = fmaxf(fmaxf(b[0], b[C]), fmaxf(b[(long long)W * C], b[(long long)W * C + C]));
  }
}

// As video_prep_s2d_w4 (elementwise.cu) with split output: per pixel [hi 64 | hi 64 | lo 64], each 64 = 4 taps x 12
// channels + 16 zeros.
template <typename T>
__global__ void video_prep_s2d_w4_split3_kernel(const T* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int Tn,
                                                int H, int W, float mul, float add, int nt) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int H2 = H / 2, W2 = W / 2;
  const long long total = (long long)B * Tn * H2 * W2 * 4;   // one thread per (pixel, tap jw)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int jw = (int)(i & 3);
    const long long pix = i >> 2;
    const int w2 = (int)(pix % W2);
    long long r = pix / W2;
    const int h2 = (int)(r % H2);
    r /= H2;
    const int t = (int)(r % Tn);
    const int b = (int)(r / Tn);
    const int ws = w2 + jw - 2;
    __nv_bfloat16* o = out + pix * (64 * nt);
#pragma unroll
    for (int ph = 0; ph < 2; ++ph)
#pragma unroll
      for (int pw = 0; pw < 2; ++pw)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float v = 0.f;
          if (ws >= 0 && ws < W2) {
            const T* p = in + ((((long long)b * 3 + c) * Tn + t) * H + (2 * h2 + ph)) * W + 2 * ws + pw;
            v = fmaf((float)(*p), mul, add);
          }
          __nv_bfloat16 pc[3];
          split_3(v, pc);
          const int ch = jw * 12 + (ph * 2 + pw) * 3 + c;
          for (int k = 0; k < nt; ++k) o[k * 64 + ch] = pc[act_piece(nt, k)];
        }
    if (jw == 0) {
      const __nv_bfloat16 z = __float2bfloat16(0.f);
      for (int k = 0; k < nt; ++k)
        for (int c = 48; c < 64; ++c) o[k * 64 + c] = z;
    }
  }
}

// 3x3 / stride 2 / pad 1 max-pool over [F][H][W][C] float32 (-inf padding, as nn.MaxPool)
__global__ void maxpool3s2_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int F, int H, int W, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  const long long total = (long long)F * P * Q * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int q = (int)(r % Q);
    r /= Q;
    const int p = (int)(r % P);
    const int f = (int)(r / P);
    float best = -INFINITY;
    for (int kh = 0; kh < 3; ++kh) {
      const int h = 2 * p - 1 + kh;
      if (h < 0 || h >= H) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int w = 2 * q - 1 + kw;
        if (w < 0 || w >= W) continue;
        best = fmaxf(best, x[(((long long)f * H + h) * W + w) * C + c]);
      }
    }
    out[i] = best;
  }
}

__global__ void avgpool_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int F, int HW, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const long long total = (long long)F * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long f = i / C;
    float s = 0.f;
    for (int k = 0; k < HW; ++k) s += x[(f * HW + k) * C + c];
    out[i] = s / (float)HW;
  }
}

// models/att_fusion.py:21-25 in float32 (expf, not the fast intrinsic)
__global__ void att_mix_f32_kernel(const float* __restrict__ xa, const float* __restrict__ xv,
                                   const float* __restrict__ sa, const float* __restrict__ sv, float* __restrict__ f,
                                   long long rows, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const long long total = rows * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const float hv = 1.f / (1.f + expf(-sv[r])), ha = 1.f / (1.f + expf(-sa[r]));
    const float ev = expf(hv), ea = expf(ha);
    const float wv = ev / (ev + ea), wa = ea / (ev + ea);
    f[i] = wv * xv[i] + wa * xa[i];
  }
}

// One GRU time step, float32 FFMA: block = (direction, group of 8 batch rows, 8 hidden units per warp pass).
//   gh = h_{t-1} W_hh^T + b_hh ; r,z = sigmoid(gi + gh) ; n = tanh(gi_n + r * gh_n) ; h = (1-z) n + z h_{t-1}
// gi f32 [B*T][2][3H];  w f32 [2][3H][H];  out f32 [B][T][2H] (h_t is written there and read back as h_{t-1}).
constexpr int kGruF32Rows = 8;
__global__ void __launch_bounds__(256) gru_step_f32_kernel(const float* __restrict__ gi, const float* __restrict__ w,
                                                           const float* __restrict__ b_hh, float* __restrict__ out,
                                                           int B, int T, int H, int step) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  extern __shared__ float hs[];   // [kGruF32Rows][H]
  const int dir = blockIdx.z;
  const int b0 = blockIdx.y * kGruF32Rows;
  const int t = dir == 0 ? step : T - 1 - step;
  const int tprev = dir == 0 ? t - 1 : t + 1;
  const int nrows = min(kGruF32Rows, B - b0);
  for (int i = threadIdx.x; i < kGruF32Rows * H; i += blockDim.x) {
    const int rr = i / H, k = i - rr * H;
    hs[i] = (step > 0 && rr < nrows) ? out[((long long)(b0 + rr) * T + tprev) * 2 * H + dir * H + k] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int units_per_block = (H + (int)gridDim.x - 1) / (int)gridDim.x;
  const int j0 = blockIdx.x * units_per_block;
  for (int j = j0 + warp; j < min(H, j0 + units_per_block); j += 8) {
    const float* wr = w + ((long long)dir * 3 * H + j) * H;
    const float* wz = wr + (long long)H * H;
    const float* wn = wz + (long long)H * H;
    float ar[kGruF32Rows], az[kGruF32Rows], an[kGruF32Rows];
#pragma unroll
    for (int rr = 0; rr < kGruF32Rows; ++rr) ar[rr] = az[rr] = an[rr] = 0.f;
    for (int k = lane; k < H; k += 32) {
      const float vr = __ldg(wr + k), vz = __ldg(wz + k), vn = __ldg(wn + k);
#pragma unroll
      for (int rr = 0; rr < kGruF32Rows; ++rr) {
        const float hv = hs[rr * H + k];
        ar[rr] = fmaf(vr, hv, ar[rr]);
        az[rr] = fmaf(vz, hv, az[rr]);
        an[rr] = fmaf(vn, hv, an[rr]);
      }
    }
#pragma unroll
    for (int rr = 0; rr < kGruF32Rows; ++rr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ar[rr] += __shfl_xor_sync(0xffffffffu, ar[rr], o);
        az[rr] += __shfl_xor_sync(0xffffffffu, az[rr], o);
        an[rr] += __shfl_xor_sync(0xffffffffu, an[rr], o);
      }
    }
    if (lane < nrows) {
      const int rr = lane;
      // every lane holds all sums after the xor butterfly; lane rr finishes batch row rr
      float sr = 0.f, sz = 0.f, sn = 0.f;
#pragma unroll
      for (int q = 0; q < kGruF32Rows; ++q)
        if (q == rr) { sr = ar[q]; sz = az[q]; sn = an[q]; }
      const long long row = (long long)(b0 + rr) * T + t;
      const float* g = gi + (row * 2 + dir) * 3 * H;
      const float* bh = b_hh + (long long)dir * 3 * H;
      const float r = 1.f / (1.f + expf(-(g[j] + sr + bh[j])));
      const float z = 1.f / (1.f + expf(-(g[H + j] + sz + bh[H + j])));
      const float n = tanhf(g[2 * H + j] + r * (sn + bh[2 * H + j]));
      const float hp = hs[rr * H + j];
      out[row * 2 * H + dir * H + j] = (1.f - z) * n + z * hp;
    }
  }
}

}  // namespace m3t

using namespace m3t;

extern "C" int m3t_split3_bf16(const float* x, const float* res, int relu, float* out_f32, void* out3, long long rows,
                               int C, int nterms, void* stream) {
  if (rows <= 0 || C <= 0 || (nterms != 3 && nterms != 6)) return -1;
  m3t::launch_k(split3_kernel, dim3(fp_blocks(rows * C)), dim3(kFpThreads), 0, reinterpret_cast<cudaStream_t>(stream), 
      x, res, relu, out_f32, reinterpret_cast<__nv_bfloat16*>(out3), rows, C, nterms);
  count_launch();
  return launch_status();
}

extern "C" int m3t_pack_split3_bf16(const float* w, void* out, long long N, int G, int C, int tap_minor, int cpad,
                                    int nterms, void* stream) {
  if (N <= 0 || G <= 0 || C <= 0 || (nterms != 3 && nterms != 6)) return -1;
  if (cpad <= 0) cpad = nterms * C;
  if (cpad < nterms * C) return -1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cpad > nterms * C && cudaMemsetAsync(out, 0, (size_t)N * G * cpad * 2, st) != cudaSuccess) return -21;
  m3t::launch_k(pack_split3_kernel, dim3(fp_blocks(N * G * C)), dim3(kFpThreads), 0, st, w, reinterpret_cast<__nv_bfloat16*>(out), N, G, C,
                                                                   tap_minor, cpad, nterms);
  count_launch();
  return launch_status();
}

extern "C" int m3t_video_prep_s2d_split3(const void* video, int is_u8, void* out, int B, int T, int H, int W, float mul,
                                         float add, int nterms, int cpad, void* stream) {
  if ((H | W) & 1 || (nterms != 3 && nterms != 6) || cpad < 16 * nterms) return -1;
  const long long items = (long long)B * T * (H / 2) * (W / 2);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (is_u8)
    m3t::launch_k(video_prep_s2d_split3_kernel<uint8_t>, dim3(fp_blocks(items)), dim3(kFpThreads), 0, st, 
        reinterpret_cast<const uint8_t*>(video), reinterpret_cast<__nv_bfloat16*>(out), B, T, H, W, mul, add, nterms, cpad);
  else
    m3t::launch_k(video_prep_s2d_split3_kernel<float>, dim3(fp_blocks(items)), dim3(kFpThreads), 0, st, 
        reinterpret_cast<const float*>(video), reinterpret_cast<__nv_bfloat16*>(out), B, T, H, W, mul, add, nterms, cpad);
  count_launch();
  return launch_status();
}

extern "C" int m3t_maxpool2x2_f32(const float* x, float* out, int F, int H, int W, int C, void* stream) {
  m3t::launch_k(maxpool2x2_f32_kernel, dim3(fp_blocks((long long)F * (H / 2) * (W / 2) * C)), dim3(kFpThreads), 0, reinterpret_cast<cudaStream_t>(stream), x, out, F, H, W, C);
  count_launch();
  return launch_status();
}

extern "C" int m3t_video_prep_s2d_w4_split3(const void* video, int is_u8, void* out, int B, int T, int H, int W,
                                            float mul, float add, int nterms, void* stream) {
  if ((H | W) & 1 || (nterms != 3 && nterms != 6)) return -1;
  const long long items = (long long)B * T * (H / 2) * (W / 2) * 4;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (is_u8)
    m3t::launch_k(video_prep_s2d_w4_split3_kernel<uint8_t>, dim3(fp_blocks(items)), dim3(kFpThreads), 0, st, 
        reinterpret_cast<const uint8_t*>(video), reinterpret_cast<__nv_bfloat16*>(out), B, T, H, W, mul, add, nterms);
  else
    m3t::launch_k(video_prep_s2d_w4_split3_kernel<float>, dim3(fp_blocks(items)), dim3(kFpThreads), 0, st, 
        reinterpret_cast<const float*>(video), reinterpret_cast<__nv_bfloat16*>(out), B, T, H, W, mul, add, nterms);
  count_launch();
  return launch_status();
}

extern "C" int m3t_maxpool3s2_f32(const float* x, float* out, int F, int H, int W, int C, void* stream) {
  const int P = (H - 1) / 2 + 1, Q = (W - 1) / 2 + 1;
  m3t::launch_k(maxpool3s2_f32_kernel, dim3(fp_blocks((long long)F * P * Q * C)), dim3(kFpThreads), 0, reinterpret_cast<cudaStream_t>(stream), 
      x, out, F, H, W, C);
  count_launch();
  return launch_status();
}

extern "C" int m3t_avgpool_f32(const float* x, float* out, int F, int HW, int C, void* stream) {
  m3t::launch_k(avgpool_f32_kernel, dim3(fp_blocks((long long)F * C)), dim3(kFpThreads), 0, reinterpret_cast<cudaStream_t>(stream), x, out, F,
                                                                                                           HW, C);
  count_launch();
  return launch_status();
}

extern "C" int m3t_att_mix_f32(const float* x_a, const float* x_v, const float* s_a, const float* s_v, float* f,
                               long long rows, int C, void* stream) {
  m3t::launch_k(att_mix_f32_kernel, dim3(fp_blocks(rows * C)), dim3(kFpThreads), 0, reinterpret_cast<cudaStream_t>(stream), x_a, x_v, s_a, s_v,
                                                                                                    f, rows, C);
  count_launch();
  return launch_status();
}

extern "C" int m3t_gru_fwd_f32(const float* gi, const float* w_hh, const float* b_hh, float* out, int B, int T, int H,
                               void* stream) {
  if (B <= 0 || T <= 0 || H <= 0) return -1;
  const size_t smem = (size_t)kGruF32Rows * H * sizeof(float);
  if (smem > 200 * 1024) return -7;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(gru_step_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return -20;
    attr_done = true;
  }
  const int gx = H >= 64 ? H / 64 : 1;   // 64 hidden units per block (8 warps x 8 passes)
  dim3 grid((unsigned)gx, (unsigned)((B + kGruF32Rows - 1) / kGruF32Rows), 2);
  for (int step = 0; step < T; ++step) {
    m3t::launch_k(gru_step_f32_kernel, dim3(grid), dim3(256), smem, reinterpret_cast<cudaStream_t>(stream), gi, w_hh, b_hh, out, B, T, H, step);
    count_launch();
  }
  return launch_status();
}
