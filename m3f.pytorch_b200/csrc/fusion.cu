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
  m3t::pdl_wait();
  m3t::pdl_launch();
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
  m3t::pdl_wait();
  m3t::pdl_launch();
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
  m3t::launch_k(att_mix_fwd_kernel, dim3((int)blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
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
  m3t::launch_k(att_mix_bwd_kernel, dim3((int)blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(df), reinterpret_cast<const __nv_bfloat16*>(x_a),
      reinterpret_cast<const __nv_bfloat16*>(x_v), s_a, s_v, reinterpret_cast<__nv_bfloat16*>(dx_a),
      reinterpret_cast<__nv_bfloat16*>(dx_v), ds_a, ds_v, rows, C);
  count_launch();
  return launch_status();
}

namespace m3t {

// ------------------------------------------------------------------------------------------------------------
// Training loss of the task module and its gradient in ONE launch (reference models/model.py:132-144,146-182;
// models/utils.py:6-17):   L = lambda * (1 - CCC(v_hat, v)) + (1 - lambda) * (1 - CCC(a_hat, a))
//                              [+ w_ce * mean_i( valid_i * CE(logits_i, class_i) )]          (loss 'ccc_mtl')
// CCC over the flattened local batch with the reference's mixed estimators: biased covariance, UNBIASED variances.
// y_hat [N][C] fp32 (C = 9: 7 expression logits, valence, arousal; C = 2: valence, arousal), labels fp32 [N],
// classes int64 [N], valid uint8 [N].  One CTA (N = B*T <= a few thousand rows); every reduction is a fixed-order
// shared-memory tree, so the loss and dL/dy_hat are bit-reproducible.  The eager PyTorch version of this was ~90
// launches per step (5 % of the launches of the whole training step).
// out[0] = L, out[1] = L_v, out[2] = L_a, out[3] = CE term (before w_ce); dy [N][C] = dL/dy_hat.
// ------------------------------------------------------------------------------------------------------------
constexpr int kLossThreads = 1024;

__device__ __forceinline__ double block_sum_d(double v, double* red) {
  // fixed-order tree over the block; result broadcast to every thread
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) red[32] = s;
  }
  __syncthreads();
  return red[32];
}

__global__ void __launch_bounds__(kLossThreads, 1)
av_loss_kernel(const float* __restrict__ y, const float* __restrict__ lab_v, const float* __restrict__ lab_a,
               const long long* __restrict__ cls, const unsigned char* __restrict__ valid, int N, int C, int iv, int ia,
               int n_logits, float lambda, float w_ce, float* __restrict__ out, float* __restrict__ dy) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  __shared__ double red[33];
  const int tid = threadIdx.x;
  // ---- first moments ----
  double sx[2] = {0, 0}, sy[2] = {0, 0};
  for (int i = tid; i < N; i += kLossThreads) {
    sx[0] += y[(long long)i * C + iv];
    sx[1] += y[(long long)i * C + ia];
    sy[0] += lab_v[i];
    sy[1] += lab_a[i];
  }
  double mx[2], my[2];
  for (int k = 0; k < 2; ++k) {
    mx[k] = block_sum_d(sx[k], red) / N;
    my[k] = block_sum_d(sy[k], red) / N;
  }
  // ---- centred second moments ----
  double sxx[2] = {0, 0}, syy[2] = {0, 0}, sxy[2] = {0, 0};
  for (int i = tid; i < N; i += kLossThreads) {
    const double xv = y[(long long)i * C + iv] - mx[0], xa = y[(long long)i * C + ia] - mx[1];
    const double yv = lab_v[i] - my[0], ya = lab_a[i] - my[1];
    sxx[0] += xv * xv; syy[0] += yv * yv; sxy[0] += xv * yv;
    sxx[1] += xa * xa; syy[1] += ya * ya; sxy[1] += xa * ya;
  }
  double cov[2], vx[2], vy[2], den[2], ccc[2];
  for (int k = 0; k < 2; ++k) {
    cov[k] = block_sum_d(sxy[k], red) / N;                 // biased
    vx[k] = block_sum_d(sxx[k], red) / (N - 1);            // unbiased (torch.var)
    vy[k] = block_sum_d(syy[k], red) / (N - 1);
    den[k] = vx[k] + vy[k] + (mx[k] - my[k]) * (mx[k] - my[k]);
    ccc[k] = 2.0 * cov[k] / den[k];
  }
  // ---- masked cross-entropy over the expression logits, and the whole gradient ----
  const double wk[2] = {(double)lambda, 1.0 - (double)lambda};
  double ce = 0.0;
  for (int i = tid; i < N; i += kLossThreads) {
    float* d = dy + (long long)i * C;
    const float* yi = y + (long long)i * C;
    for (int j = 0; j < C; ++j) d[j] = 0.f;
    if (n_logits > 0) {
      float m = -INFINITY;
      for (int j = 0; j < n_logits; ++j) m = fmaxf(m, yi[j]);
      float se = 0.f;
      for (int j = 0; j < n_logits; ++j) se += expf(yi[j] - m);
      const float lse = m + logf(se);
      const int c = (int)cls[i];
      const bool ok = valid[i] != 0 && c >= 0 && c < n_logits;
      if (ok) {
        ce += (double)(lse - yi[c]);
        const float g = w_ce / (float)N;
        for (int j = 0; j < n_logits; ++j) d[j] = g * (expf(yi[j] - lse) - (j == c ? 1.f : 0.f));
      }
    }
    const int col[2] = {iv, ia};
    const float* lab[2] = {lab_v, lab_a};
    for (int k = 0; k < 2; ++k) {
      const double xc = (double)yi[col[k]] - mx[k], yc = (double)lab[k][i] - my[k];
      // d ccc / d x_i = 2 [ (y_i - my)/N * D - cov * (2 (x_i - mx)/(N-1) + 2 (mx - my)/N) ] / D^2
      const double dccc = 2.0 * (yc / N * den[k] - cov[k] * (2.0 * xc / (N - 1) + 2.0 * (mx[k] - my[k]) / N)) /
                          (den[k] * den[k]);
      d[col[k]] += (float)(-wk[k] * dccc);
    }
  }
  const double ce_mean = block_sum_d(ce, red) / N;
  if (tid == 0) {
    const double lv = 1.0 - ccc[0], la = 1.0 - ccc[1];
    out[1] = (float)lv;
    out[2] = (float)la;
    out[3] = (float)ce_mean;
    out[0] = (float)(wk[0] * lv + wk[1] * la + (n_logits > 0 ? (double)w_ce * ce_mean : 0.0));
  }
}

}  // namespace m3t

extern "C" int m3t_av_loss(const float* y_hat, const float* label_v, const float* label_a, const long long* cls,
                           const unsigned char* valid, int N, int C, int idx_v, int idx_a, int n_logits, float lambda,
                           float w_ce, float* out4, float* dy, void* stream) {
  if (N < 2 || C < 2 || idx_v < 0 || idx_v >= C || idx_a < 0 || idx_a >= C || n_logits < 0 || n_logits > C) return -1;
  if (n_logits > 0 && (!cls || !valid)) return -1;
  m3t::launch_k(m3t::av_loss_kernel, dim3(1), dim3(m3t::kLossThreads), 0, reinterpret_cast<cudaStream_t>(stream), 
      y_hat, label_v, label_a, cls, valid, N, C, idx_v, idx_a, n_logits, lambda, w_ce, out4, dy);
  m3t::count_launch();
  return m3t::launch_status();
}
