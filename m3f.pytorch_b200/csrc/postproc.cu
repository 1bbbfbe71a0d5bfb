// Evaluation post-processing on the device: the steps right after the model in the reference's
// eval flow, which it runs as host-side Python loops over per-video dicts:
//   * overlap-add of half-stride window predictions into per-video frame tracks
//     (models/model.py:281-297 validation_end, :358-366 test_end);
//   * Wiener smoothing of every track, window 35 (get_smoothed_ccc.py:15-16 -> models/utils.py:29-33 ->
//     scipy.signal.wiener, which promotes the float32 predictions to float64 through its float64 ones() kernel);
//   * masked per-video and global concordance correlation (models/utils.py:19-21, get_smoothed_ccc.py:17-26).
// All videos are processed at once as ragged sequences: `seq_off[v] .. seq_off[v+1]` are the frames of video v in
// one flat [frames][C] array.  Arithmetic type: f32 for the overlap-add (as the reference's torch tensors), f64 for
// the filter and the correlation (as scipy / numpy do on these inputs).  HBM-bound, tiny; one thread per element.
#include "../../include/m3t_b200.h"
#include "common.cuh"

namespace m3t {

constexpr int kPpThreads = 256;
static inline int pp_blocks(long long n) {
  long long b = (n + kPpThreads - 1) / kPpThreads;
  if (b > 148LL * 16) b = 148LL * 16;
  return (int)(b < 1 ? 1 : b);
}

// out[(seg_base[s] + seg_start[s] + i) * C + c] += pred[(s*L + i) * C + c]   for i < seg_len[s]
__global__ void overlap_add_kernel(const float* __restrict__ pred, const int* __restrict__ seg_start,
                                   const int* __restrict__ seg_len, const long long* __restrict__ seg_base,
                                   float* __restrict__ out, long long S, int L, int C) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const long long total = S * L * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long r = i / C;
    const int t = (int)(r % L);
    const long long s = r / L;
    if (t < seg_len[s]) atomicAdd(out + (seg_base[s] + seg_start[s] + t) * C + c, pred[i]);
  }
}

// frames [window/2, n) of every video were covered by two windows: halve them (models/model.py:295-298, :364-365)
__global__ void overlap_halve_kernel(float* __restrict__ out, const long long* __restrict__ seq_off, int V, int C,
                                     int half) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int v = blockIdx.x;
  if (v >= V) return;
  const long long lo = (seq_off[v] + half) * C, hi = seq_off[v + 1] * C;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) out[i] *= 0.5f;
}

// Pass 1 of scipy.signal.wiener: local mean / variance over a centred window with zero padding ('same' correlation
// with ones(W) / W), and the per-sequence, per-channel sum of the local variances (the noise estimate's numerator).
__global__ void wiener_stats_kernel(const float* __restrict__ x, const long long* __restrict__ seq_off, int V, int C,
                                    int W, double* __restrict__ lmean, double* __restrict__ lvar,
                                    double* __restrict__ noise_sum) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int v = blockIdx.y;
  const long long lo = seq_off[v], hi = seq_off[v + 1];
  const long long n = hi - lo;
  const int half = W / 2;
  extern __shared__ double sh[];   // [C] block partials
  for (int i = threadIdx.x; i < C; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n * C;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long t = i / C;
    double s1 = 0.0, s2 = 0.0;
    const long long a = t - half < 0 ? 0 : t - half;
    const long long b = t + half >= n ? n - 1 : t + half;
    for (long long j = a; j <= b; ++j) {
      const float xf = x[(lo + j) * C + c];
      const float x2 = __fmul_rn(xf, xf);     // scipy squares the float32 input BEFORE the float64 correlation
      s1 += (double)xf;
      s2 += (double)x2;
    }
    const double m = s1 / (double)W;
    const double var = s2 / (double)W - m * m;
    lmean[(lo + t) * C + c] = m;
    lvar[(lo + t) * C + c] = var;
    atomicAdd(&sh[c], var);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(noise_sum + (long long)v * C + i, sh[i]);
}

// Pass 2: out = lVar < noise ? lMean : (x - lMean) * (1 - noise / lVar) + lMean,  noise = mean(lVar) per sequence
__global__ void wiener_apply_kernel(const float* __restrict__ x, const long long* __restrict__ seq_off, int V, int C,
                                    const double* __restrict__ lmean, const double* __restrict__ lvar,
                                    const double* __restrict__ noise_sum, double* __restrict__ out) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int v = blockIdx.y;
  const long long lo = seq_off[v], hi = seq_off[v + 1];
  const long long n = hi - lo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n * C;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long e = lo * C + i;
    const double noise = noise_sum[(long long)v * C + c] / (double)n;
    const double m = lmean[e], var = lvar[e];
    double res = ((double)x[e] - m);
    res *= (1.0 - noise / var);
    res += m;
    out[e] = var < noise ? m : res;
  }
}

// Per (video, channel) moments over the valid frames: {n, sum a, sum b, sum a^2, sum b^2, sum ab}, a = prediction
// (f64), b = ground truth (f32); valid = every ground-truth channel of the frame >= -1 (get_smoothed_ccc.py:21).
__global__ void ccc_moments_kernel(const double* __restrict__ pred, const float* __restrict__ gt,
                                   const long long* __restrict__ seq_off, int C, double* __restrict__ mom) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int v = blockIdx.x, c = blockIdx.y;
  const long long lo = seq_off[v], hi = seq_off[v + 1];
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (long long t = lo + threadIdx.x; t < hi; t += blockDim.x) {
    bool valid = true;
    for (int k = 0; k < C; ++k) valid = valid && gt[t * C + k] >= -1.f;
    if (valid) {
      const double a = pred[t * C + c], b = (double)gt[t * C + c];
      acc[0] += 1.0; acc[1] += a; acc[2] += b; acc[3] += a * a; acc[4] += b * b; acc[5] += a * b;
    }
  }
  __shared__ double sh[6][kPpThreads];
  for (int k = 0; k < 6; ++k) sh[k][threadIdx.x] = acc[k];
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s)
      for (int k = 0; k < 6; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x < 6) mom[((long long)v * C + c) * 6 + threadIdx.x] = sh[threadIdx.x][0];
}

}  // namespace m3t

using namespace m3t;

extern "C" int m3t_overlap_add_f32(const float* pred, const int* seg_start, const int* seg_len,
                                   const long long* seg_base, const long long* seq_off, float* out, long long S, int L,
                                   int C, int V, long long total_frames, int window, void* stream) {
  if (S < 0 || L <= 0 || C <= 0 || V <= 0 || window <= 0) return -1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(out, 0, (size_t)total_frames * C * sizeof(float), st) != cudaSuccess) return -21;
  if (S > 0) {
    m3t::launch_k(overlap_add_kernel, dim3(pp_blocks(S * L * C)), dim3(kPpThreads), 0, st, pred, seg_start, seg_len, seg_base, out, S, L, C);
    count_launch();
  }
  m3t::launch_k(overlap_halve_kernel, dim3(V), dim3(kPpThreads), 0, st, out, seq_off, V, C, window / 2);
  count_launch();
  return launch_status();
}

extern "C" int m3t_wiener1d_f64(const float* x, const long long* seq_off, int V, int C, int window,
                                long long max_len, double* lmean, double* lvar, double* noise_sum, double* out,
                                void* stream) {
  if (V <= 0 || C <= 0 || window <= 0 || (window & 1) == 0 || max_len <= 0) return -1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(noise_sum, 0, (size_t)V * C * sizeof(double), st) != cudaSuccess) return -21;
  long long bx = (max_len * C + kPpThreads - 1) / kPpThreads;
  if (bx > 64) bx = 64;
  dim3 grid((unsigned)bx, (unsigned)V);
  m3t::launch_k(wiener_stats_kernel, dim3(grid), dim3(kPpThreads), C * sizeof(double), st, x, seq_off, V, C, window, lmean, lvar, noise_sum);
  m3t::launch_k(wiener_apply_kernel, dim3(grid), dim3(kPpThreads), 0, st, x, seq_off, V, C, lmean, lvar, noise_sum, out);
  count_launch(2);
  return launch_status();
}

extern "C" int m3t_ccc_moments_f64(const double* pred, const float* gt, const long long* seq_off, int V, int C,
                                   double* moments, void* stream) {
  if (V <= 0 || C <= 0) return -1;
  dim3 grid((unsigned)V, (unsigned)C);
  m3t::launch_k(ccc_moments_kernel, dim3(grid), dim3(kPpThreads), 0, reinterpret_cast<cudaStream_t>(stream), pred, gt, seq_off, C, moments);
  count_launch();
  return launch_status();
}
