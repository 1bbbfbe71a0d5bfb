// Fused log-Mel front end (reference: process/extract_melspec.py:13-20 = librosa.feature.melspectrogram(n_fft=512,
// hop_length=hop, win_length=400, n_mels=40) followed by librosa.power_to_db).
// One CTA per STFT frame:  frame the (centre-padded) 16 kHz signal -> Hann(400) zero-padded to 512 -> 512-point FFT
// in shared memory (fp32, radix-2) -> power spectrum (257 bins) -> 40-band mel projection -> 10*log10(max(S,1e-10)),
// and a running global maximum (for the top_db clip, applied by a second tiny pass).
// Algorithmic bytes per frame: 4*hop samples in (frames overlap; 4*512 touched) + 4*40 out.
#include "../../include/m3t_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace m3t {

constexpr int kNfft = 512;
constexpr int kBins = kNfft / 2 + 1;
constexpr int kMelThreads = 128;

__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

// pad_mode: 0 = constant (zeros, librosa >= 0.10), 1 = reflect (librosa < 0.10)
__global__ void __launch_bounds__(kMelThreads) melspec_kernel(const float* __restrict__ wav, long long n_samples,
                                                              int hop, int win_length, int n_mels,
                                                              const float* __restrict__ melfb /* [n_mels][257] */,
                                                              int pad_mode, float* __restrict__ out_db,
                                                              int* __restrict__ gmax_ordered, long long n_frames) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  __shared__ float re[kNfft], im[kNfft];
  __shared__ float pw[kBins + 3];
  __shared__ float red[kMelThreads / 32];
  const long long frame = blockIdx.x;
  if (frame >= n_frames) return;
  const int tid = threadIdx.x;
  const int lpad = (kNfft - win_length) / 2;
  // ---- framing + window, stored in bit-reversed order for the in-place DIT FFT ----
  for (int n = tid; n < kNfft; n += kMelThreads) {
    long long s = frame * hop + n - kNfft / 2;   // centre padding of n_fft/2 on both sides
    float x = 0.f;
    if (pad_mode == 1) {
      if (s < 0) s = -s;
      if (s >= n_samples) s = 2 * (n_samples - 1) - s;
    }
    if (s >= 0 && s < n_samples) x = __ldg(wav + s);
    float w = 0.f;
    const int k = n - lpad;
    if (k >= 0 && k < win_length) w = 0.5f - 0.5f * cospif(2.0f * (float)k / (float)win_length);  // periodic Hann
    const int r = __brev((unsigned)n) >> (32 - 9);
    re[r] = x * w;
    im[r] = 0.f;
  }
  __syncthreads();
  // ---- 9 radix-2 stages, 256 butterflies each ----
#pragma unroll 1
  for (int s = 1; s <= 9; ++s) {
    const int half = 1 << (s - 1);
    for (int b = tid; b < kNfft / 2; b += kMelThreads) {
      const int j = b & (half - 1);
      const int i0 = ((b >> (s - 1)) << s) + j;
      const int i1 = i0 + half;
      float sn, cs;
      sincospif(-(float)j / (float)half, &sn, &cs);   // exp(-i*pi*j/half)
      const float xr = re[i1], xi = im[i1];
      const float tr = xr * cs - xi * sn, ti = xr * sn + xi * cs;
      const float ur = re[i0], ui = im[i0];
      re[i0] = ur + tr; im[i0] = ui + ti;
      re[i1] = ur - tr; im[i1] = ui - ti;
    }
    __syncthreads();
  }
  for (int k = tid; k < kBins; k += kMelThreads) pw[k] = re[k] * re[k] + im[k] * im[k];
  __syncthreads();
  // ---- mel projection + dB ----
  float local_max = -INFINITY;
  for (int m = tid; m < n_mels; m += kMelThreads) {
    const float* fb = melfb + (long long)m * kBins;
    float acc = 0.f;
    for (int k = 0; k < kBins; ++k) acc = fmaf(__ldg(fb + k), pw[k], acc);
    const float db = 10.0f * log10f(fmaxf(acc, 1e-10f));
    out_db[frame * n_mels + m] = db;
    local_max = fmaxf(local_max, db);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
  if ((tid & 31) == 0) red[tid >> 5] = local_max;
  __syncthreads();
  if (tid == 0) {
    float m = red[0];
    for (int i = 1; i < kMelThreads / 32; ++i) m = fmaxf(m, red[i]);
    atomicMax(gmax_ordered, float_to_ordered(m));
  }
}

__global__ void melspec_clip_kernel(float* __restrict__ db, long long n, const int* __restrict__ gmax_ordered,
                                    float top_db) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const float floor_db = ordered_to_float(*gmax_ordered) - top_db;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    db[i] = fmaxf(db[i], floor_db);
}

// 200-d stacked audio features (reference: models/dataset.py:83-95): out[t][j*n_mels + m] = mel[3*(start+t)+j][m],
// j < 5, zero beyond the end of the spectrogram.
__global__ void mel_stack_kernel(const float* __restrict__ mel, long long n_frames, int n_mels, long long start,
                                 int w_len, float* __restrict__ out) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  const long long total = (long long)w_len * 5 * n_mels;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i % n_mels);
    const long long r = i / n_mels;
    const int j = (int)(r % 5);
    const long long t = r / 5;
    const long long f = (start + t) * 3 + j;
    out[i] = f < n_frames ? mel[f * n_mels + m] : 0.f;
  }
}

}  // namespace m3t

using namespace m3t;

extern "C" int m3t_logmel(const float* wav, long long n_samples, int hop, int win_length, int n_mels,
                          const float* mel_fb, int pad_mode, float top_db, float* out_db, int* scratch,
                          void* stream) {
  if (n_samples <= 0 || hop <= 0 || win_length > kNfft || n_mels <= 0) return -1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long n_frames = 1 + n_samples / hop;
  // 0x80808080 orders below every finite dB value (it decodes to about -3e38)
  if (cudaMemsetAsync(scratch, 0x80, sizeof(int), st) != cudaSuccess) return -22;
  m3t::launch_k(melspec_kernel, dim3((unsigned)n_frames), dim3(kMelThreads), 0, st, wav, n_samples, hop, win_length, n_mels, mel_fb, pad_mode,
                                                           out_db, scratch, n_frames);
  count_launch();
  if (top_db > 0.f) {
    const long long n = n_frames * n_mels;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    m3t::launch_k(melspec_clip_kernel, dim3((int)blocks), dim3(256), 0, st, out_db, n, scratch, top_db);
    count_launch();
  }
  return launch_status();
}

extern "C" int m3t_mel_stack(const float* mel, long long n_frames, int n_mels, long long start, int w_len,
                             float* out, void* stream) {
  const long long total = (long long)w_len * 5 * n_mels;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  m3t::launch_k(mel_stack_kernel, dim3((int)blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), mel, n_frames, n_mels, start,
                                                                                    w_len, out);
  count_launch();
  return launch_status();
}
