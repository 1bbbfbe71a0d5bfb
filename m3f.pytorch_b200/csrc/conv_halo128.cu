// 3x3 / stride 1 / pad 1 convolution for the 128 -> 128 channel layers (ResNet layer2 fprop and dgrad at 14x14):
// persistent halo-tile implicit GEMM, one whole image per tile.
//
// Why: the im2col kernel (umma_kernel.cuh) re-reads every input pixel once per filter tap - a 256x128 CTA tile moves
// 590 KB of activations + 295 KB of filter through L2 for 75 MFLOP, i.e. L2->SM-bound at ~540 TFLOP/s (measured).
// Here
//   * a tile is ONE image in the W-padded pixel space: (H+2) x (W+2) halo pixels x 128 channels = 2 TMA boxes of 64
//     channels (zero padding by out-of-bounds fill), double-buffered; H*(W+2) <= 256 output positions = two 128-row
//     accumulators (the second one overlaps the first when there are fewer than 256 positions);
//   * filter tap (r,s) of channel block cb is the same box read from row offset r*(W+2)+s (SWIZZLE_128B is a function
//     of the absolute shared-memory address, so the descriptor start may move by whole 128-byte rows);
//   * the filter streams through a 5-slot ring as 18 [128 co][64 ci] blocks per image (L2-resident, 295 KB), each
//     used by both accumulators; activations cost 64 KB per image instead of 1.2 MB;
//   * N = 128 MMAs, double-buffered TMEM (2 x 256 columns), BatchNorm statistics reduced per tile by a warp
//     butterfly into four per-lane column accumulators and flushed once per CTA.
#include "../../include/m3t_b200.h"
#include "common.cuh"
#include "tmap.cuh"
#include "ptx.cuh"

namespace m3t {

constexpr int kH128Threads = 192;
constexpr int kH128AStages = 2;
constexpr int kH128BSlots = 5;
constexpr int kH128BBytes = 128 * 128;   // [128 co][64 ci] bf16

struct Halo128Params {
  int F, H, W, Wp;
  int npos, nsub;      // H*Wp output positions, 128-row accumulators per image
  int cb_bytes;        // bytes of one 64-channel halo box ((H+2)*Wp*128)
  __nv_bfloat16* y;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  int relu;
  float* stats;        // [2][128] or null
  long long det_stride;   // != 0: statistics of CTA b go to stats + (1 + b) * det_stride
};

template <int N>
__device__ __forceinline__ void h128_butterfly(float (&v)[32], uint32_t lane) {
  if constexpr (N >= 1) {
    const bool upper = (lane & N) != 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float send = upper ? v[i] : v[i + N];
      const float keep = upper ? v[i + N] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, N);
    }
    h128_butterfly<N / 2>(v, lane);
  }
}
template <>
__device__ __forceinline__ void h128_butterfly<0>(float (&)[32], uint32_t) {}

__global__ void __launch_bounds__(kH128Threads, 1)
conv3x3_c128_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                         const Halo128Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_stage = 2 * p.cb_bytes;
  uint8_t* sA = smem;                                    // kH128AStages x (2 channel blocks)
  uint8_t* sB = smem + kH128AStages * a_stage;           // kH128BSlots x 16 KB
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + kH128BSlots * kH128BBytes);
  uint64_t* a_empty = a_full + kH128AStages;
  uint64_t* b_full = a_empty + kH128AStages;
  uint64_t* b_empty = b_full + kH128BSlots;
  uint64_t* tmem_full = b_empty + kH128BSlots;           // [2]
  uint64_t* tmem_empty = tmem_full + 2;                  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* stat_smem = reinterpret_cast<float*>(tmem_slot + 2);   // [4 warps][2][128]

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < kH128AStages; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kH128BSlots; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < 4 * 2 * 128; i += kH128Threads) stat_smem[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barriers, TMEM and descriptor prefetch above touch nothing a previous kernel wrote: they overlap its tail
  m3t::pdl_wait();
  m3t::pdl_launch();

  if (warp == 0) {
    if (lane == 0) {
      auto load_a = [&](int it, int f) {
        const int stage = it % kH128AStages;
        mbar_wait(&a_empty[stage], ((it / kH128AStages) & 1) ^ 1, 700 + stage);
        mbar_arrive_expect_tx(&a_full[stage], 2 * p.cb_bytes);
        tma_load_4d(&tmX, &a_full[stage], sA + stage * a_stage, 0, -1, -1, f);
        tma_load_4d(&tmX, &a_full[stage], sA + stage * a_stage + p.cb_bytes, 64, -1, -1, f);
      };
      int it = 0, bit = 0;
      if ((int)blockIdx.x < p.F) load_a(0, blockIdx.x);
      for (int f = blockIdx.x; f < p.F; f += gridDim.x, ++it) {
        for (int j = 0; j < 18; ++j, ++bit) {
          // the next image's halo is requested a few filter blocks into this one, so that it lands before it is needed
          if (j == 4 && f + (int)gridDim.x < p.F) load_a(it + 1, f + gridDim.x);
          const int slot = bit % kH128BSlots;
          mbar_wait(&b_empty[slot], ((bit / kH128BSlots) & 1) ^ 1, 710 + slot);
          mbar_arrive_expect_tx(&b_full[slot], kH128BBytes);
          tma_load_2d(&tmW, &b_full[slot], sB + slot * kH128BBytes, j * 64, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
      int it = 0, bit = 0;
      for (int f = blockIdx.x; f < p.F; f += gridDim.x, ++it) {
        const int stage = it % kH128AStages;
        const int acc = it & 1;
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1, 720 + acc);
        mbar_wait(&a_full[stage], (it / kH128AStages) & 1, 730 + stage);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + stage * a_stage);
#pragma unroll 1
        for (int j = 0; j < 18; ++j, ++bit) {
          const int tap = j >> 1, cb = j & 1;
          const int slot = bit % kH128BSlots;
          mbar_wait(&b_full[slot], (bit / kH128BSlots) & 1, 740 + slot);
          tc_fence_after();
          const uint32_t b_base = smem_u32(sB + slot * kH128BBytes);
          const uint32_t a_tap = a_base + cb * p.cb_bytes + (uint32_t)((tap / 3) * p.Wp + (tap % 3)) * 128u;
          for (int sub = 0; sub < p.nsub; ++sub) {
            const uint32_t a_sub = a_tap + (uint32_t)(sub == 0 ? 0 : p.npos - 128) * 128u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adesc = make_smem_desc(a_sub + k * 32, 16, 1024, SWZ_128B);
              const uint64_t bdesc = make_smem_desc(b_base + k * 32, 16, 1024, SWZ_128B);
              umma_bf16(tmem_base + acc * 256 + sub * 128, adesc, bdesc, idesc, (j | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&b_empty[slot]);
        }
        umma_commit(&a_empty[stage]);
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + (int)lane;
    const bool want_stats = p.stats != nullptr;
    float acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};   // column (c0 + lane) sums, c0 = 32*i
    int it = 0;
    for (int f = blockIdx.x; f < p.F; f += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&tmem_full[acc], (it >> 1) & 1, 750 + acc);
      tc_fence_after();
      for (int sub = 0; sub < p.nsub; ++sub) {
        const int base = sub == 0 ? 0 : p.npos - 128;
        const int pos = base + row;
        const int hh = pos / p.Wp, ww = pos - hh * p.Wp;
        // the second accumulator repeats the positions below 128: only its new rows are stored / counted
        const bool ok = ww < p.W && hh < p.H && (sub == 0 || pos >= 128);
        const long long pix = ((long long)f * p.H + hh) * p.W + ww;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256 + sub * 128;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c0 = ci * 32;
          float v[32];
          {
            uint32_t r[16];
            tmem_ld16(t_lane + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
            tmem_ld16(t_lane + c0 + 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[16 + i] = __uint_as_float(r[i]);
          }
          if (ci == 3 && sub == p.nsub - 1) {   // accumulators fully read: hand the TMEM buffer back
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
          }
          if (ok) {
            float o[32];
            epi_scale_shift32(o, v, p.scale, p.shift, c0);
            if (p.residual) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.residual + pix * 128 + c0);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint4 rv = __ldg(rp + g);
                const uint32_t w4[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  o[g * 8 + 2 * j] += bf16lo(w4[j]);
                  o[g * 8 + 2 * j + 1] += bf16hi(w4[j]);
                }
              }
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = fmaxf(o[i], 0.f);
            }
            uint4* op = reinterpret_cast<uint4*>(p.y + pix * 128 + c0);
#pragma unroll
            for (int g = 0; g < 4; ++g)
              op[g] = make_uint4(pack_bf16x2(o[8 * g], o[8 * g + 1]), pack_bf16x2(o[8 * g + 2], o[8 * g + 3]),
                                 pack_bf16x2(o[8 * g + 4], o[8 * g + 5]), pack_bf16x2(o[8 * g + 6], o[8 * g + 7]));
          }
          if (want_stats) {
            float s1[32], s2[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float x = ok ? v[i] : 0.f;
              s1[i] = x;
              s2[i] = x * x;
            }
            h128_butterfly<16>(s1, lane);
            h128_butterfly<16>(s2, lane);
            acc1[ci] += s1[0];
            acc2[ci] += s2[0];
          }
        }
      }
    }
    if (want_stats) {
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        stat_smem[(quad * 2 + 0) * 128 + ci * 32 + lane] = acc1[ci];
        stat_smem[(quad * 2 + 1) * 128 + ci * 32 + lane] = acc2[ci];
      }
      named_bar_sync(1, 128);
      for (int i = threadIdx.x - 64; i < 2 * 128; i += 128) {
        const int which = i / 128, c = i - which * 128;
        const float s = stat_smem[(0 * 2 + which) * 128 + c] + stat_smem[(1 * 2 + which) * 128 + c] +
                        stat_smem[(2 * 2 + which) * 128 + c] + stat_smem[(3 * 2 + which) * 128 + c];
        atomicAdd(p.stats + (p.det_stride ? (1 + (long long)blockIdx.x) * p.det_stride : 0) + which * 128 + c, s);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace m3t

using namespace m3t;

// x, y: bf16 [F][H][W][128]; w_packed: bf16 [128][9*128] (tap-major, channel-minor); epilogue as m3t_conv_fprop_bf16.
extern "C" int m3t_conv3x3_c128_halo(const void* x, const void* w_packed, void* y, int F, int H, int W,
                                     const float* scale, const float* shift, const void* residual, int relu,
                                     float* stats, void* stream) {
  if (F <= 0 || H <= 0 || W <= 0) return -1;
  Halo128Params p;
  memset(&p, 0, sizeof(p));
  p.F = F; p.H = H; p.W = W; p.Wp = W + 2;
  p.npos = H * p.Wp;
  if (p.npos < 128 || p.npos > 256) return -8;
  p.nsub = p.npos > 128 ? 2 : 1;
  const int halo_px = (H + 2) * p.Wp;
  if (halo_px > 256 || (halo_px * 128) % 1024 != 0 || p.Wp > 256 || H + 2 > 256) return -8;
  // rows read past the box (largest tap offset 2*Wp+2 beyond position npos-1) must stay inside the allocation: they
  // fall into the next channel block / stage / filter ring and only feed discarded padding columns
  p.cb_bytes = halo_px * 128;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.scale = scale; p.shift = shift;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.relu = relu & 1; p.stats = stats;
  p.det_stride = (relu & 256) && stats ? 2 * 128 : 0;
  CUtensorMap tmX, tmW;
  uint64_t dims[4] = {128, (uint64_t)W, (uint64_t)H, (uint64_t)F};
  uint64_t strides[3] = {256, (uint64_t)W * 256, (uint64_t)H * W * 256};
  uint32_t box[4] = {64, (uint32_t)p.Wp, (uint32_t)(H + 2), 1};
  int rc = make_tmap_tiled_bf16(&tmX, x, 4, dims, strides, box, 128);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmW, w_packed, 1152, 128, 1152, 64, 128);
  if (rc) return rc;
  const int smem = kH128AStages * 2 * p.cb_bytes + kH128BSlots * kH128BBytes + (2 * kH128AStages + 2 * kH128BSlots + 4) * 8 +
                   16 + 4 * 2 * 128 * 4 + 1024;
  if (smem > 227 * 1024) return -7;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(conv3x3_c128_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
        cudaSuccess)
      return -20;
    attr_done = true;
  }
  const int sms = m3t::usable_sms();
  const int grid = F < sms ? F : sms;
  m3t::launch_k(conv3x3_c128_halo_kernel, dim3(grid), dim3(kH128Threads), smem, reinterpret_cast<cudaStream_t>(stream), tmX, tmW, p);
  count_launch();
  return launch_status();
}
