// 3x3 / stride 1 / pad 1 convolution for the 64 -> 64 channel layers (ResNet layer1 fprop and dgrad): persistent,
// filter-resident, halo-tile implicit GEMM.
//
// Why: with one im2col load per filter tap (umma_kernel.cuh) a 128x64 tile moves 9 x (16 KB + 8 KB) through L2 for
// 4.7 MMAC - 192 B/cycle/SM demanded against ~42 B/cycle/SM of L2->SM bandwidth, i.e. L2-bound at ~20 % of the MMA
// rate (measured 251 TFLOP/s).  Here
//   * the 9 filter taps (9 x [64 co][64 ci] bf16 = 72 KB) are loaded ONCE per CTA and stay in shared memory;
//   * per tile ONE tiled-TMA box brings the input halo: (TR+3) rows x (W+2) columns x 64 channels of one image, with
//     the zero padding produced by TMA out-of-bounds fill (coordinates start at -1);
//   * an output tile is 128 consecutive positions of the W-padded row-major pixel space, so filter tap (r,s) is the
//     SAME smem tile read from row offset r*(W+2)+s: the UMMA descriptor start address simply moves by that many
//     128-byte rows (the SWIZZLE_128B pattern is a function of the absolute smem address, so any row offset is
//     legal with base_offset = 0 - verified on hardware by m3t_debug_rowshift);
//   * CTAs are persistent (one per SM) with a double-buffered TMEM accumulator, so the epilogue of tile i overlaps
//     the MMAs of tile i+1; BatchNorm statistics are accumulated across all tiles of a CTA and flushed once.
// L2->SM traffic per tile drops from 216 KB to 26 KB; 2 of every W+2 positions (and the tail of the last row block)
// are computed and discarded.
#include "../../include/m3t_b200.h"
#include "common.cuh"
#include "tmap.cuh"
#include "ptx.cuh"

namespace m3t {

constexpr int kHaloThreads = 192;
constexpr int kHaloStages = 4;

struct HaloParams {
  int F, H, W;     // images, height, width; Cin = Cout = 64
  int Wp;          // W + 2
  int TR;          // output rows per tile (TR * Wp <= 128)
  int tiles_per_img, num_tiles;
  int a_stage;     // bytes per A stage (1024-aligned)
  int box_bytes;   // bytes one halo box delivers
  __nv_bfloat16* y;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  int relu;
  float* stats;    // [2][64] or null
  long long det_stride;   // != 0: statistics of CTA b go to stats + (1 + b) * det_stride (UmmaParams::det_stride)
};

template <int N>
__device__ __forceinline__ void halo_butterfly(float (&v)[32], uint32_t lane) {
  if constexpr (N >= 1) {
    const bool upper = (lane & N) != 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float send = upper ? v[i] : v[i + N];
      const float keep = upper ? v[i + N] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, N);
    }
    halo_butterfly<N / 2>(v, lane);
  }
}
template <>
__device__ __forceinline__ void halo_butterfly<0>(float (&)[32], uint32_t) {}

__global__ void __launch_bounds__(kHaloThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                    const HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;                                   // 9 x 8 KB
  uint8_t* sA = smem + 9 * 8192;                        // kHaloStages x a_stage
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sA + kHaloStages * p.a_stage);
  uint64_t* empty_bar = full_bar + kHaloStages;
  uint64_t* w_bar = empty_bar + kHaloStages;
  uint64_t* tmem_full = w_bar + 1;                      // [2]
  uint64_t* tmem_empty = tmem_full + 2;                 // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* stat_smem = reinterpret_cast<float*>(tmem_slot + 2);   // [4 warps][2][64]

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < kHaloStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(w_bar, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  for (int i = threadIdx.x; i < 4 * 2 * 64; i += kHaloThreads) stat_smem[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barriers, TMEM and descriptor prefetch above touch nothing a previous kernel wrote: they overlap its tail
  m3t::pdl_wait();
  m3t::pdl_launch();

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w_bar, 9 * 8192);
      for (int t = 0; t < 9; ++t) tma_load_2d(&tmW, w_bar, sW + t * 8192, t * 64, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int stage = it % kHaloStages;
        const uint32_t phase = (it / kHaloStages) & 1;
        mbar_wait(&empty_bar[stage], phase ^ 1, 400 + stage);
        const int f = tile / p.tiles_per_img;
        const int h0 = (tile - f * p.tiles_per_img) * p.TR;
        mbar_arrive_expect_tx(&full_bar[stage], p.box_bytes);
        tma_load_4d(&tmX, &full_bar[stage], sA + stage * p.a_stage, 0, -1, h0 - 1, f);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      mbar_wait(w_bar, 0, 410);
      const uint32_t w_base = smem_u32(sW);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int stage = it % kHaloStages;
        const uint32_t phase = (it / kHaloStages) & 1;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 420 + acc);
        mbar_wait(&full_bar[stage], phase, 430 + stage);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + stage * p.a_stage);
#pragma unroll 1
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const uint32_t a_tap = a_base + (uint32_t)(r * p.Wp + s) * 128u;
            const uint32_t b_tap = w_base + (uint32_t)(r * 3 + s) * 8192u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adesc = make_smem_desc(a_tap + k * 32, 16, 1024, SWZ_128B);
              const uint64_t bdesc = make_smem_desc(b_tap + k * 32, 16, 1024, SWZ_128B);
              umma_bf16(tmem_base + acc * 64, adesc, bdesc, idesc, (r | s | k) != 0 ? 1u : 0u);
            }
          }
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + (int)lane;
    const int hh = row / p.Wp, ww = row - hh * p.Wp;
    const bool want_stats = p.stats != nullptr;
    // BatchNorm statistics: every thread owns one accumulator row, so it keeps running per-column sums of ITS rows
    // over all tiles of this persistent CTA (2 x 64 registers); rows are reduced across lanes once, at the end.
    float st1[64], st2[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) st1[i] = st2[i] = 0.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int f = tile / p.tiles_per_img;
      const int h0 = (tile - f * p.tiles_per_img) * p.TR;
      const bool ok = hh < p.TR && ww < p.W && (h0 + hh) < p.H;
      const long long pix = ((long long)f * p.H + h0 + hh) * p.W + ww;
      mbar_wait(&tmem_full[acc], acc_phase, 440 + acc);
      tc_fence_after();
      const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 64;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
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
        if (c0 == 32) {   // accumulator fully read: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          mbar_arrive(&tmem_empty[acc]);
        }
        if (ok) {
          float o[32];
          epi_scale_shift32(o, v, p.scale, p.shift, c0);
          if (p.residual) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + pix * 64 + c0);
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
          uint4* op = reinterpret_cast<uint4*>(p.y + pix * 64 + c0);
#pragma unroll
          for (int g = 0; g < 4; ++g)
            op[g] = make_uint4(pack_bf16x2(o[8 * g], o[8 * g + 1]), pack_bf16x2(o[8 * g + 2], o[8 * g + 3]),
                               pack_bf16x2(o[8 * g + 4], o[8 * g + 5]), pack_bf16x2(o[8 * g + 6], o[8 * g + 7]));
        }
        if (want_stats && ok) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            st1[c0 + i] += v[i];
            st2[c0 + i] = fmaf(v[i], v[i], st2[c0 + i]);
          }
        }
      }
    }
    if (want_stats) {
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        float s1[32], s2[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          s1[i] = st1[c0 + i];
          s2[i] = st2[c0 + i];
        }
        halo_butterfly<16>(s1, lane);
        halo_butterfly<16>(s2, lane);
        stat_smem[(quad * 2 + 0) * 64 + c0 + lane] = s1[0];
        stat_smem[(quad * 2 + 1) * 64 + c0 + lane] = s2[0];
      }
      named_bar_sync(1, 128);
      for (int i = threadIdx.x - 64; i < 2 * 64; i += 128) {
        const int which = i / 64, c = i - which * 64;
        const float s = stat_smem[(0 * 2 + which) * 64 + c] + stat_smem[(1 * 2 + which) * 64 + c] +
                        stat_smem[(2 * 2 + which) * 64 + c] + stat_smem[(3 * 2 + which) * 64 + c];
        atomicAdd(p.stats + (p.det_stride ? (1 + (long long)blockIdx.x) * p.det_stride : 0) + which * 64 + c, s);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}


// ------------------------------------------------------------------------------------------------------------
// Stem forward over the W-unrolled space-to-depth image: (5,4,1) filter, 64 -> 64 channels (reference:
// models/backbone.py:328).  Same halo idea along H: a tile is TR = 8 whole image rows (448 positions = four 128-row
// accumulators, the last one overlapping the third); per temporal tap kt ONE box brings rows h0-2 .. h0+TR of frame
// t+kt-2 (zero-filled outside the image / clip) together with that tap's 4 x [64][64] filter slices, and the four
// vertical taps jh are row-shifted views (jh * W2 rows) of the box.  L2->SM traffic per 128 positions drops from
// 480 KB (20 taps x (16 + 8) KB) to 159 KB.  Persistent CTAs, double-buffered TMEM (2 x 256 columns).
// ------------------------------------------------------------------------------------------------------------
struct StemHaloParams {
  int B, T, H2, W2;
  int TR, npos, nsub;           // rows per tile, positions per tile, 128-row accumulators per tile
  int tiles_per_img, num_tiles;
  int a_bytes, stage_bytes;
  __nv_bfloat16* y;
  const float* scale;
  const float* shift;
  int relu;
  float* stats;
  long long det_stride;
};

__global__ void __launch_bounds__(kHaloThreads, 1)
stem_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                 const StemHaloParams p) {
  constexpr int STAGES = 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * p.stage_bytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* stat_smem = reinterpret_cast<float*>(tmem_slot + 2);   // [4 warps][2][64]
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < 4 * 2 * 64; i += kHaloThreads) stat_smem[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barriers, TMEM and descriptor prefetch above touch nothing a previous kernel wrote: they overlap its tail
  m3t::pdl_wait();
  m3t::pdl_launch();

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int f = tile / p.tiles_per_img;
        const int h0 = (tile - f * p.tiles_per_img) * p.TR;
        const int b = f / p.T, t = f - b * p.T;
        for (int kt = 0; kt < 5; ++kt, ++it) {
          const int stage = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1;
          mbar_wait(&empty_bar[stage], phase ^ 1, 600 + stage);
          uint8_t* sA = smem + stage * p.stage_bytes;
          uint8_t* sB = sA + p.a_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], p.a_bytes + 4 * 8192);
          tma_load_5d(&tmX, &full_bar[stage], sA, 0, 0, h0 - 2, t + kt - 2, b);
#pragma unroll
          for (int jh = 0; jh < 4; ++jh) tma_load_2d(&tmW, &full_bar[stage], sB + jh * 8192, (kt * 4 + jh) * 64, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      int it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
        const int acc = tcount & 1;
        const uint32_t acc_phase = (tcount >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 620 + acc);
        for (int kt = 0; kt < 5; ++kt, ++it) {
          const int stage = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1;
          mbar_wait(&full_bar[stage], phase, 630 + stage);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * p.stage_bytes);
          const uint32_t b_base = a_base + p.a_bytes;
#pragma unroll 1
          for (int jh = 0; jh < 4; ++jh) {
#pragma unroll 1
            for (int j = 0; j < p.nsub; ++j) {
              const int sub_start = min(128 * j, p.npos - 128);
              const uint32_t a_tap = a_base + (uint32_t)(jh * p.W2 + sub_start) * 128u;
              const uint32_t b_tap = b_base + (uint32_t)jh * 8192u;
              // channels 48..63 of every tap are structural zeros (m3t_video_prep_s2d_w4): three K steps, not four
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                const uint64_t adesc = make_smem_desc(a_tap + k * 32, 16, 1024, SWZ_128B);
                const uint64_t bdesc = make_smem_desc(b_tap + k * 32, 16, 1024, SWZ_128B);
                umma_bf16(tmem_base + acc * 256 + j * 64, adesc, bdesc, idesc, (kt | jh | k) != 0 ? 1u : 0u);
              }
            }
          }
          umma_commit(&empty_bar[stage]);
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + (int)lane;
    const bool want_stats = p.stats != nullptr;
    float st1[64], st2[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) st1[i] = st2[i] = 0.f;
    int tcount = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
      const int acc = tcount & 1;
      const uint32_t acc_phase = (tcount >> 1) & 1;
      const int f = tile / p.tiles_per_img;
      const int h0 = (tile - f * p.tiles_per_img) * p.TR;
      mbar_wait(&tmem_full[acc], acc_phase, 640 + acc);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < p.nsub; ++j) {
        const int sub_start = min(128 * j, p.npos - 128);
        const int q = sub_start + row;
        const int hh = q / p.W2, ww = q - hh * p.W2;
        const bool ok = q >= 128 * j && (h0 + hh) < p.H2;
        const long long pix = ((long long)f * p.H2 + h0 + hh) * p.W2 + ww;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256 + j * 64;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
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
          if (j == p.nsub - 1 && c0 == 32) {   // whole tile read: release the TMEM buffer
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
          }
          if (ok) {
            float o[32];
            epi_scale_shift32(o, v, p.scale, p.shift, c0);
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = fmaxf(o[i], 0.f);
            }
            uint4* op = reinterpret_cast<uint4*>(p.y + pix * 64 + c0);
#pragma unroll
            for (int g = 0; g < 4; ++g)
              op[g] = make_uint4(pack_bf16x2(o[8 * g], o[8 * g + 1]), pack_bf16x2(o[8 * g + 2], o[8 * g + 3]),
                                 pack_bf16x2(o[8 * g + 4], o[8 * g + 5]), pack_bf16x2(o[8 * g + 6], o[8 * g + 7]));
            if (want_stats) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                st1[c0 + i] += v[i];
                st2[c0 + i] = fmaf(v[i], v[i], st2[c0 + i]);
              }
            }
          }
        }
      }
    }
    if (want_stats) {
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        float s1[32], s2[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          s1[i] = st1[c0 + i];
          s2[i] = st2[c0 + i];
        }
        halo_butterfly<16>(s1, lane);
        halo_butterfly<16>(s2, lane);
        stat_smem[(quad * 2 + 0) * 64 + c0 + lane] = s1[0];
        stat_smem[(quad * 2 + 1) * 64 + c0 + lane] = s2[0];
      }
      named_bar_sync(1, 128);
      for (int i = threadIdx.x - 64; i < 2 * 64; i += 128) {
        const int which = i / 64, c = i - which * 64;
        const float s = stat_smem[(0 * 2 + which) * 64 + c] + stat_smem[(1 * 2 + which) * 64 + c] +
                        stat_smem[(2 * 2 + which) * 64 + c] + stat_smem[(3 * 2 + which) * 64 + c];
        atomicAdd(p.stats + (p.det_stride ? (1 + (long long)blockIdx.x) * p.det_stride : 0) + which * 64 + c, s);
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

// x: bf16 [F][H][W][64], w_packed: bf16 [64][9*64] (tap-major), y: bf16 [F][H][W][64]
extern "C" int m3t_conv3x3_c64_halo(const void* x, const void* w_packed, void* y, int F, int H, int W,
                                    const float* scale, const float* shift, const void* residual, int relu,
                                    float* stats, void* stream) {
  if (F <= 0 || H <= 0 || W <= 0 || W + 2 > 64) return -1;
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.F = F; p.H = H; p.W = W;
  p.Wp = W + 2;
  p.TR = 128 / p.Wp;
  if (p.TR < 1) return -1;
  if (p.TR > H) p.TR = H;
  p.tiles_per_img = (H + p.TR - 1) / p.TR;
  p.num_tiles = F * p.tiles_per_img;
  // rows of the halo box: enough for the largest row offset (2*Wp + 2) plus 128 positions
  const int box_rows = (2 * p.Wp + 2 + 128 + p.Wp - 1) / p.Wp;
  if (box_rows > 256 || p.Wp > 256) return -1;
  p.box_bytes = box_rows * p.Wp * 128;
  p.a_stage = (p.box_bytes + 1023) / 1024 * 1024;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.scale = scale; p.shift = shift;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.relu = relu & 1; p.stats = stats;
  p.det_stride = (relu & 256) && stats ? 2 * 64 : 0;      // bit 8: deterministic statistics (m3t_det_reduce follows)
  CUtensorMap tmX, tmW;
  uint64_t dims[4] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)F};
  uint64_t strides[3] = {128, (uint64_t)W * 128, (uint64_t)H * W * 128};
  uint32_t box[4] = {64, (uint32_t)p.Wp, (uint32_t)box_rows, 1};
  int rc = make_tmap_tiled_bf16(&tmX, x, 4, dims, strides, box, 128);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmW, w_packed, 576, 64, 576, 64, 64);
  if (rc) return rc;
  const int smem = 9 * 8192 + kHaloStages * p.a_stage + (2 * kHaloStages + 5) * 8 + 16 + 4 * 2 * 64 * 4 + 1024;
  if (smem > 227 * 1024) return -7;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(conv3x3_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
        cudaSuccess)
      return -20;
    attr_done = true;
  }
  const int sms = m3t::usable_sms();
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  m3t::launch_k(conv3x3_halo_kernel, dim3(grid), dim3(kHaloThreads), smem, reinterpret_cast<cudaStream_t>(stream), tmX, tmW, p);
  count_launch();
  return launch_status();
}

// xs: bf16 [B][T][H2][W2][64] (W-unrolled s2d), w_packed: bf16 [64][20*64] ((kt,jh) taps), y: bf16 [B*T][H2][W2][64]
extern "C" int m3t_stem_fprop_halo(const void* xs, const void* w_packed, void* y, int B, int T, int H2, int W2,
                                   const float* scale, const float* shift, int relu, float* stats, void* stream) {
  if (B <= 0 || T <= 0 || H2 <= 0 || W2 <= 0 || W2 > 256) return -1;
  StemHaloParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.T = T; p.H2 = H2; p.W2 = W2;
  // rows per tile: as many as fit (box + 4 filter slices) x 2 stages in shared memory, at most 4 accumulators
  int TR = 0;
  for (int t = 16; t >= 1; --t) {
    const int npos = t * W2;
    const int abytes = ((t + 3) * W2 * 128 + 1023) / 1024 * 1024;
    if (npos >= 128 && (npos + 127) / 128 <= 4 && 2 * (abytes + 4 * 8192) + 4096 <= 227 * 1024 - 1024) { TR = t; break; }
  }
  if (TR == 0) return -8;
  if (TR > H2) return -8;
  p.TR = TR;
  p.npos = TR * W2;
  p.nsub = (p.npos + 127) / 128;
  p.tiles_per_img = (H2 + TR - 1) / TR;
  p.num_tiles = B * T * p.tiles_per_img;
  p.a_bytes = ((TR + 3) * W2 * 128 + 1023) / 1024 * 1024;
  p.stage_bytes = p.a_bytes + 4 * 8192;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.scale = scale; p.shift = shift; p.relu = relu & 1; p.stats = stats;
  p.det_stride = (relu & 256) && stats ? 2 * 64 : 0;
  if ((TR + 3) * W2 * 128 != p.a_bytes) return -8;   // the box must fill the stage exactly (expect_tx accounting)
  CUtensorMap tmX, tmW;
  uint64_t xd[5] = {64, (uint64_t)W2, (uint64_t)H2, (uint64_t)T, (uint64_t)B};
  uint64_t xst[4] = {128, (uint64_t)W2 * 128, (uint64_t)H2 * W2 * 128, (uint64_t)T * H2 * W2 * 128};
  uint32_t xb[5] = {64, (uint32_t)W2, (uint32_t)(TR + 3), 1, 1};
  int rc = make_tmap_tiled_bf16(&tmX, xs, 5, xd, xst, xb, 128);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmW, w_packed, 1280, 64, 1280, 64, 64);
  if (rc) return rc;
  const int smem = 2 * p.stage_bytes + (2 * 2 + 4) * 8 + 16 + 4 * 2 * 64 * 4 + 1024;
  if (smem > 227 * 1024) return -7;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(stem_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return -20;
    attr_done = true;
  }
  const int sms = m3t::usable_sms();
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  m3t::launch_k(stem_halo_kernel, dim3(grid), dim3(kHaloThreads), smem, reinterpret_cast<cudaStream_t>(stream), tmX, tmW, p);
  count_launch();
  return launch_status();
}
