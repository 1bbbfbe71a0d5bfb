// Weight gradient of the 64-channel convolutions as a persistent, halo-tile tcgen05 kernel:
//   * ResNet layer1  3x3/s1/p1  64 -> 64  over [F][H][W][64]                      (9 taps, one launch)
//   * the stem over the W-unrolled space-to-depth image  (5,4,1) x 64 -> 64        (activation box resident over the
//     temporal taps, two passes: wgrad_stem_xres_kernel; fallback: one launch per temporal tap)
// The generic wgrad (umma_kernel.cuh, A_WGRAD) re-reads a 64-pixel activation tile from L2 once per filter tap: 200 KB
// of L2->SM traffic per 64 pixels for the stem, i.e. L2-bound at ~150 TFLOP/s.  Here a K-block is TR whole image rows;
// ONE TMA box brings the activation halo (TR+3 rows, zero padding by out-of-bounds fill) and ONE box brings dY in the
// same W-padded pixel space (its padding columns are zero-filled, so they contribute nothing); every filter tap is a
// row-shifted view of the activation box used as an MN-major A operand, and two taps 64 channels wide form one
// 128-row operand whose leading-dimension offset is simply the distance between the two views.  All accumulators of a
// CTA ([taps*64] x 64 fp32) stay in TMEM over all its K-blocks and are flushed once with fp32 atomics.
#include <cstdlib>
#include "../../include/m3t_b200.h"
#include "common.cuh"
#include "tmap.cuh"
#include "ptx.cuh"

namespace m3t {

constexpr int kWgThreads = 192;
constexpr int kWgMaxAcc = 5;

struct WgHaloParams {
  int n_img;            // images (frames) the K loop runs over
  int H, TR;            // image rows, rows per K-block
  int Wq;               // positions per row in the padded pixel space
  int blocks_per_img, num_kblocks;
  int ksteps;           // TR * Wq / 16
  int x_bytes, dy_bytes, stage_bytes;
  int x_w0, x_h0;       // box start offsets relative to (0, h0): (-1,-1) for 3x3/p1, (0,-2) for the stem
  int t_frames, dt;     // stem: frames per clip and temporal tap offset (image n = b*T + t reads frame t + dt); else T=0
  int n_acc;
  int shiftA[kWgMaxAcc], shiftB[kWgMaxAcc];   // row shifts of the two 64-wide atoms of accumulator a
  int tapA[kWgMaxAcc], tapB[kWgMaxAcc];       // output tap index of each atom (-1: atom unused)
  float* out;           // fp32 [64][ldo], accumulated atomically
  long long det_stride; // != 0: CTA b accumulates into out + (1 + b) * det_stride (deterministic mode)
  int ldo;
};

template <int STAGES>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                  const WgHaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * p.stage_bytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const uint32_t tm_cols = p.n_acc * 64 <= 128 ? 128 : (p.n_acc * 64 <= 256 ? 256 : 512);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tm_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barriers, TMEM and descriptor prefetch above touch nothing a previous kernel wrote: they overlap its tail
  m3t::pdl_wait();
  m3t::pdl_launch();
  const int x_stage_bytes = (p.x_bytes + 1023) / 1024 * 1024;
  int my_blocks = 0;
  for (int kb = blockIdx.x; kb < p.num_kblocks; kb += gridDim.x) ++my_blocks;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int kb = blockIdx.x; kb < p.num_kblocks; kb += gridDim.x, ++it) {
        const int stage = it % STAGES;
        const uint32_t phase = (it / STAGES) & 1;
        mbar_wait(&empty_bar[stage], phase ^ 1, 500 + stage);
        const int n = kb / p.blocks_per_img;
        const int h0 = (kb - n * p.blocks_per_img) * p.TR;
        uint8_t* sX = smem + stage * p.stage_bytes;
        uint8_t* sD = sX + x_stage_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], p.x_bytes + p.dy_bytes);
        if (p.t_frames > 0) {
          const int b = n / p.t_frames, t = n - b * p.t_frames;
          tma_load_5d(&tmX, &full_bar[stage], sX, 0, p.x_w0, h0 + p.x_h0, t + p.dt, b);
        } else {
          tma_load_5d(&tmX, &full_bar[stage], sX, 0, p.x_w0, h0 + p.x_h0, 0, n);
        }
        tma_load_4d(&tmDY, &full_bar[stage], sD, 0, 0, h0, n);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      int it = 0;
      for (int kb = blockIdx.x; kb < p.num_kblocks; kb += gridDim.x, ++it) {
        const int stage = it % STAGES;
        const uint32_t phase = (it / STAGES) & 1;
        mbar_wait(&full_bar[stage], phase, 510 + stage);
        tc_fence_after();
        const uint32_t sX = smem_u32(smem + stage * p.stage_bytes);
        const uint32_t sD = sX + x_stage_bytes;
        for (int ks = 0; ks < p.ksteps; ++ks) {
          const uint64_t bdesc = make_smem_desc(sD + ks * 2048, 8192, 1024, SWZ_128B);
          for (int a = 0; a < p.n_acc; ++a) {
            const uint32_t lbo = (uint32_t)(p.shiftB[a] - p.shiftA[a]) * 128u;
            const uint64_t adesc = make_smem_desc(sX + (uint32_t)(p.shiftA[a] + ks * 16) * 128u, lbo, 1024, SWZ_128B);
            umma_bf16(tmem_base + a * 64, adesc, bdesc, idesc, (it | ks) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[stage]);
      }
      umma_commit(tmem_full);
    }
  } else if (my_blocks > 0) {
    const int quad = warp & 3;
    const int row = quad * 32 + (int)lane;
    mbar_wait(tmem_full, 0, 520);
    tc_fence_after();
    float* const outp = p.out + (p.det_stride ? (1 + (long long)blockIdx.x) * p.det_stride : 0);
    for (int a = 0; a < p.n_acc; ++a) {
      const int tap = row < 64 ? p.tapA[a] : p.tapB[a];
      const long long m = (long long)tap * 64 + (row & 63);
      const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + a * 64;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(t_lane + c0, r);
        tmem_ld_wait();
        if (tap >= 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) atomicAdd(outp + (long long)(c0 + i) * p.ldo + m, __uint_as_float(r[i]));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tm_cols);
  }
}

static int wg_launch(const CUtensorMap& tmX, const CUtensorMap& tmDY, WgHaloParams& p, cudaStream_t st) {
  const int x_stage = (p.x_bytes + 1023) / 1024 * 1024;
  const int dy_stage = (p.dy_bytes + 1023) / 1024 * 1024;
  p.stage_bytes = x_stage + dy_stage;
  constexpr int STAGES = 2;
  const int smem = STAGES * p.stage_bytes + (2 * STAGES + 1) * 8 + 16 + 1024;
  if (smem > 227 * 1024) return -7;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(wgrad_halo_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
        cudaSuccess)
      return -20;
    attr_done = true;
  }
  const int sms = m3t::usable_sms();
  const int grid = p.num_kblocks < sms ? p.num_kblocks : sms;
  m3t::launch_k(wgrad_halo_kernel<STAGES>, dim3(grid), dim3(kWgThreads), smem, st, tmX, tmDY, p);
  count_launch();
  return launch_status();
}

// ---------------------------------------------------------------------------------------------------------------
// Stem weight gradient with the ACTIVATION box resident: one K-block = TR rows of one input frame f.  Its halo box is
// loaded once and multiplied against the dY boxes of every output frame t = f + 2 - kt that reads it (kt in
// [kt0, kt0+nkt)), so a K-block costs one X box + nkt dY boxes of L2->SM traffic for nkt*2 accumulators instead of
// nkt * (X + dY).  TMEM holds at most 8 accumulators of 64 columns, so the five temporal taps take two passes.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kXresXStages = 2;
constexpr int kXresDStages = 4;

struct WgXresParams {
  int B, T, H, TR, W;
  int blocks_per_img, num_xblocks, ksteps;
  int x_bytes, dy_bytes;
  int kt0, nkt;
  float* out;
  int ldo;
  long long det_stride;
};

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_stem_xres_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                       const WgXresParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smemD = smem + kXresXStages * p.x_bytes;
  uint64_t* xfull = reinterpret_cast<uint64_t*>(smemD + kXresDStages * p.dy_bytes);
  uint64_t* xempty = xfull + kXresXStages;
  uint64_t* dfull = xempty + kXresXStages;
  uint64_t* dempty = dfull + kXresDStages;
  uint64_t* tmem_full = dempty + kXresDStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  uint32_t* started_slot = tmem_slot + 1;
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int n_acc = p.nkt * 2;
  const uint32_t tm_cols = n_acc * 64 <= 128 ? 128 : (n_acc * 64 <= 256 ? 256 : 512);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int i = 0; i < kXresXStages; ++i) {
      mbar_init(&xfull[i], 1);
      mbar_init(&xempty[i], 1);
    }
    for (int i = 0; i < kXresDStages; ++i) {
      mbar_init(&dfull[i], 1);
      mbar_init(&dempty[i], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tm_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barriers, TMEM and descriptor prefetch above touch nothing a previous kernel wrote: they overlap its tail
  m3t::pdl_wait();
  m3t::pdl_launch();

  if (warp == 0) {
    if (lane == 0) {
      int it = 0, dit = 0;
      for (int kb = blockIdx.x; kb < p.num_xblocks; kb += gridDim.x, ++it) {
        const int xs = it % kXresXStages;
        mbar_wait(&xempty[xs], ((it / kXresXStages) & 1) ^ 1, 600 + xs);
        const int n = kb / p.blocks_per_img;
        const int h0 = (kb - n * p.blocks_per_img) * p.TR;
        const int b = n / p.T, f = n - b * p.T;
        mbar_arrive_expect_tx(&xfull[xs], p.x_bytes);
        tma_load_5d(&tmX, &xfull[xs], smem + xs * p.x_bytes, 0, 0, h0 - 2, f, b);
        for (int k = 0; k < p.nkt; ++k) {
          const int t = f + 2 - (p.kt0 + k);
          if (t < 0 || t >= p.T) continue;
          const int ds = dit % kXresDStages;
          mbar_wait(&dempty[ds], ((dit / kXresDStages) & 1) ^ 1, 610 + ds);
          mbar_arrive_expect_tx(&dfull[ds], p.dy_bytes);
          tma_load_4d(&tmDY, &dfull[ds], smemD + ds * p.dy_bytes, 0, 0, h0, b * p.T + t);
          ++dit;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      const uint32_t lbo = (uint32_t)p.W * 128u;   // the second 64-channel atom of A is the next vertical tap
      uint32_t started = 0;
      int it = 0, dit = 0;
      for (int kb = blockIdx.x; kb < p.num_xblocks; kb += gridDim.x, ++it) {
        const int xs = it % kXresXStages;
        const int n = kb / p.blocks_per_img;
        const int f = n % p.T;
        mbar_wait(&xfull[xs], (it / kXresXStages) & 1, 620 + xs);
        const uint32_t sX = smem_u32(smem + xs * p.x_bytes);
        for (int k = 0; k < p.nkt; ++k) {
          const int t = f + 2 - (p.kt0 + k);
          if (t < 0 || t >= p.T) continue;
          const int ds = dit % kXresDStages;
          mbar_wait(&dfull[ds], (dit / kXresDStages) & 1, 630 + ds);
          tc_fence_after();
          const uint32_t sD = smem_u32(smemD + ds * p.dy_bytes);
          const uint32_t first = ((started >> k) & 1u) ^ 1u;
          for (int ks = 0; ks < p.ksteps; ++ks) {
            const uint64_t bdesc = make_smem_desc(sD + ks * 2048, 8192, 1024, SWZ_128B);
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {
              const uint64_t adesc =
                  make_smem_desc(sX + (uint32_t)(2 * jp * p.W + ks * 16) * 128u, lbo, 1024, SWZ_128B);
              umma_bf16(tmem_base + (k * 2 + jp) * 64, adesc, bdesc, idesc, (first && ks == 0) ? 0u : 1u);
            }
          }
          started |= 1u << k;
          umma_commit(&dempty[ds]);
          ++dit;
        }
        umma_commit(&xempty[xs]);
      }
      *started_slot = started;
      __threadfence_block();
      umma_commit(tmem_full);
    }
  } else if (blockIdx.x < p.num_xblocks) {
    const int quad = warp & 3;
    const int row = quad * 32 + (int)lane;
    mbar_wait(tmem_full, 0, 640);
    tc_fence_after();
    float* const outp = p.out + (p.det_stride ? (1 + (long long)blockIdx.x) * p.det_stride : 0);
    const uint32_t started = *reinterpret_cast<volatile uint32_t*>(started_slot);
    for (int a = 0; a < n_acc; ++a) {
      if (!((started >> (a >> 1)) & 1u)) continue;
      const int tap = (p.kt0 + (a >> 1)) * 4 + (a & 1) * 2 + (row >> 6);
      const long long m = (long long)tap * 64 + (row & 63);
      const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + a * 64;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(t_lane + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) atomicAdd(outp + (long long)(c0 + i) * p.ldo + m, __uint_as_float(r[i]));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tm_cols);
  }
}

static int wg_xres_launch(const CUtensorMap& tmX, const CUtensorMap& tmDY, const WgXresParams& p, cudaStream_t st) {
  const int smem = kXresXStages * p.x_bytes + kXresDStages * p.dy_bytes + (2 * kXresXStages + 2 * kXresDStages + 1) * 8 +
                   16 + 1024;
  if (smem > 227 * 1024 || p.x_bytes % 1024 || p.dy_bytes % 1024) return -7;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(wgrad_stem_xres_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
        cudaSuccess)
      return -20;
    attr_done = true;
  }
  const int sms = m3t::usable_sms();
  const int grid = p.num_xblocks < sms ? p.num_xblocks : sms;
  m3t::launch_k(wgrad_stem_xres_kernel, dim3(grid), dim3(kWgThreads), smem, st, tmX, tmDY, p);
  count_launch();
  return launch_status();
}

}  // namespace m3t

using namespace m3t;

// 3x3 / stride 1 / pad 1, Cin = Cout = 64.  x, dy: bf16 [F][H][W][64]; dw_packed: fp32 [64][9*64] += (caller zero-fills)
static int wgrad3x3_c64_halo_impl(const void* x, const void* dy, float* dw_packed, int F, int H, int W, void* stream,
                                  int det);

extern "C" int m3t_wgrad3x3_c64_halo(const void* x, const void* dy, float* dw_packed, int F, int H, int W,
                                     void* stream) {
  return wgrad3x3_c64_halo_impl(x, dy, dw_packed, F, H, W, stream, 0);
}
// Deterministic variant: dw_packed is the first of (1 + m3t_det_cta_slots()) consecutive zero-filled copies; CTA b
// accumulates into copy 1 + b; the caller sums the copies in order with m3t_det_reduce.
extern "C" int m3t_wgrad3x3_c64_halo_det(const void* x, const void* dy, float* dw_packed, int F, int H, int W,
                                         void* stream) {
  return wgrad3x3_c64_halo_impl(x, dy, dw_packed, F, H, W, stream, 1);
}

static int wgrad3x3_c64_halo_impl(const void* x, const void* dy, float* dw_packed, int F, int H, int W, void* stream,
                                  int det) {
  const int Wp = W + 2;
  if (F <= 0 || H <= 0 || W <= 0 || Wp > 64) return -1;
  // rows per K-block: TR*Wp must be a multiple of 16 and the boxes must fit two pipeline stages
  int TR = 0;
  for (int t = 16; t >= 1; --t)
    if ((t * Wp) % 16 == 0 && (t + 3) * Wp * 128 + t * Wp * 128 <= 100 * 1024) { TR = t; break; }
  if (TR == 0) return -8;
  WgHaloParams p;
  memset(&p, 0, sizeof(p));
  p.n_img = F; p.H = H; p.TR = TR; p.Wq = Wp;
  p.blocks_per_img = (H + TR - 1) / TR;
  p.num_kblocks = F * p.blocks_per_img;
  p.ksteps = TR * Wp / 16;
  const int RX = TR + 3;
  p.x_bytes = RX * Wp * 128;
  p.dy_bytes = TR * Wp * 128;
  p.x_w0 = -1; p.x_h0 = -1;
  p.t_frames = 0; p.dt = 0;
  p.n_acc = 5;
  for (int a = 0; a < 5; ++a) {
    const int t0 = 2 * a, t1 = 2 * a + 1;
    p.tapA[a] = t0;
    p.shiftA[a] = (t0 / 3) * Wp + (t0 % 3);
    if (t1 < 9) { p.tapB[a] = t1; p.shiftB[a] = (t1 / 3) * Wp + (t1 % 3); }
    else { p.tapB[a] = -1; p.shiftB[a] = p.shiftA[a]; }
  }
  p.out = dw_packed; p.ldo = 576;
  p.det_stride = det ? 64LL * 576 : 0;
  CUtensorMap tmX, tmDY;
  uint64_t xd[5] = {64, (uint64_t)W, (uint64_t)H, 1, (uint64_t)F};
  uint64_t xs[4] = {128, (uint64_t)W * 128, (uint64_t)H * W * 128, (uint64_t)H * W * 128};
  uint32_t xb[5] = {64, (uint32_t)Wp, (uint32_t)RX, 1, 1};
  int rc = make_tmap_tiled_bf16(&tmX, x, 5, xd, xs, xb, 128);
  if (rc) return rc;
  uint64_t dd[4] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)F};
  uint64_t ds[3] = {128, (uint64_t)W * 128, (uint64_t)H * W * 128};
  uint32_t db[4] = {64, (uint32_t)Wp, (uint32_t)TR, 1};
  rc = make_tmap_tiled_bf16(&tmDY, dy, 4, dd, ds, db, 128);
  if (rc) return rc;
  return wg_launch(tmX, tmDY, p, reinterpret_cast<cudaStream_t>(stream));
}

// Stem: xs bf16 [B][T][H2][W2][64] (W-unrolled space-to-depth), dy bf16 [B*T][H2][W2][64];
// dw_packed fp32 [64][20*64] (tap = kt*4 + jh) += .  Two activation-resident passes (M3T_STEM_WGRAD_SPLIT taps each,
// default 3 + 2); falls back to five launches, one per temporal tap kt, when the boxes do not fit.
static int wgrad_stem_halo_impl(const void* xs, const void* dy, float* dw_packed, int B, int T, int H2, int W2,
                                void* stream, int det);

extern "C" int m3t_wgrad_stem_halo(const void* xs, const void* dy, float* dw_packed, int B, int T, int H2, int W2,
                                   void* stream) {
  return wgrad_stem_halo_impl(xs, dy, dw_packed, B, T, H2, W2, stream, 0);
}
extern "C" int m3t_wgrad_stem_halo_det(const void* xs, const void* dy, float* dw_packed, int B, int T, int H2, int W2,
                                       void* stream) {
  return wgrad_stem_halo_impl(xs, dy, dw_packed, B, T, H2, W2, stream, 1);
}

static int wgrad_stem_halo_impl(const void* xs, const void* dy, float* dw_packed, int B, int T, int H2, int W2,
                                void* stream, int det) {
  if (B <= 0 || T <= 0 || H2 <= 0 || W2 <= 0 || W2 > 256) return -1;
  int TR = 0;
  for (int t = 8; t >= 1; --t)
    if ((t * W2) % 16 == 0 && (t + 3) * W2 * 128 + t * W2 * 128 <= 100 * 1024) { TR = t; break; }
  if (TR == 0) return -8;
  const int RX = TR + 3;
  CUtensorMap tmX, tmDY;
  uint64_t xd[5] = {64, (uint64_t)W2, (uint64_t)H2, (uint64_t)T, (uint64_t)B};
  uint64_t xst[4] = {128, (uint64_t)W2 * 128, (uint64_t)H2 * W2 * 128, (uint64_t)T * H2 * W2 * 128};
  uint32_t xb[5] = {64, (uint32_t)W2, (uint32_t)RX, 1, 1};
  int rc = make_tmap_tiled_bf16(&tmX, xs, 5, xd, xst, xb, 128);
  if (rc) return rc;
  uint64_t dd[4] = {64, (uint64_t)W2, (uint64_t)H2, (uint64_t)B * T};
  uint64_t dst[3] = {128, (uint64_t)W2 * 128, (uint64_t)H2 * W2 * 128};
  uint32_t db[4] = {64, (uint32_t)W2, (uint32_t)TR, 1};
  rc = make_tmap_tiled_bf16(&tmDY, dy, 4, dd, dst, db, 128);
  if (rc) return rc;
  // Activation-resident passes (see wgrad_stem_xres_kernel) when the boxes are whole 1 KB swizzle atoms and two X
  // stages + four dY stages fit in shared memory; `split` = temporal taps per pass (TMEM holds four).
  {
    int TX = 0;
    for (int t = 8; t >= 1; --t) {
      const int xb = (t + 3) * W2 * 128, db = t * W2 * 128;
      if ((t * W2) % 16 == 0 && xb % 1024 == 0 && db % 1024 == 0 &&
          kXresXStages * xb + kXresDStages * db <= 220 * 1024) { TX = t; break; }
    }
    static int split = -1;
    if (split < 0) {
      const char* e = getenv("M3T_STEM_WGRAD_SPLIT");
      split = e ? atoi(e) : 3;
      if (split < 0 || split > 4) split = 3;
    }
    if (TX > 0 && split > 0) {
      uint32_t xb2[5] = {64, (uint32_t)W2, (uint32_t)(TX + 3), 1, 1};
      rc = make_tmap_tiled_bf16(&tmX, xs, 5, xd, xst, xb2, 128);
      if (rc) return rc;
      uint32_t db2[4] = {64, (uint32_t)W2, (uint32_t)TX, 1};
      rc = make_tmap_tiled_bf16(&tmDY, dy, 4, dd, dst, db2, 128);
      if (rc) return rc;
      for (int kt0 = 0; kt0 < 5; kt0 += split) {
        WgXresParams q;
        memset(&q, 0, sizeof(q));
        q.B = B; q.T = T; q.H = H2; q.TR = TX; q.W = W2;
        q.blocks_per_img = (H2 + TX - 1) / TX;
        q.num_xblocks = B * T * q.blocks_per_img;
        q.ksteps = TX * W2 / 16;
        q.x_bytes = (TX + 3) * W2 * 128;
        q.dy_bytes = TX * W2 * 128;
        q.kt0 = kt0; q.nkt = 5 - kt0 < split ? 5 - kt0 : split;
        q.out = dw_packed; q.ldo = 20 * 64;
        q.det_stride = det ? 64LL * 20 * 64 : 0;
        rc = wg_xres_launch(tmX, tmDY, q, reinterpret_cast<cudaStream_t>(stream));
        if (rc) return rc;
      }
      return 0;
    }
  }
  for (int kt = 0; kt < 5; ++kt) {
    WgHaloParams p;
    memset(&p, 0, sizeof(p));
    p.n_img = B * T; p.H = H2; p.TR = TR; p.Wq = W2;
    p.blocks_per_img = (H2 + TR - 1) / TR;
    p.num_kblocks = p.n_img * p.blocks_per_img;
    p.ksteps = TR * W2 / 16;
    p.x_bytes = RX * W2 * 128;
    p.dy_bytes = TR * W2 * 128;
    p.x_w0 = 0; p.x_h0 = -2;
    p.t_frames = T; p.dt = kt - 2;
    p.n_acc = 2;
    for (int a = 0; a < 2; ++a) {
      p.tapA[a] = kt * 4 + 2 * a;     p.shiftA[a] = (2 * a) * W2;
      p.tapB[a] = kt * 4 + 2 * a + 1; p.shiftB[a] = (2 * a + 1) * W2;
    }
    p.out = dw_packed; p.ldo = 20 * 64;
    p.det_stride = det ? 64LL * 20 * 64 : 0;
    rc = wg_launch(tmX, tmDY, p, reinterpret_cast<cudaStream_t>(stream));
    if (rc) return rc;
  }
  return 0;
}
