// Persistent variant of the tcgen05 pipeline of umma_kernel.cuh for K-major GEMM / implicit-GEMM conv forward with
// the EPI_STORE epilogue: one CTA per SM walks over output tiles, the shared-memory ring keeps rolling across tile
// boundaries and the TMEM accumulator is double-buffered, so
//   * barrier set-up, TMEM allocation and tensor-map prefetch are paid once per SM instead of once per tile
//     (the stride-2 dgrad parity convolutions and the 1x1 downsample convs have 1-8 k-iterations per tile: the
//     one-tile-per-CTA kernel spent most of their time there, tensor pipe 4-12 %);
//   * the epilogue of tile i overlaps the loads and MMAs of tile i+1 inside the same CTA.
// Tiles are taken in the order t = blockIdx.x + i*gridDim.x with the N tile fastest; the host makes gridDim.x a
// multiple of tiles_n so a CTA keeps one column block and BatchNorm statistics accumulate in registers over all its
// tiles (flushed when the column block changes and at the end).
#pragma once
#include "umma_kernel.cuh"

namespace m3t {

template <int BN, int MT, int STAGES, int AKIND, bool TCN = false>
__global__ void __launch_bounds__(kUmmaThreads, 1)
umma_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const UmmaParams p, const int num_tiles) {
  static_assert(AKIND == A_TILED || AKIND == A_IM2COL, "K-major operands only");
  static_assert(2 * MT * BN <= 512, "two accumulator buffers must fit TMEM");
  constexpr int A_STAGE = MT * 128 * 128;
  constexpr int B_STAGE = (BN < 8 ? 8 : BN) * 128;
  constexpr int STAGE = A_STAGE + B_STAGE;
  constexpr int ACC_COLS = MT * BN;
  constexpr int TM_COLS_RAW = 2 * ACC_COLS;
  constexpr int TM_COLS = TM_COLS_RAW <= 32 ? 32 : TM_COLS_RAW <= 64 ? 64 : TM_COLS_RAW <= 128 ? 128
                          : TM_COLS_RAW <= 256 ? 256 : 512;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
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
  if (warp == 1) tmem_alloc(tmem_slot, TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barriers, TMEM and descriptor prefetch above touch nothing a previous kernel wrote: they overlap its tail
  m3t::pdl_wait();
  m3t::pdl_launch();

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.tiles_n;
        const int m_tile = tile / p.tiles_n;
        PixCoord pc[MT];
        if constexpr (AKIND == A_IM2COL) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) pc[mt] = pixel_base(p, (m_tile * MT + mt) * 128);
        }
        for (int it = 0; it < p.k_iters; ++it) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 800 + stage);
          uint8_t* sA = smem + stage * STAGE;
          uint8_t* sB = sA + A_STAGE;
          mbar_arrive_expect_tx(&full_bar[stage], STAGE);
          if constexpr (AKIND == A_TILED) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              tma_load_2d(&tmA, &full_bar[stage], sA + mt * 16384, it * kBlockK, (m_tile * MT + mt) * 128);
            tma_load_2d(&tmB, &full_bar[stage], sB, it * kBlockK, n_tile * BN);
          } else {
            int tap = it / p.cblocks;
            int cb = it - tap * p.cblocks;
            if (p.cb_major) {
              const int taps = p.T * p.R * p.S;
              cb = it / taps;
              tap = it - cb * taps;
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              im2col_load(&tmA, &full_bar[stage], sA + mt * 16384, p, pc[mt], cb * 64, tap);
            tma_load_2d(&tmB, &full_bar[stage], sB, tap * p.Cin + cb * 64, n_tile * BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0u, 0u);
      int stage = 0;
      uint32_t phase = 0;
      int ti = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
        const int acc = ti & 1;
        mbar_wait(&tmem_empty[acc], ((ti >> 1) & 1) ^ 1, 810 + acc);
        tc_fence_after();
        for (int it = 0; it < p.k_iters; ++it) {
          mbar_wait(&full_bar[stage], phase, 820 + stage);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * STAGE);
          const uint32_t sB = sA + A_STAGE;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t bdesc = make_smem_desc(sB + k * 32, 16, 1024, SWZ_128B);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const uint64_t adesc = make_smem_desc(sA + mt * 16384 + k * 32, 16, 1024, SWZ_128B);
              umma_bf16(tmem_base + acc * ACC_COLS + mt * BN, adesc, bdesc, idesc, (it | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ======================================= epilogue =======================================
    const int quad = warp & 3;
    const int row = quad * 32 + (int)lane;
    const bool want_stats = p.stats != nullptr;
    constexpr int NCH = (BN + 31) / 32;
    float st1[NCH], st2[NCH];   // sums of column (32*ci + lane) over this warp's rows of all tiles with n_tile == cur_n
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) st1[ci] = st2[ci] = 0.f;
    int cur_n = -1;
    float* const stats_base = p.stats + (p.det_stride ? (1 + 4 * (long long)blockIdx.x + quad) * p.det_stride : 0);
    auto flush_stats = [&]() {
      if (cur_n < 0) return;
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        const int c = ci * 32 + (int)lane;
        if (c < BN && cur_n * BN + c < p.N) {
          atomicAdd(stats_base + cur_n * BN + c, st1[ci]);
          atomicAdd(stats_base + p.N + cur_n * BN + c, st2[ci]);
        }
        st1[ci] = st2[ci] = 0.f;
      }
    };
    int ti = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
      const int n_tile = tile % p.tiles_n;
      const int m_tile = tile / p.tiles_n;
      const int acc = ti & 1;
      const int n0 = n_tile * BN;
      if (want_stats && n_tile != cur_n) {
        flush_stats();
        cur_n = n_tile;
      }
      mbar_wait(&tmem_full[acc], (ti >> 1) & 1, 830 + acc);
      tc_fence_after();
      const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * ACC_COLS;
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const long long m = (long long)(m_tile * MT + mt) * 128 + row;
        const bool row_ok = m < p.M;
        const long long mo = row_ok ? out_row(p, m) : 0;
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int c0 = ci * 32;
          float v[32];
          {
            uint32_t r[16];
            tmem_ld16(t_lane + mt * BN + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
            if (c0 + 16 < BN) {
              tmem_ld16(t_lane + mt * BN + c0 + 16, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[16 + i] = __uint_as_float(r[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[16 + i] = 0.f;
            }
          }
          if (mt == MT - 1 && ci == NCH - 1) {   // accumulator buffer fully read: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
          }
          if (row_ok) epi_store_chunk<TCN>(p, v, m, mo, n0 + c0, BN - c0);
          if (want_stats) {
            float s1[32], s2[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float x = row_ok ? v[i] : 0.f;
              s1[i] = x;
              s2[i] = x * x;
            }
            butterfly_colsum<16>(s1, lane);
            butterfly_colsum<16>(s2, lane);
            st1[ci] += s1[0];
            st2[ci] += s2[0];
          }
        }
      }
    }
    if (want_stats) flush_stats();
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

template <int BN, int MT, int STAGES>
constexpr int umma_persist_smem_bytes() {
  return STAGES * (MT * 128 * 128 + (BN < 8 ? 8 : BN) * 128) + (2 * STAGES + 4) * 8 + 16 + 1024;
}

}  // namespace m3t
