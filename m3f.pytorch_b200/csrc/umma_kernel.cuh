// One warp-specialised tcgen05 pipeline for every GEMM-shaped op on the hot path:
//   * Linear / GRU input projections (plain GEMM, any operand major),
//   * Conv1d/2d/3d forward and stride-1 dgrad as implicit GEMM (TMA im2col feeds the A operand),
//   * Conv wgrad (TMA im2col feeds the M side, dY the N side, contraction over pixels, split-K).
// Layout of a CTA (192 threads):  warp 0 = TMA producer (one lane), warp 1 = TMEM owner + MMA issuer (one lane),
// warps 2..5 = epilogue (TMEM lane quadrant = warp_idx % 4).  One 128*MT x BN output tile per CTA; two CTAs
// co-reside per SM whenever shared memory and TMEM columns allow, so one CTA's epilogue overlaps the
// other's main loop.  Accumulators live in TMEM (MT*BN fp32 columns).
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace m3t {

constexpr int kUmmaThreads = 192;
constexpr int kBlockK = 64;  // contraction elements per pipeline stage (= one 128B swizzle row of bf16)

enum AKind : int { A_TILED = 0, A_IM2COL = 1, A_WGRAD = 2 };
enum EpiKind : int {
  EPI_STORE = 0,     // out[m][n] = act(acc * scale[n] + shift[n] + residual[m][n]); optional column stats
  EPI_ATOMIC_T = 1,  // out_f32[n][m] += acc           (conv wgrad, split-K)
};

struct UmmaParams {
  int M, N;           // extents of the D tile space (rows, cols)
  int k_iters;        // pipeline iterations per CTA
  int tiles_n;        // number of N tiles (blockIdx.x = m_tile * tiles_n + n_tile)
  // ---- implicit-GEMM geometry (A_IM2COL / A_WGRAD) ----
  int rank;           // tensor-map rank: 3 = (c,w,n)  4 = (c,w,h,n)  5 = (c,w,h,d,n)
  int Q, P, Z;        // output extents along w, h, d
  int sw, sh, sd;     // conv strides
  int pw, ph, pd;     // paddings (lower)
  int dw, dh, dd;     // dilations
  int S, R, T;        // filter extents along w, h, d
  int cblocks;        // Cin / CK
  int Cin;
  int atoms;          // A_WGRAD: taps * cblocks (number of CK-wide M atoms)
  // ---- epilogue ----
  void* out;
  long long ldc;      // row stride of out (elements)
  int out_f32;        // 0: bf16, 1: fp32
  const float* scale; // per-column (n) multiplier or null
  const float* shift; // per-column (n) addend (bias / folded BN shift) or null
  const __nv_bfloat16* residual;  // [M][ldr] or null
  long long ldr;
  int relu;
  float* stats;       // [2][N] (sum, sum of squares of the raw accumulator over valid rows) or null
  // ---- optional output row map (stride-2 dgrad by parity): row m = (img, a, b) over [.][op][oq] is stored at pixel
  //      img*o_img + a*o_row + b*o_px (x ldc elements) instead of m.  oq == 0: linear.
  int oq, op;
  long long o_img, o_row, o_px;
  int res_mapped;     // residual rows follow the same map (in-place accumulation)
  // K order of the implicit GEMM: 0 = tap-major (tap, channel block), 1 = channel-block-major (all taps of block 0, then
  // block 1, ...).  The fp32-parity mode puts the small correction terms in the first channel blocks and the main term
  // last, so the tensor core's per-MMA accumulator truncation acts on the full magnitude for 1/3 of the K steps only.
  int cb_major;
  // ---- fused TemporalBlock epilogue (models/tcn.py:19-33,43-46): t = drop(relu(acc*scale+shift)); out2 = t (saved for
  //      backward when a residual follows); out = residual ? relu(t + residual) : t.  pre_act != 0 selects this order
  //      (the activation BEFORE the residual); drop_thresh == 0: no dropout.
  // ---- deterministic accumulation (train.py:17 asks for reproducible training): when det_stride != 0 the fp32 atomics
  //      of this launch go to a PRIVATE copy of the destination - slot s at dst + (1 + s) * det_stride - so no two CTAs
  //      (or warps) ever add to the same address; m3t_det_reduce then sums the slots in index order.  Slot = 4 *
  //      blockIdx.x + epilogue warp for the BatchNorm statistics of the persistent kernel, the split index for wgrad.
  long long det_stride;
  int pre_act;
  unsigned drop_thresh;
  float drop_scale;
  unsigned long long drop_seed;
  const unsigned long long* drop_seed_dev;   // optional: *drop_seed_dev is added to drop_seed (a device-resident
                                             // step counter: a replayed CUDA graph draws a new mask every replay)
  __nv_bfloat16* out2;
};

__device__ __forceinline__ long long out_row(const UmmaParams& p, long long m) {
  if (p.oq == 0) return m;
  const int mi = (int)m;
  const int b = mi % p.oq;
  const int t = mi / p.oq;
  const int a = t % p.op;
  return (long long)(t / p.op) * p.o_img + (long long)a * p.o_row + (long long)b * p.o_px;
}

// Decompose a linear output-pixel index into the im2col base coordinates of filter tap 0.
struct PixCoord {
  int w, h, d, n;
};
__device__ __forceinline__ PixCoord pixel_base(const UmmaParams& p, int m) {
  PixCoord c;
  int q = m % p.Q;
  int t = m / p.Q;
  c.w = q * p.sw - p.pw;
  if (p.rank == 3) {
    c.h = 0; c.d = 0; c.n = t;
    return c;
  }
  int pp = t % p.P;
  t /= p.P;
  c.h = pp * p.sh - p.ph;
  if (p.rank == 4) {
    c.d = 0; c.n = t;
    return c;
  }
  int z = t % p.Z;
  c.d = z * p.sd - p.pd;
  c.n = t / p.Z;
  return c;
}
__device__ __forceinline__ void im2col_load(const CUtensorMap* tm, uint64_t* bar, void* dst, const UmmaParams& p,
                                            const PixCoord& c, int chan, int tap) {
  if (p.rank == 4) {
    int r = tap / p.S, s = tap - r * p.S;
    tma_load_im2col_4d(tm, bar, dst, chan, c.w, c.h, c.n, (uint16_t)(s * p.dw), (uint16_t)(r * p.dh));
  } else if (p.rank == 3) {
    tma_load_im2col_3d(tm, bar, dst, chan, c.w, c.n, (uint16_t)(tap * p.dw));
  } else {
    int rs = p.R * p.S;
    int t = tap / rs;
    int rem = tap - t * rs;
    int r = rem / p.S, s = rem - r * p.S;
    tma_load_im2col_5d(tm, bar, dst, chan, c.w, c.h, c.d, c.n, (uint16_t)(s * p.dw), (uint16_t)(r * p.dh),
                       (uint16_t)(t * p.dd));
  }
}

// Transposing butterfly: on entry every lane holds N values (one per column); on exit v[0] of lane l holds the
// sum over the 32 lanes of column l (N == 32).  31 shuffles instead of 32*5.
template <int N>
__device__ __forceinline__ void butterfly_colsum(float (&v)[32], uint32_t lane) {
  if constexpr (N >= 1) {
    const bool upper = (lane & N) != 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float send = upper ? v[i] : v[i + N];
      float keep = upper ? v[i + N] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, N);
    }
    butterfly_colsum<N / 2>(v, lane);
  }
}
template <>
__device__ __forceinline__ void butterfly_colsum<0>(float (&)[32], uint32_t) {}

// Epilogue output of one 32-column chunk of one accumulator row: v * scale + shift (+ residual) (ReLU) -> bf16 / fp32.
// m = row of D, mo = output row (after the optional scatter map), nc0 = first global column of the chunk, cols_left =
// columns of the CTA tile from this chunk on.
// TCN = true compiles the fused TemporalBlock epilogue in (UmmaParams::pre_act ...); the ordinary instantiations do not
// carry its registers (the 256-column persistent kernel went from 160 to 220 registers with it).
template <bool TCN = false>
__device__ __forceinline__ void epi_store_chunk(const UmmaParams& p, const float (&v)[32], long long m, long long mo,
                                                int nc0, int cols_left) {
    float o[32];
    // whole chunk inside N and 16-byte aligned constants (parameters living in a flat arena are only 4-byte aligned):
    // vector loads of the per-column constants, no per-element guards
    const bool vec_ok = nc0 + 32 <= p.N &&
                        ((reinterpret_cast<uintptr_t>(p.scale) | reinterpret_cast<uintptr_t>(p.shift)) & 15) == 0;
    if (vec_ok) {
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = v[i];
      if (p.scale) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 sv = __ldg(reinterpret_cast<const float4*>(p.scale + nc0) + g);
          o[4 * g] *= sv.x; o[4 * g + 1] *= sv.y; o[4 * g + 2] *= sv.z; o[4 * g + 3] *= sv.w;
        }
      }
      if (p.shift) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(p.shift + nc0) + g);
          o[4 * g] += bv.x; o[4 * g + 1] += bv.y; o[4 * g + 2] += bv.z; o[4 * g + 3] += bv.w;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int n = nc0 + i;
        float x = v[i];
        if (n < p.N) {
          if (p.scale) x *= __ldg(p.scale + n);
          if (p.shift) x += __ldg(p.shift + n);
        }
        o[i] = x;
      }
    }
    const int ncols = min(32, min(cols_left, p.N - nc0));
    if (TCN && p.pre_act) {
      // ReLU -> inverted dropout on the bf16-rounded activation (what the stand-alone pass m3t_dropout_bf16 sees) ->
      // optional store of this pre-residual tensor
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = fmaxf(o[i], 0.f);
      if (p.drop_thresh) {
        const long long e0 = mo * p.ldc + nc0;
        const unsigned long long seed = p.drop_seed + (p.drop_seed_dev ? __ldg(p.drop_seed_dev) : 0ull);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          o[i] = dropout_u32(seed, e0 + i) >= p.drop_thresh
                     ? __bfloat162float(__float2bfloat16(o[i])) * p.drop_scale : 0.f;
      }
      if (p.out2 && ncols > 0) {
        __nv_bfloat16* op2 = p.out2 + mo * p.ldc + nc0;
        if (ncols == 32 && (p.ldc & 7) == 0) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            reinterpret_cast<uint4*>(op2)[g] =
                make_uint4(pack_bf16x2(o[8 * g], o[8 * g + 1]), pack_bf16x2(o[8 * g + 2], o[8 * g + 3]),
                           pack_bf16x2(o[8 * g + 4], o[8 * g + 5]), pack_bf16x2(o[8 * g + 6], o[8 * g + 7]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < ncols) op2[i] = __float2bfloat16(o[i]);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __bfloat162float(__float2bfloat16(o[i]));   // the sum reads the stored t
      }
    }
    if (p.residual) {
      const __nv_bfloat16* rp = p.residual + (p.res_mapped ? mo : m) * p.ldr + nc0;
      if (ncols == 32 && (p.ldr & 7) == 0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 rv = __ldg(reinterpret_cast<const uint4*>(rp) + g);
          const uint32_t w4[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            o[g * 8 + 2 * j] += bf16lo(w4[j]);
            o[g * 8 + 2 * j + 1] += bf16hi(w4[j]);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < ncols) o[i] += __bfloat162float(rp[i]);
      }
    }
    if (p.relu) {
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = fmaxf(o[i], 0.f);
    }
    if (ncols > 0) {
      if (p.out_f32) {
        float* op = reinterpret_cast<float*>(p.out) + mo * p.ldc + nc0;
        if (ncols == 32 && (p.ldc & 3) == 0) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            reinterpret_cast<float4*>(op)[g] = make_float4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < ncols) op[i] = o[i];
        }
      } else {
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + mo * p.ldc + nc0;
        if (ncols == 32 && (p.ldc & 7) == 0) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            reinterpret_cast<uint4*>(op)[g] =
                make_uint4(pack_bf16x2(o[8 * g], o[8 * g + 1]), pack_bf16x2(o[8 * g + 2], o[8 * g + 3]),
                           pack_bf16x2(o[8 * g + 4], o[8 * g + 5]), pack_bf16x2(o[8 * g + 6], o[8 * g + 7]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < ncols) op[i] = __float2bfloat16(o[i]);
        }
      }
    }
}

// CK = channels per filter-tap chunk of the implicit-GEMM operand: 64 (one 128B swizzle row per pixel) for the
// trunk, 16 (32B rows, SWIZZLE_32B, four taps per pipeline stage) for the space-to-depth stem whose Cin is 16.
template <int BN, int MT, int STAGES, int AKIND, bool A_MN, bool B_MN, int EPI, int CK = 64>
__global__ void __launch_bounds__(kUmmaThreads, 1)
umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const UmmaParams p) {
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N");
  static_assert(MT == 1 || MT == 2, "MT");
  static_assert(!(A_MN && MT != 1), "MN-major A supports MT == 1");
  static_assert(!B_MN || BN % 64 == 0, "MN-major B needs 64-wide atoms");
  static_assert(CK == 64 || CK == 16, "CK");
  static_assert(CK == 64 || AKIND != A_TILED, "CK=16 is an implicit-GEMM mode");
  constexpr int TPS = kBlockK / CK;              // filter taps per pipeline stage (1 or 4)
  constexpr int ATOMS_PER_TILE = MT * 128 / CK;  // A_WGRAD: CK-wide M atoms per (128*MT)-row CTA tile
  constexpr int ATOM_BYTES = kBlockK * CK * 2;   // A_WGRAD: one atom = 64 pixels x CK channels
  constexpr uint32_t A_SWZ = CK == 64 ? SWZ_128B : SWZ_32B;
  constexpr uint32_t A_SBO = 8 * CK * 2;         // 8 rows of CK bf16
  constexpr int A_STAGE = MT * 128 * 128;                 // bytes
  constexpr int B_STAGE = (BN < 8 ? 8 : BN) * 128;        // bytes
  constexpr int STAGE = A_STAGE + B_STAGE;
  constexpr int TM_COLS_RAW = MT * BN;
  constexpr int TM_COLS = TM_COLS_RAW <= 32 ? 32 : TM_COLS_RAW <= 64 ? 64 : TM_COLS_RAW <= 128 ? 128
                          : TM_COLS_RAW <= 256 ? 256 : 512;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem base is only guaranteed 16B aligned: round up to 1024 (swizzle-128B atom)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* red_smem = reinterpret_cast<float*>(tmem_slot + 2);  // [4 warps][2][BN] column partials (stats)

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int n_tile = blockIdx.x % p.tiles_n;
  const int m_tile = blockIdx.x / p.tiles_n;
  const int split = blockIdx.y;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full, 1);
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
      PixCoord pc[MT];
      if constexpr (AKIND == A_IM2COL) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) pc[mt] = pixel_base(p, (m_tile * MT + mt) * 128);
      }
      uint32_t tx_bytes = STAGE;
      if constexpr (AKIND == A_WGRAD) {
        // the second 64-wide atom of the last M tile may not exist
        const int missing = (m_tile + 1) * ATOMS_PER_TILE - p.atoms;
        if (missing > 0) tx_bytes -= missing * ATOM_BYTES;
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < p.k_iters; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
        uint8_t* sA = smem + stage * STAGE;
        uint8_t* sB = sA + A_STAGE;
        mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
        const int kglob = split * p.k_iters + it;  // global K-block index (split-K aware)
        if constexpr (AKIND == A_TILED) {
          if constexpr (!A_MN) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              tma_load_2d(&tmA, &full_bar[stage], sA + mt * 16384, kglob * kBlockK, (m_tile * MT + mt) * 128);
          } else {
            tma_load_2d(&tmA, &full_bar[stage], sA, m_tile * 128, kglob * kBlockK);
            tma_load_2d(&tmA, &full_bar[stage], sA + 8192, m_tile * 128 + 64, kglob * kBlockK);
          }
          if constexpr (!B_MN) {
            tma_load_2d(&tmB, &full_bar[stage], sB, kglob * kBlockK, n_tile * BN);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(&tmB, &full_bar[stage], sB + j * 8192, n_tile * BN + j * 64, kglob * kBlockK);
          }
        } else if constexpr (AKIND == A_IM2COL) {
          if constexpr (CK == 64) {
            int tap = kglob / p.cblocks;
            int cb = kglob - tap * p.cblocks;
            if (p.cb_major) {
              const int taps = p.T * p.R * p.S;
              cb = kglob / taps;
              tap = kglob - cb * taps;
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              im2col_load(&tmA, &full_bar[stage], sA + mt * 16384, p, pc[mt], cb * 64, tap);
            tma_load_2d(&tmB, &full_bar[stage], sB, tap * p.Cin + cb * 64, n_tile * BN);
          } else {
#pragma unroll
            for (int j = 0; j < TPS; ++j) {
              const int tap = kglob * TPS + j;
#pragma unroll
              for (int mt = 0; mt < MT; ++mt)
                im2col_load(&tmA, &full_bar[stage], sA + mt * 16384 + j * (128 * CK * 2), p, pc[mt], 0, tap);
              tma_load_2d(&tmB, &full_bar[stage], sB + j * (BN * CK * 2), tap * CK, n_tile * BN);
            }
          }
        } else {  // A_WGRAD: contraction over pixels, 64 pixels per stage
          const int pix0 = kglob * kBlockK;
          const PixCoord c = pixel_base(p, pix0);
#pragma unroll
          for (int j = 0; j < ATOMS_PER_TILE; ++j) {
            const int atom = m_tile * ATOMS_PER_TILE + j;
            if (atom < p.atoms) {
              const int tap = atom / p.cblocks;
              const int cb = atom - tap * p.cblocks;
              im2col_load(&tmA, &full_bar[stage], sA + j * ATOM_BYTES, p, c, cb * CK, tap);
            }
          }
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_2d(&tmB, &full_bar[stage], sB + j * 8192, n_tile * BN + j * 64, pix0);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    if (lane == 0) {
      constexpr bool a_mn = A_MN || (AKIND == A_WGRAD);
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, a_mn ? 1u : 0u, B_MN ? 1u : 0u);
      // K-major: LBO unused (16), SBO = 8 rows; k-step = 16 elements along the row (CK=64) or the next tap
      // sub-tile (CK=16).  MN-major: LBO = one atom (64 k-rows x CK), SBO = 8 k-rows, k-step = 16 k-rows.
      constexpr uint32_t a_lbo = a_mn ? (uint32_t)ATOM_BYTES : 16u, b_lbo = B_MN ? 8192u : 16u;
      constexpr uint32_t a_kstep = a_mn ? 16u * CK * 2 : (CK == 64 ? 32u : 128u * CK * 2);
      constexpr uint32_t b_kstep = B_MN ? 2048u : (CK == 64 ? 32u : (uint32_t)BN * CK * 2);
      constexpr uint32_t B_SWZ = B_MN ? SWZ_128B : A_SWZ;
      constexpr uint32_t B_SBO = B_MN ? 1024u : A_SBO;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < p.k_iters; ++it) {
        mbar_wait(&full_bar[stage], phase, 200 + stage);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + stage * STAGE);
        const uint32_t sB = sA + A_STAGE;
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          const uint64_t bdesc = make_smem_desc(sB + k * b_kstep, b_lbo, B_SBO, B_SWZ);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t adesc = make_smem_desc(sA + mt * 16384 + k * a_kstep, a_lbo, A_SBO, A_SWZ);
            umma_bf16(tmem_base + mt * BN, adesc, bdesc, idesc, (it | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above have read it
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full);  // accumulator complete
    }
  } else {
    // ======================================= epilogue =======================================
    const int quad = warp & 3;
    const int row = quad * 32 + (int)lane;
    mbar_wait(tmem_full, 0, 300);
    tc_fence_after();
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int n0 = n_tile * BN;

    if constexpr (EPI == EPI_STORE) {
      const bool want_stats = p.stats != nullptr;
      if (want_stats) {
        for (int i = threadIdx.x - 64; i < 4 * 2 * BN; i += 128) red_smem[i] = 0.f;
        named_bar_sync(1, 128);
      }
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const long long m = (long long)(m_tile * MT + mt) * 128 + row;
        const bool row_ok = m < p.M;
        const long long mo = row_ok ? out_row(p, m) : 0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
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
          // ---- output ----
          if (row_ok) epi_store_chunk(p, v, m, mo, n0 + c0, BN - c0);
          // ---- per-column statistics of the raw accumulator (train-mode BatchNorm) ----
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
            if (c0 + (int)lane < BN) {
              red_smem[(quad * 2 + 0) * BN + c0 + lane] += s1[0];
              red_smem[(quad * 2 + 1) * BN + c0 + lane] += s2[0];
            }
          }
        }
      }
      if (want_stats) {
        named_bar_sync(1, 128);
        for (int i = threadIdx.x - 64; i < 2 * BN; i += 128) {
          const int which = i / BN, c = i - which * BN;
          if (n0 + c < p.N) {
            const float s = red_smem[(0 * 2 + which) * BN + c] + red_smem[(1 * 2 + which) * BN + c] +
                            red_smem[(2 * 2 + which) * BN + c] + red_smem[(3 * 2 + which) * BN + c];
            atomicAdd(p.stats + which * p.N + n0 + c, s);
          }
        }
      }
    } else {  // EPI_ATOMIC_T : out_f32[n][m] += acc  (rows of D are contiguous in the output)
      float* outp = reinterpret_cast<float*>(p.out) + (p.det_stride ? (1 + split) * p.det_stride : 0);
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const long long m = (long long)(m_tile * MT + mt) * 128 + row;
        const bool row_ok = m < p.M;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(t_lane + mt * BN + c0, r);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int n = n0 + c0 + i;
              if (n < p.N) atomicAdd(outp + (long long)n * p.ldc + m, __uint_as_float(r[i]));
            }
          }
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

template <int BN, int MT, int STAGES>
constexpr int umma_smem_bytes() {
  return STAGES * (MT * 128 * 128 + (BN < 8 ? 8 : BN) * 128) + (2 * STAGES + 1) * 8 + 16 + 4 * 2 * BN * 4 + 1024;
}

}  // namespace m3t
