// Hardware-semantics probes (not on the product path).  m3t_debug_rowshift answers one question for the round-2
// halo-tile convolution: can a K-major SWIZZLE_128B operand start at an arbitrary 128-byte row of a TMA-written tile
// (start address not 1024-byte aligned), and does the descriptor's base_offset field have to carry (addr >> 7) & 7?
#include "../../include/m3t_b200.h"
#include "common.cuh"
#include "tmap.cuh"
#include "ptx.cuh"

namespace m3t {

__global__ void __launch_bounds__(128, 1)
dbg_rowshift_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    float* __restrict__ out, int shift_rows, int mode) {
  m3t::pdl_wait();
  m3t::pdl_launch();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 256 rows x 128 B
  uint8_t* sB = smem + 256 * 128;     // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 64 * 128);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 256 * 128 + 64 * 128);
    tma_load_2d(&tmA, bar, sA, 0, 0);
    tma_load_2d(&tmB, bar, sB, 0, 0);
    mbar_wait(bar, 0, 900);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sA) + (uint32_t)shift_rows * 128u;
    const uint32_t b0 = smem_u32(sB);
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    for (int k = 0; k < 4; ++k) {
      uint64_t ad = make_smem_desc(a0 + k * 32, 16, 1024, SWZ_128B);
      if (mode == 1) ad |= static_cast<uint64_t>((a0 >> 7) & 7) << 49;
      const uint64_t bd = make_smem_desc(b0 + k * 32, 16, 1024, SWZ_128B);
      umma_bf16(tmem, ad, bd, idesc, k != 0);
    }
    umma_commit(done);
  }
  mbar_wait(done, 0, 901);
  tc_fence_after();
  const int row = (warp & 3) * 32 + (int)lane;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[row * 64 + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

}  // namespace m3t

using namespace m3t;

extern "C" int m3t_debug_rowshift(const void* A /*[256][64] bf16*/, const void* B /*[64][64] bf16*/,
                                  float* out /*[128][64]*/, int shift_rows, int mode, void* stream) {
  if (shift_rows < 0 || shift_rows > 128) return -1;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d_bf16(&tmA, A, 64, 256, 64, 64, 256);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmB, B, 64, 64, 64, 64, 64);
  if (rc) return rc;
  const int smem = 256 * 128 + 64 * 128 + 64 + 1024;
  if (cudaFuncSetAttribute(dbg_rowshift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return -20;
  m3t::launch_k(dbg_rowshift_kernel, dim3(1), dim3(128), smem, reinterpret_cast<cudaStream_t>(stream), tmA, tmB, out, shift_rows, mode);
  count_launch();
  return launch_status();
}
