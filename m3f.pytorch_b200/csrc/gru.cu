// Bidirectional GRU recurrence as persistent kernels (forward and BPTT).
//
// The input projection  gi = x . W_ih^T + b_ih  of all time steps is one tcgen05 GEMM (m3t_gemm_bf16); what is
// left is the serial part.  One launch runs a whole layer, both directions:
//   grid = (H/32 hidden slices, batch-slice lanes, 2 directions), 256 threads, one CTA per SM (cooperative launch).
//   A CTA keeps its slice of W_hh (3 gates x 32 units x H, bf16) in shared memory for the whole sequence, and per
//   step multiplies the previous hidden state of its batch slice (re-read from the layer output in L2, bf16) with
//   it on the legacy tensor path (mma.sync m16n8k16, fp32 accumulate) -- the step is latency/sync bound, not
//   throughput bound, so the 128-row tcgen05 tile would only add TMEM round trips.  The fp32 hidden state of the
//   (batch, unit) elements a thread owns never leaves its registers.  CTAs that share a (direction, batch slice)
//   synchronise once per step through a global arrival counter.
// Gate order and arithmetic follow nn.GRU (reference: models/rnn.py:17,72-75):
//   r = s(gi_r + W_hr h + b_hr), z = s(gi_z + W_hz h + b_hz), n = tanh(gi_n + r*(W_hn h + b_hn)), h' = (1-z)n + z h.
#include "../../include/m3t_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace m3t {

constexpr int kGruThreads = 256;
constexpr int kJS = 32;          // hidden units per CTA
constexpr int kFwdBS = 64;       // batch rows per slice (forward)
constexpr int kBwdBS = 32;       // batch rows per slice (backward)
constexpr int kMaxSlicesPerCta = 4;

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(saddr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 ld_cg_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
// 16-byte async copy global(L2) -> shared; src_bytes = 0 zero-fills the destination (rows past the batch)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ float ld_cg_bf16(const __nv_bfloat16* p) {
  unsigned short v;
  asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return __uint_as_float(((uint32_t)v) << 16);
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// all threads of the CTA call; thread 0 arrives / polls
// Arrival = CTA barrier (orders every thread's global writes before thread 0) + ONE gpu-scope release reduction:
// release is cumulative over what thread 0 observed through bar.sync, so no per-thread __threadfence is needed.
__device__ __forceinline__ void group_arrive(unsigned* counter) {
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}
__device__ __forceinline__ void group_wait(const unsigned* counter, unsigned target) {
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while (ld_acquire(counter) < target) {
      if (clock64() - t0 > 4000000000LL) {
        printf("m3t: gru group barrier timeout block=(%d,%d,%d) target=%u have=%u\n", blockIdx.x, blockIdx.y,
               blockIdx.z, target, ld_acquire(counter));
        __trap();
      }
    }
  }
  __syncthreads();   // the acquire load of thread 0 + this barrier order the other threads' reads after the arrivals
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

struct GruParams {
  int B, T, H;
  const float* gi;              // [B*T][2][3H]   x-projection incl. b_ih
  const __nv_bfloat16* w;       // fwd: [2][3H][H] (W_hh)      bwd: [2][H][3H] (W_hh^T)
  const float* b_hh;            // [2][3H]
  __nv_bfloat16* out;           // [B][T][2H]     hidden states (bf16)
  float* out_f32;               // optional fp32 copy of out (API boundary) or null
  float* saved;                 // [B*T][2][4][H] r, z, n, hn  (fwd writes when non-null; bwd reads)
  // backward only
  const __nv_bfloat16* dout;    // [B][T][2H] gradient wrt out
  __nv_bfloat16* dgi;           // [B*T][2][3H]
  __nv_bfloat16* dgh;           // [B*T][2][3H]
  __nv_bfloat16* hprev;         // [B*T][2][H]   h_{t-1} copy (zero at sequence start) for the dW_hh GEMM
  unsigned* counters;           // [2][nslices_b] zero-initialised
  int nbslices;                 // number of batch slices
  float* dbias;                 // bwd, optional: [2 (ih, hh)][2 dirs][3H] fp32 += column sums of dgi / dgh (bias grads)
};

// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGruThreads, 1) gru_fwd_kernel(const GruParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int H = p.H, T = p.T, B = p.B;
  const int ldk = H + 8;  // padded row (elements): rows shift by 16 B -> conflict-free ldmatrix
  __nv_bfloat16* Wsm = reinterpret_cast<__nv_bfloat16*>(smem);   // [96][ldk]
  __nv_bfloat16* hsm = Wsm + 96 * ldk;                           // [64][ldk]
  const int js = blockIdx.x, dir = blockIdx.z;
  const int nsl = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int vec_per_row = H / 8;

  // resident W_hh slice: smem row g*32 + j  <-  W_hh[dir][g*H + js*32 + j][:]
  for (int i = tid; i < 96 * vec_per_row; i += kGruThreads) {
    const int row = i / vec_per_row, v = i - row * vec_per_row;
    const int g = row >> 5, j = row & 31;
    const uint4 val =
        __ldg(reinterpret_cast<const uint4*>(p.w + ((long long)dir * 3 * H + g * H + js * kJS + j) * H) + v);
    *reinterpret_cast<uint4*>(Wsm + row * ldk + v * 8) = val;
  }
  const int mrow = (warp & 3) * 16, jhalf = warp >> 2;
  // per-thread owned elements: (half, e): row = mrow + lane/4 + (e>>1)*8 ; j = 16*jhalf + 8*half + (lane%4)*2 + (e&1)
  float bhh[3][4];  // [gate][half*2 + (e&1)]
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
      for (int e = 0; e < 2; ++e)
        bhh[g][hh * 2 + e] =
            p.b_hh[dir * 3 * H + g * H + js * kJS + 16 * jhalf + 8 * hh + (lane & 3) * 2 + e];
  float hreg[kMaxSlicesPerCta][8];
#pragma unroll
  for (int s = 0; s < kMaxSlicesPerCta; ++s)
#pragma unroll
    for (int e = 0; e < 8; ++e) hreg[s][e] = 0.f;
  __syncthreads();

  for (int step = 0; step < T; ++step) {
    const int t = dir == 0 ? step : T - 1 - step;
    const int tprev = dir == 0 ? t - 1 : t + 1;
#pragma unroll
    for (int si = 0; si < kMaxSlicesPerCta; ++si) {
      const int bs = blockIdx.y + si * gridDim.y;
      if (bs >= p.nbslices) break;
      const int b0 = bs * kFwdBS;
      unsigned* counter = p.counters + dir * p.nbslices + bs;
      float acc[6][4];
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
      // the x-projection of this step does not depend on h: fetch it before waiting on the other slices
      // (element pairs (e&1) are adjacent hidden units -> 8-byte loads)
      float2 gir[4], giz[4], gin[4];
#pragma unroll
      for (int pr = 0; pr < 4; ++pr) {
        const int hh = pr >> 1, rs = pr & 1;
        const int b = b0 + mrow + (lane >> 2) + rs * 8;
        const int j = js * kJS + 16 * jhalf + 8 * hh + (lane & 3) * 2;
        gir[pr] = giz[pr] = gin[pr] = make_float2(0.f, 0.f);
        if (b < B) {
          const float* gi = p.gi + ((long long)b * T + t) * 6 * H + dir * 3 * H + j;
          gir[pr] = __ldg(reinterpret_cast<const float2*>(gi));
          giz[pr] = __ldg(reinterpret_cast<const float2*>(gi + H));
          gin[pr] = __ldg(reinterpret_cast<const float2*>(gi + 2 * H));
        }
      }
      if (step > 0) {
        group_wait(counter, (unsigned)(step * nsl));
        // stage h_{t-1} of this batch slice (bf16, written by the nsl CTAs of the group) -> smem
        for (int i = tid; i < kFwdBS * vec_per_row; i += kGruThreads) {
          const int row = i / vec_per_row, v = i - row * vec_per_row;
          const int b = b0 + row;
          const bool ok = b < B;
          cp_async16(hsm + row * ldk + v * 8,
                     p.out + ((long long)(ok ? b : 0) * T + tprev) * 2 * H + dir * H + v * 8, ok ? 16 : 0);
        }
        cp_async_wait_all();
        __syncthreads();
        const uint32_t a_base = smem_u32(hsm + (mrow + (lane & 7) + ((lane >> 3) & 1) * 8) * ldk + (lane >> 4) * 8);
        const uint32_t b_base =
            smem_u32(Wsm + (16 * jhalf + (lane & 7) + (lane >> 4) * 8) * ldk + ((lane >> 3) & 1) * 8);
        for (int kk = 0; kk < H / 16; ++kk) {
          uint32_t a[4];
          ldmatrix_x4(a, a_base + kk * 32);
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            uint32_t bb[4];
            ldmatrix_x4(bb, b_base + (g * 32 * ldk) * 2 + kk * 32);
            mma_16816(acc[g * 2 + 0], a, bb[0], bb[1]);
            mma_16816(acc[g * 2 + 1], a, bb[2], bb[3]);
          }
        }
      }
      // gate math on the 8 owned (b, j) elements, as 4 pairs of adjacent hidden units
#pragma unroll
      for (int pr = 0; pr < 4; ++pr) {
        const int hh = pr >> 1, rs = pr & 1;
        const int b = b0 + mrow + (lane >> 2) + rs * 8;
        const int j = js * kJS + 16 * jhalf + 8 * hh + (lane & 3) * 2;
        if (b < B) {
          const long long row = (long long)b * T + t;
          float hnew[2], rr[2], zz[2], nn[2], hnn[2];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int e = rs * 2 + q;
            const float hp = hreg[si][hh * 4 + e];
            const float hr = acc[0 + hh][e] + bhh[0][hh * 2 + q];
            const float hz = acc[2 + hh][e] + bhh[1][hh * 2 + q];
            const float hn = acc[4 + hh][e] + bhh[2][hh * 2 + q];
            const float r = sigmoidf_((q ? gir[pr].y : gir[pr].x) + hr);
            const float z = sigmoidf_((q ? giz[pr].y : giz[pr].x) + hz);
            const float n = tanhf((q ? gin[pr].y : gin[pr].x) + r * hn);
            hnew[q] = (1.f - z) * n + z * hp;
            hreg[si][hh * 4 + e] = hnew[q];
            rr[q] = r; zz[q] = z; nn[q] = n; hnn[q] = hn;
          }
          *reinterpret_cast<uint32_t*>(p.out + row * 2 * H + dir * H + j) = pack_bf16x2(hnew[0], hnew[1]);
          if (p.out_f32) *reinterpret_cast<float2*>(p.out_f32 + row * 2 * H + dir * H + j) = make_float2(hnew[0], hnew[1]);
          if (p.saved) {
            float* sv = p.saved + (row * 2 + dir) * 4 * H + j;
            *reinterpret_cast<float2*>(sv) = make_float2(rr[0], rr[1]);
            *reinterpret_cast<float2*>(sv + H) = make_float2(zz[0], zz[1]);
            *reinterpret_cast<float2*>(sv + 2 * H) = make_float2(nn[0], nn[1]);
            *reinterpret_cast<float2*>(sv + 3 * H) = make_float2(hnn[0], hnn[1]);
          }
        }
      }
      group_arrive(counter);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Small-batch inference variant (B <= 16: long-sequence evaluation at 16 clips per GPU, BASELINE config 5; config 1).
// There a step of gru_fwd_kernel is pure latency: an L2 round trip to publish h, a gpu-scope arrival counter to poll,
// another L2 round trip to fetch h (5.7 us per step measured).  Here the CTAs of one direction form ONE thread-block
// cluster and the hidden state never leaves the SMs:
//   grid (H/64, ceil(B/16), 2), cluster (H/64, 1, 1): a cluster serves 16 batch rows of one direction (rows are
//   independent, so a larger batch is simply more clusters); CTA r owns hidden units [64r, 64r+64); its W_hh slice
//   (3 gates x 64 units x H, bf16, 196 KB at H = 512) stays in shared memory for the whole sequence;
//   per step: h' slice (16 rows x 64 units, bf16) -> own shared memory (double-buffered) -> ONE hardware cluster
//   barrier (release/acquire) -> every CTA pulls the H/64 slices over DSMEM (ld.shared::cluster, 16-byte vectors)
//   into its h tile -> mma.sync m16n8k16 over K = H (one m16 tile = all batch rows; warp w owns units 8w..8w+7 of
//   the three gates) -> gates on the 4 (b, j) elements a thread owns (fp32 state in registers).
// Same operand values, same k order and same gate arithmetic as gru_fwd_kernel.
// ------------------------------------------------------------------------------------------------------------
constexpr int kCJS = 64;         // hidden units per CTA (cluster variant)
constexpr int kCRows = 16;       // batch rows (one m16 tile)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// The two halves of the cluster barrier.  Global stores issued BETWEEN them are not covered by this release (they are
// by the next one, a whole step later, when they have long completed): a release with this step's global stores still
// in flight showed up as the `membar` stall of both cluster kernels (15 % of their stall cycles).
__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint4 ld_dsmem_u4(uint32_t local_saddr, uint32_t rank) {
  uint32_t ra;
  uint4 v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_saddr), "r"(rank));
  asm volatile("ld.shared::cluster.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(ra)
               : "memory");
  return v;
}

__global__ void __launch_bounds__(kGruThreads, 1) gru_fwd_cluster_kernel(const GruParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int H = p.H, T = p.T, B = p.B;
  const int ldk = H + 8;
  __nv_bfloat16* Wsm = reinterpret_cast<__nv_bfloat16*>(smem);   // [192][ldk]  row g*64 + j
  __nv_bfloat16* hsm = Wsm + 3 * kCJS * ldk;                     // [16][ldk]   h_{t-1}, all H units
  __nv_bfloat16* sl = hsm + kCRows * ldk;                        // [2][16][64] own h' slice, double-buffered
  const int js = (int)cluster_ctarank(), dir = blockIdx.z;
  const int ncta = gridDim.x;                                    // = H / 64 = cluster size
  const int b0 = blockIdx.y * kCRows;                            // batch rows of this cluster (rows are independent:
                                                                 // larger batches are more clusters, each with W_hh)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int vec_per_row = H / 8;

  for (int i = tid; i < 3 * kCJS * vec_per_row; i += kGruThreads) {
    const int row = i / vec_per_row, v = i - row * vec_per_row;
    const int g = row / kCJS, j = row - g * kCJS;
    const uint4 val =
        __ldg(reinterpret_cast<const uint4*>(p.w + ((long long)dir * 3 * H + g * H + js * kCJS + j) * H) + v);
    *reinterpret_cast<uint4*>(Wsm + row * ldk + v * 8) = val;
  }
  // owned elements: rows b = lane/4 + 8*rs (rs = 0, 1), hidden units j0 + q (q = 0, 1) of this CTA's 64
  const int j0 = 8 * warp + (lane & 3) * 2;
  float bhh[3][2];
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int q = 0; q < 2; ++q) bhh[g][q] = p.b_hh[dir * 3 * H + g * H + js * kCJS + j0 + q];
  float hreg[4] = {0.f, 0.f, 0.f, 0.f};     // [rs*2 + q]
  __syncthreads();

  const uint32_t a_base = smem_u32(hsm + ((lane & 7) + ((lane >> 3) & 1) * 8) * ldk + (lane >> 4) * 8);
  // gates r and z in one ldmatrix.x4 (lanes 0-15: r rows, k lo / k hi; lanes 16-31: z rows), gate n in an x2
  const uint32_t brz_base =
      smem_u32(Wsm + ((lane >> 4) * kCJS + 8 * warp + (lane & 7)) * ldk + ((lane >> 3) & 1) * 8);
  const uint32_t bn_base = smem_u32(Wsm + (2 * kCJS + 8 * warp + (lane & 7)) * ldk + ((lane >> 3) & 1) * 8);

  for (int step = 0; step < T; ++step) {
    const int t = dir == 0 ? step : T - 1 - step;
    float acc[3][4];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[g][e] = 0.f;
    // this step's input projection does not depend on h: issue the loads before touching the peers
    float2 gir[2], giz[2], gin[2];
#pragma unroll
    for (int rs = 0; rs < 2; ++rs) {
      const int b = b0 + (lane >> 2) + rs * 8;
      gir[rs] = giz[rs] = gin[rs] = make_float2(0.f, 0.f);
      if (b < B) {
        const float* gi = p.gi + ((long long)b * T + t) * 6 * H + dir * 3 * H + js * kCJS + j0;
        gir[rs] = __ldg(reinterpret_cast<const float2*>(gi));
        giz[rs] = __ldg(reinterpret_cast<const float2*>(gi + H));
        gin[rs] = __ldg(reinterpret_cast<const float2*>(gi + 2 * H));
      }
    }
    if (step > 0) {
      // pull h_{t-1}: slice r of buffer (step-1)&1 from CTA r, for every r of the cluster (own slice included)
      const uint32_t sl_local = smem_u32(sl + ((step - 1) & 1) * kCRows * kCJS);
      const int nvec = kCRows * ncta * 8;                 // 16-byte vectors: 8 per (row, slice)
      for (int v0 = tid; v0 < nvec; v0 += 4 * kGruThreads) {
        uint4 val[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int v = v0 + u * kGruThreads;
          if (v < nvec) {
            const int peer = v >> 7, rem = v & 127;       // 128 vectors per slice
            val[u] = ld_dsmem_u4(sl_local + (uint32_t)((rem >> 3) * kCJS + (rem & 7) * 8) * 2u, (uint32_t)peer);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int v = v0 + u * kGruThreads;
          if (v < nvec) {
            const int peer = v >> 7, rem = v & 127;
            *reinterpret_cast<uint4*>(hsm + (rem >> 3) * ldk + peer * kCJS + (rem & 7) * 8) = val[u];
          }
        }
      }
      __syncthreads();
      for (int kk = 0; kk < H / 16; ++kk) {
        uint32_t a[4], brz[4], bn[2];
        ldmatrix_x4(a, a_base + kk * 32);
        ldmatrix_x4(brz, brz_base + kk * 32);
        ldmatrix_x2(bn, bn_base + kk * 32);
        mma_16816(acc[0], a, brz[0], brz[1]);
        mma_16816(acc[1], a, brz[2], brz[3]);
        mma_16816(acc[2], a, bn[0], bn[1]);
      }
    }
    __nv_bfloat16* slw = sl + (step & 1) * kCRows * kCJS;
    float hnew[2][2], rr[2][2], zz[2][2], nn[2][2], hnn[2][2];
    uint32_t packed[2] = {0u, 0u};
#pragma unroll
    for (int rs = 0; rs < 2; ++rs) {
      const int lrow = (lane >> 2) + rs * 8;
      if (b0 + lrow < B) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int e = rs * 2 + q;
          const float hp = hreg[e];
          const float hr = acc[0][e] + bhh[0][q];
          const float hz = acc[1][e] + bhh[1][q];
          const float hn = acc[2][e] + bhh[2][q];
          const float r = sigmoidf_((q ? gir[rs].y : gir[rs].x) + hr);
          const float z = sigmoidf_((q ? giz[rs].y : giz[rs].x) + hz);
          const float n = tanhf((q ? gin[rs].y : gin[rs].x) + r * hn);
          hnew[rs][q] = (1.f - z) * n + z * hp;
          hreg[e] = hnew[rs][q];
          rr[rs][q] = r; zz[rs][q] = z; nn[rs][q] = n; hnn[rs][q] = hn;
        }
        packed[rs] = pack_bf16x2(hnew[rs][0], hnew[rs][1]);
      }
      *reinterpret_cast<uint32_t*>(slw + lrow * kCJS + j0) = packed[rs];  // rows past the batch stay zero
    }
    // One barrier per step: buffer step&1 is complete in every CTA after it, and nobody still reads buffer
    // (step+1)&1 (those pulls happened before the pullers' MMAs of this step).  The last one also keeps every
    // CTA's shared memory alive until its peers are done with it.  The step's global stores go between the two
    // halves: the peers only need the shared-memory slice.
    cluster_arrive_release();
#pragma unroll
    for (int rs = 0; rs < 2; ++rs) {
      const int b = b0 + (lane >> 2) + rs * 8;
      if (b < B) {
        const long long row = (long long)b * T + t;
        if (p.saved) {      // training: gates for BPTT, same layout as gru_fwd_kernel
          float* sv = p.saved + (row * 2 + dir) * 4 * H + js * kCJS + j0;
          *reinterpret_cast<float2*>(sv) = make_float2(rr[rs][0], rr[rs][1]);
          *reinterpret_cast<float2*>(sv + H) = make_float2(zz[rs][0], zz[rs][1]);
          *reinterpret_cast<float2*>(sv + 2 * H) = make_float2(nn[rs][0], nn[rs][1]);
          *reinterpret_cast<float2*>(sv + 3 * H) = make_float2(hnn[rs][0], hnn[rs][1]);
        }
        *reinterpret_cast<uint32_t*>(p.out + row * 2 * H + dir * H + js * kCJS + j0) = packed[rs];
        if (p.out_f32)
          *reinterpret_cast<float2*>(p.out_f32 + row * 2 * H + dir * H + js * kCJS + j0) =
              make_float2(hnew[rs][0], hnew[rs][1]);
      }
    }
    cluster_wait_acquire();
  }
}

// ------------------------------------------------------------------------------------------------------------
// BPTT.  Per step and batch slice (32 rows):
//   phase A (elementwise, owned (b, j)):  dh = dout_t + dh_rec ; gate gradients ; write dgi, dgh (bf16), hprev copy
//   group barrier
//   phase B: dh_rec[b][k in own slice] = dh*z + sum_{g,j} dgh[b][g][j] * W_hh[g][j][k]   (mma.sync over 3H)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGruThreads, 1) gru_bwd_kernel(const GruParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int H = p.H, T = p.T, B = p.B;
  const int K3 = 3 * H;
  const int ldk = K3 + 8;
  __nv_bfloat16* Wsm = reinterpret_cast<__nv_bfloat16*>(smem);  // [32][ldk] rows = own k, cols = (g, j)
  __nv_bfloat16* gsm = Wsm + 32 * ldk;                          // [32][ldk] dgh tile of the batch slice
  const int js = blockIdx.x, dir = blockIdx.z;
  const int nsl = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int vec_per_row = K3 / 8;
  for (int i = tid; i < 32 * vec_per_row; i += kGruThreads) {
    const int row = i / vec_per_row, v = i - row * vec_per_row;
    const uint4 val = __ldg(reinterpret_cast<const uint4*>(p.w + ((long long)dir * H + js * kJS + row) * K3) + v);
    *reinterpret_cast<uint4*>(Wsm + row * ldk + v * 8) = val;
  }
  const int mrow = (warp & 1) * 16, ncol = (warp >> 1) * 8;
  float dhrec[kMaxSlicesPerCta][4];
#pragma unroll
  for (int s = 0; s < kMaxSlicesPerCta; ++s)
#pragma unroll
    for (int e = 0; e < 4; ++e) dhrec[s][e] = 0.f;
  __syncthreads();

  // bias gradients = column sums of the gate gradients over (b, t): per-thread partials for the two hidden units this
  // thread owns (e & 1), reduced over the warp's rows and added atomically once, after the last step
  float sb_r[2] = {0.f, 0.f}, sb_z[2] = {0.f, 0.f}, sb_n[2] = {0.f, 0.f}, sb_nr[2] = {0.f, 0.f};

  for (int step = 0; step < T; ++step) {
    const int t = dir == 0 ? T - 1 - step : step;       // reverse of the forward order
    const int tprev = dir == 0 ? t - 1 : t + 1;         // time index of h_{prev} in forward order
    const bool has_prev = tprev >= 0 && tprev < T;
    float dh_direct[kMaxSlicesPerCta][4];
    // ---- phase A for every batch slice of this CTA, each followed by its group arrival: the barrier latency of
    //      slice i overlaps the element-wise work of slice i+1 ----
#pragma unroll
    for (int si = 0; si < kMaxSlicesPerCta; ++si) {
      const int bs = blockIdx.y + si * gridDim.y;
      if (bs >= p.nbslices) break;
      const int b0 = bs * kBwdBS;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int b = b0 + mrow + (lane >> 2) + (e >> 1) * 8;
        const int j = js * kJS + ncol + (lane & 3) * 2 + (e & 1);
        dh_direct[si][e] = 0.f;
        if (b < B) {
          const long long row = (long long)b * T + t;
          const float dh = __bfloat162float(p.dout[row * 2 * H + dir * H + j]) + dhrec[si][e];
          const float* sv = p.saved + (row * 2 + dir) * 4 * H + j;
          const float r = __ldg(sv), z = __ldg(sv + H), n = __ldg(sv + 2 * H), hn = __ldg(sv + 3 * H);
          const float hp =
              has_prev ? __bfloat162float(p.out[((long long)b * T + tprev) * 2 * H + dir * H + j]) : 0.f;
          const float dn = dh * (1.f - z);
          const float dz = dh * (hp - n);
          const float dan = dn * (1.f - n * n);
          const float daz = dz * z * (1.f - z);
          const float dar = dan * hn * r * (1.f - r);
          const long long g0 = row * 6 * H + dir * 3 * H + j;
          sb_r[e & 1] += dar;
          sb_z[e & 1] += daz;
          sb_n[e & 1] += dan;
          sb_nr[e & 1] += dan * r;
          p.dgi[g0] = __float2bfloat16(dar);
          p.dgi[g0 + H] = __float2bfloat16(daz);
          p.dgi[g0 + 2 * H] = __float2bfloat16(dan);
          p.dgh[g0] = __float2bfloat16(dar);
          p.dgh[g0 + H] = __float2bfloat16(daz);
          p.dgh[g0 + 2 * H] = __float2bfloat16(dan * r);
          p.hprev[(row * 2 + dir) * H + j] = __float2bfloat16(hp);
          dh_direct[si][e] = dh * z;
        }
      }
      if (step + 1 < T) group_arrive(p.counters + dir * p.nbslices + bs);
    }
    if (step + 1 == T) break;  // gradient wrt h_0 is not needed
    // ---- phase B ----
#pragma unroll
    for (int si = 0; si < kMaxSlicesPerCta; ++si) {
      const int bs = blockIdx.y + si * gridDim.y;
      if (bs >= p.nbslices) break;
      const int b0 = bs * kBwdBS;
      group_wait(p.counters + dir * p.nbslices + bs, (unsigned)((step + 1) * nsl));
      for (int i = tid; i < kBwdBS * vec_per_row; i += kGruThreads) {
        const int row = i / vec_per_row, v = i - row * vec_per_row;
        const int b = b0 + row;
        const bool ok = b < B;
        cp_async16(gsm + row * ldk + v * 8, p.dgh + ((long long)(ok ? b : 0) * T + t) * 6 * H + dir * 3 * H + v * 8,
                   ok ? 16 : 0);
      }
      cp_async_wait_all();
      __syncthreads();
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t a_base = smem_u32(gsm + (mrow + (lane & 7) + ((lane >> 3) & 1) * 8) * ldk + (lane >> 4) * 8);
      const uint32_t b_base = smem_u32(Wsm + (ncol + (lane & 7)) * ldk + ((lane >> 3) & 1) * 8);
      for (int kk = 0; kk < K3 / 16; ++kk) {
        uint32_t a[4], bb[2];
        ldmatrix_x4(a, a_base + kk * 32);
        ldmatrix_x2(bb, b_base + kk * 32);
        mma_16816(acc, a, bb[0], bb[1]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) dhrec[si][e] = dh_direct[si][e] + acc[e];
      __syncthreads();  // gsm is rewritten by the next slice / step
    }
  }
  if (p.dbias) {
    // lanes with equal (lane & 3) own the same two hidden units on different rows: butterfly over lane bits 2..4
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        sb_r[jj] += __shfl_xor_sync(0xffffffffu, sb_r[jj], o);
        sb_z[jj] += __shfl_xor_sync(0xffffffffu, sb_z[jj], o);
        sb_n[jj] += __shfl_xor_sync(0xffffffffu, sb_n[jj], o);
        sb_nr[jj] += __shfl_xor_sync(0xffffffffu, sb_nr[jj], o);
      }
    }
    if (lane < 4) {
      float* bih = p.dbias + (long long)dir * 3 * H;            // [ih][dir][3H]
      float* bhh = p.dbias + (long long)(2 + dir) * 3 * H;      // [hh][dir][3H]
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = js * kJS + ncol + lane * 2 + jj;
        atomicAdd(bih + j, sb_r[jj]);
        atomicAdd(bih + H + j, sb_z[jj]);
        atomicAdd(bih + 2 * H + j, sb_n[jj]);
        atomicAdd(bhh + j, sb_r[jj]);
        atomicAdd(bhh + H + j, sb_z[jj]);
        atomicAdd(bhh + 2 * H + j, sb_nr[jj]);
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------
// BPTT on a thread-block cluster (H = 32 * CS, CS = 4, 8 or 16 CTAs).  The kernel above gathers the whole (rows, 3H)
// gate-gradient tile through L2 every step and multiplies it with a 32-column slice of W_hh^T: a 96 KB read and a
// serial chain of 3H/16 MMAs behind a global-memory barrier (9.2 us per step at any batch).  Here the contraction
// is split the other way: CTA js owns hidden units [32 js, 32 js + 32) both as the ELEMENT-WISE owner (its threads
// hold dh for them in registers) and as a K-SLICE of the recurrent product, so the gate gradients it needs as MMA
// input are the ones it has just computed (a 16 x 96 tile in its own shared memory, never re-read from L2):
//     partial[b][k] = sum_{g, j in own 32} dgh[b][g][j] * W_hh[g][j][k]        (all H columns k, K = 96)
// and the partials are reduce-scattered over distributed shared memory: after ONE cluster barrier every CTA reads
// the 32 columns it owns from each peer's partial tile (CS x 2 KB) and adds them in rank order — a fixed order, so
// the kernel is deterministic as it stands.  W_hh^T[:, own (g, j)] stays in shared memory (H x 96 bf16 = 98 KB) for
// the whole sequence.  Batch rows are independent: a cluster owns up to 6 16-row slices (blockIdx.y, + gridDim.y,
// ...), more rows are more clusters; with two or more slices the items (slice, step) are software-pipelined: an
// item's barrier wait and reduce-scatter run after the NEXT item's phase A + MMA (three partial buffers).
// ------------------------------------------------------------------------------------------------------------
constexpr int kClMaxSlices = 6;       // 16-row slices per cluster (B = 256 at H = 512, where 7 clusters of 16 CTAs are
                                      // resident: 3 clusters per direction x 6 slices; 2.4 us per slice-step pipelined)
constexpr int kBRows = 16;             // batch rows per slice (one m16 tile)
constexpr int kBLd = 3 * kJS + 8;      // padded K row (bf16 elements): 208 B, conflict-free ldmatrix

__device__ __forceinline__ float2 ld_dsmem_f2(uint32_t local_saddr, uint32_t rank) {
  uint32_t ra;
  float2 v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_saddr), "r"(rank));
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(ra) : "memory");
  return v;
}

__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_saddr, uint32_t rank) {
  uint32_t ra;
  float4 v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_saddr), "r"(rank));
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(ra)
               : "memory");
  return v;
}

struct BwdIn {            // what phase A of one (slice, step) reads from global memory, for one (row, unit pair)
  float2 r, z, n, hn;
  uint32_t dout2, hp2;
  bool ok;
};

template <int CS>
__global__ void __launch_bounds__(kGruThreads, 1) gru_bwd_cluster_kernel(const GruParams p) {
  pdl_wait();
  pdl_launch();
  constexpr int H = CS * kJS;
  constexpr int K3 = 3 * H;
  constexpr int PLD = H + 8;            // fp32 row of a partial tile
  constexpr int NTW = CS / 2;           // 8-column n-tiles per warp (H / 8 tiles over 8 warps)
  extern __shared__ __align__(16) uint8_t smem[];
  __nv_bfloat16* Wsm = reinterpret_cast<__nv_bfloat16*>(smem);     // [H][kBLd]  row k, col g*32 + jl
  __nv_bfloat16* Asm = Wsm + H * kBLd;                             // [2][16][kBLd] own gate gradients of an item
  float* Psm = reinterpret_cast<float*>(Asm + 2 * kBRows * kBLd);  // [3][16][PLD] partial dh_rec, all H columns
  float2* Dsm = reinterpret_cast<float2*>(Psm + 3 * kBRows * PLD); // [kClMaxSlices][256] recurrent gradient slots
  const int T = p.T, B = p.B;
  const int js = (int)cluster_ctarank(), dir = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < H * 12; i += kGruThreads) {
    const int k = i / 12, rem = i - k * 12, g = rem >> 2, v = rem & 3;
    const uint4 val =
        __ldg(reinterpret_cast<const uint4*>(p.w + ((long long)dir * H + k) * K3 + g * H + js * kJS) + v);
    *reinterpret_cast<uint4*>(Wsm + k * kBLd + g * kJS + v * 8) = val;
  }
  // element-wise ownership: row = tid / 16 of the slice, hidden units jl, jl + 1 of this CTA's 32
  const int row = tid >> 4, jl = (tid & 15) * 2;
  const int j = js * kJS + jl;
  int nsl = 0;                                             // slices of this cluster
  for (int bs = blockIdx.y; bs < p.nbslices; bs += gridDim.y) ++nsl;

  auto load_in = [&](int step, int si) {
    BwdIn in;
    const int t = dir == 0 ? T - 1 - step : step;
    const int tprev = dir == 0 ? t - 1 : t + 1;
    const int b = (blockIdx.y + si * gridDim.y) * kBRows + row;
    in.ok = b < B;
    in.hp2 = 0u;
    if (in.ok) {
      const long long rw = (long long)b * T + t;
      const float* sv = p.saved + (rw * 2 + dir) * 4 * H + j;
      in.r = __ldg(reinterpret_cast<const float2*>(sv));
      in.z = __ldg(reinterpret_cast<const float2*>(sv + H));
      in.n = __ldg(reinterpret_cast<const float2*>(sv + 2 * H));
      in.hn = __ldg(reinterpret_cast<const float2*>(sv + 3 * H));
      in.dout2 = __ldg(reinterpret_cast<const uint32_t*>(p.dout + rw * 2 * H + dir * H + j));
      if (tprev >= 0 && tprev < T)
        in.hp2 = __ldg(reinterpret_cast<const uint32_t*>(p.out + ((long long)b * T + tprev) * 2 * H + dir * H + j));
    }
    return in;
  };

  // recurrent gradient of the owned (row, unit pair) of every slice: a private slot per thread (dynamic slice index)
  for (int si = 0; si < kClMaxSlices; ++si) Dsm[si * kGruThreads + tid] = make_float2(0.f, 0.f);
  float sb_r[2] = {0.f, 0.f}, sb_z[2] = {0.f, 0.f}, sb_n[2] = {0.f, 0.f}, sb_nr[2] = {0.f, 0.f};

  // B fragments of two k-steps per ldmatrix.x4: matrices (lane >> 3) = k offsets 0, 8, 16, 24
  const uint32_t b_base = smem_u32(Wsm + (warp * NTW * 8 + (lane & 7)) * kBLd + (lane >> 3) * 8);

  struct Out {                          // one (slice, step)'s results for the GEMMs that follow the recurrence
    uint32_t a_r, a_z, a_n, a_nr, hp2;
    long long rw;
    bool ok;
  };
  // ---- phase A: gate gradients of the owned (row, unit pair) of slice si at time t ----
  auto phase_a = [&](const BwdIn& in, int si, int t, float (&dh_direct)[2]) {
    Out o;
    o.a_r = o.a_z = o.a_n = o.a_nr = 0u;
    o.hp2 = in.hp2;
    o.ok = in.ok;
    o.rw = 0;
    dh_direct[0] = dh_direct[1] = 0.f;
    if (in.ok) {
      const float2 rec = Dsm[si * kGruThreads + tid];
      const float dr[2] = {rec.x, rec.y};
      const float rr[2] = {in.r.x, in.r.y}, zz[2] = {in.z.x, in.z.y}, nn[2] = {in.n.x, in.n.y};
      const float hh[2] = {in.hn.x, in.hn.y};
      const float dd[2] = {__uint_as_float(in.dout2 << 16), __uint_as_float(in.dout2 & 0xffff0000u)};
      const float hp[2] = {__uint_as_float(in.hp2 << 16), __uint_as_float(in.hp2 & 0xffff0000u)};
      float dar[2], daz[2], dan[2], danr[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float dh = dd[e] + dr[e];
        const float dn = dh * (1.f - zz[e]);
        const float dz = dh * (hp[e] - nn[e]);
        dan[e] = dn * (1.f - nn[e] * nn[e]);
        daz[e] = dz * zz[e] * (1.f - zz[e]);
        dar[e] = dan[e] * hh[e] * rr[e] * (1.f - rr[e]);
        danr[e] = dan[e] * rr[e];
        sb_r[e] += dar[e];
        sb_z[e] += daz[e];
        sb_n[e] += dan[e];
        sb_nr[e] += danr[e];
        dh_direct[e] = dh * zz[e];
      }
      o.a_r = pack_bf16x2(dar[0], dar[1]);
      o.a_z = pack_bf16x2(daz[0], daz[1]);
      o.a_nr = pack_bf16x2(danr[0], danr[1]);
      o.a_n = pack_bf16x2(dan[0], dan[1]);
      const int b = (blockIdx.y + si * gridDim.y) * kBRows + row;
      o.rw = (long long)b * T + t;
    }
    return o;
  };
  auto store_out = [&](const Out& o) {
    if (!o.ok) return;
    const long long g0 = o.rw * 6 * H + dir * 3 * H + j;
    *reinterpret_cast<uint32_t*>(p.dgi + g0) = o.a_r;
    *reinterpret_cast<uint32_t*>(p.dgi + g0 + H) = o.a_z;
    *reinterpret_cast<uint32_t*>(p.dgi + g0 + 2 * H) = o.a_n;
    *reinterpret_cast<uint32_t*>(p.dgh + g0) = o.a_r;
    *reinterpret_cast<uint32_t*>(p.dgh + g0 + H) = o.a_z;
    *reinterpret_cast<uint32_t*>(p.dgh + g0 + 2 * H) = o.a_nr;
    *reinterpret_cast<uint32_t*>(p.hprev + (o.rw * 2 + dir) * H + j) = o.hp2;
  };
  // ---- reduce-scatter of partial tile `buf`: own 32 columns of every peer's tile, added in rank order.  16-byte
  //      requests: lanes l, l ^ 1 share a 4-unit group; each reads it from half of the ranks (even lane: ranks
  //      0 .. CS/2-1, odd lane: the rest), the halves meet through one shuffle ----
  const int hsel = tid & 1;
  auto pull = [&](int buf, int si, const float (&dh_direct)[2]) {
    const uint32_t pl = smem_u32(Psm + buf * kBRows * PLD + row * PLD + js * kJS + ((tid & 15) >> 1) * 4);
    float4 v[CS / 2];
#pragma unroll
    for (int s2 = 0; s2 < CS / 2; ++s2) v[s2] = ld_dsmem_f4(pl, (uint32_t)(hsel * (CS / 2) + s2));
    float4 sm = v[0];
#pragma unroll
    for (int s2 = 1; s2 < CS / 2; ++s2) {
      sm.x += v[s2].x;
      sm.y += v[s2].y;
      sm.z += v[s2].z;
      sm.w += v[s2].w;
    }
    float4 ot;
    ot.x = __shfl_xor_sync(0xffffffffu, sm.x, 1);
    ot.y = __shfl_xor_sync(0xffffffffu, sm.y, 1);
    ot.z = __shfl_xor_sync(0xffffffffu, sm.z, 1);
    ot.w = __shfl_xor_sync(0xffffffffu, sm.w, 1);
    // low ranks + high ranks, the same expression in both lanes
    const float s0 = hsel ? ot.z + sm.z : sm.x + ot.x;
    const float s1 = hsel ? ot.w + sm.w : sm.y + ot.y;
    Dsm[si * kGruThreads + tid] = make_float2(dh_direct[0] + s0, dh_direct[1] + s1);
  };

  // Items q = 0 .. Q-1 in (step, slice) order are the ones with a recurrent product (the last time step has none).
  // Item q: phase A, MMA into partial buffer q % 3, ARRIVE at the cluster barrier; its WAIT and reduce-scatter follow
  //  - with >= 2 slices per cluster after the NEXT item's phase A + MMA (that item belongs to another slice, its
  //    recurrent gradient is two pulls old): the barrier latency hides behind them.  Three partial buffers make
  //    that safe: buffer (q+1) % 3 was last read in pull q-2, which every peer finished before it arrived at
  //    barrier q-1, and barrier q-1 has been waited for.
  //  - with one slice per cluster before the next item's phase A (it needs this very pull).
  const bool pipe = nsl >= 2;
  const int Q = (T - 1) * nsl, items = T * nsl;
  BwdIn nxt = load_in(0, 0);
  __syncthreads();
  float dd_prev[2] = {0.f, 0.f};
  int si_prev = 0, buf_prev = 0;
  bool pending = false;
  int step = 0, si = 0;
  for (int q = 0; q < items; ++q) {
    const int t = dir == 0 ? T - 1 - step : step;
    const BwdIn in = nxt;
    const int si2 = si + 1 < nsl ? si + 1 : 0, step2 = si + 1 < nsl ? step : step + 1;
    if (q + 1 < items) nxt = load_in(step2, si2);   // inputs do not depend on the recurrence: latency hides here
    if (pending && (!pipe || q >= Q)) {             // one slice per cluster, or the tail: this item needs the pull
      cluster_wait_acquire();
      pull(buf_prev, si_prev, dd_prev);
      pending = false;
    }
    float dh_direct[2];
    const Out o = phase_a(in, si, t, dh_direct);
    if (q >= Q) {                                   // last time step: no recurrent product (dh_0 is not needed)
      store_out(o);
    } else {
      __nv_bfloat16* A = Asm + (q & 1) * kBRows * kBLd;
      *reinterpret_cast<uint32_t*>(A + row * kBLd + jl) = o.a_r;            // rows past the batch: zeros
      *reinterpret_cast<uint32_t*>(A + row * kBLd + kJS + jl) = o.a_z;
      *reinterpret_cast<uint32_t*>(A + row * kBLd + 2 * kJS + jl) = o.a_nr;
      __syncthreads();
      // ---- phase B: partial[16][H] = A[16][96] . Wsm[H][96]^T, this warp's NTW column tiles ----
      float acc[NTW][4];
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
      const uint32_t a_base = smem_u32(A + ((lane & 7) + ((lane >> 3) & 1) * 8) * kBLd + (lane >> 4) * 8);
      uint32_t a[6][4];
#pragma unroll
      for (int kk = 0; kk < 6; ++kk) ldmatrix_x4(a[kk], a_base + kk * 32);
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt) {
#pragma unroll
        for (int k2 = 0; k2 < 3; ++k2) {
          uint32_t bb[4];
          ldmatrix_x4(bb, b_base + nt * 8 * kBLd * 2 + k2 * 64);
          mma_16816(acc[nt], a[2 * k2], bb[0], bb[1]);
          mma_16816(acc[nt], a[2 * k2 + 1], bb[2], bb[3]);
        }
      }
      const int buf = q % 3;
      float* P = Psm + buf * kBRows * PLD;
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt) {
        const int col = (warp * NTW + nt) * 8 + (lane & 3) * 2;
        *reinterpret_cast<float2*>(P + (lane >> 2) * PLD + col) = make_float2(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<float2*>(P + ((lane >> 2) + 8) * PLD + col) = make_float2(acc[nt][2], acc[nt][3]);
      }
      if (pending) {                                // >= 2 slices: the previous item's barrier and reduce-scatter
        cluster_wait_acquire();                     // (waiting first and overlapping the PULL with this item's phase A
        pull(buf_prev, si_prev, dd_prev);           //  + MMA instead measured slower: 18.5 vs 16.0 us per step at B = 256)
      }
      cluster_arrive_release();                     // partial tile q is complete in this CTA
      store_out(o);                                 // global stores between the halves of the barrier
      pending = true;
      dd_prev[0] = dh_direct[0];
      dd_prev[1] = dh_direct[1];
      si_prev = si;
      buf_prev = buf;
    }
    si = si2;
    step = step2;
  }
  if (pending) {                                    // T == 1 never arrives; otherwise the tail consumed it
    cluster_wait_acquire();
    pending = false;
  }
  if (p.dbias) {
    // lanes l and l ^ 16 own the same unit pair on two rows; the 8 warps add their sums atomically
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      sb_r[e] += __shfl_xor_sync(0xffffffffu, sb_r[e], 16);
      sb_z[e] += __shfl_xor_sync(0xffffffffu, sb_z[e], 16);
      sb_n[e] += __shfl_xor_sync(0xffffffffu, sb_n[e], 16);
      sb_nr[e] += __shfl_xor_sync(0xffffffffu, sb_nr[e], 16);
    }
    if (lane < 16) {
      float* bih = p.dbias + (long long)dir * 3 * H;
      float* bhh = p.dbias + (long long)(2 + dir) * 3 * H;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        atomicAdd(bih + j + e, sb_r[e]);
        atomicAdd(bih + H + j + e, sb_z[e]);
        atomicAdd(bih + 2 * H + j + e, sb_n[e]);
        atomicAdd(bhh + j + e, sb_r[e]);
        atomicAdd(bhh + H + j + e, sb_z[e]);
        atomicAdd(bhh + 2 * H + j + e, sb_nr[e]);
      }
    }
  }
  cluster_sync_all();      // a CTA's shared memory must outlive its peers' last reads
}

}  // namespace m3t

using namespace m3t;

static int gru_launch(bool fwd, GruParams& p, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p.H % 32 != 0 || p.H > 512 || p.B <= 0 || p.T <= 0) return -1;
  const int bs_rows = fwd ? kFwdBS : kBwdBS;
  p.nbslices = (p.B + bs_rows - 1) / bs_rows;
  const int nsl = p.H / kJS;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int gy = sms / (2 * nsl);   // all CTAs must be co-resident (1 CTA per SM)
  if (gy < 1) return -2;
  if (gy > p.nbslices) gy = p.nbslices;
  if ((p.nbslices + gy - 1) / gy > kMaxSlicesPerCta) return -3;
  const size_t smem = fwd ? (size_t)(96 + 64) * (p.H + 8) * 2 : (size_t)64 * (3 * p.H + 8) * 2;
  auto kern = fwd ? gru_fwd_kernel : gru_bwd_kernel;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -20;
  if (cudaMemsetAsync(p.counters, 0, sizeof(unsigned) * 2 * p.nbslices, st) != cudaSuccess) return -22;
  dim3 grid(nsl, gy, 2);
  void* args[] = {&p};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)kern, grid, dim3(kGruThreads), args, smem, st);
  count_launch();
  if (e != cudaSuccess) return -21;
  return launch_status();
}

extern "C" int m3t_gru_fwd(const float* gi, const void* w_hh_bf16, const float* b_hh, void* out_bf16, float* out_f32,
                           float* saved, unsigned* counters, int B, int T, int H, void* stream) {
  GruParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.T = T; p.H = H;
  p.gi = gi;
  p.w = reinterpret_cast<const __nv_bfloat16*>(w_hh_bf16);
  p.b_hh = b_hh;
  p.out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  p.out_f32 = out_f32;
  p.saved = saved;
  p.counters = counters;
  return gru_launch(true, p, stream);
}

extern "C" int m3t_gru_fwd_cluster(const float* gi, const void* w_hh_bf16, const float* b_hh, void* out_bf16,
                                   float* out_f32, float* saved, int B, int T, int H, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (H % kCJS != 0 || H > 512 || B <= 0 || T <= 0) return -1;
  GruParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.T = T; p.H = H;
  p.gi = gi;
  p.w = reinterpret_cast<const __nv_bfloat16*>(w_hh_bf16);
  p.b_hh = b_hh;
  p.out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  p.out_f32 = out_f32;
  p.saved = saved;
  const int ncta = H / kCJS;     // 2, 4 or 8: a portable cluster size
  const size_t smem = (size_t)(3 * kCJS + kCRows) * (H + 8) * 2 + (size_t)2 * kCRows * kCJS * 2;
  if (cudaFuncSetAttribute(gru_fwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
      cudaSuccess)
    return -20;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(ncta, (B + kCRows - 1) / kCRows, 2);
  cfg.blockDim = dim3(kGruThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ncta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // every cluster must be resident at once (a second wave would double the sequence latency): 8 clusters (64 rows)
  // always are; beyond that ask the occupancy calculator for this cluster size and shared-memory footprint
  static int max_clusters[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};     // by cluster size, stored + 1 (0 = not asked yet)
  const int need = 2 * ((B + kCRows - 1) / kCRows);
  if (need > 8) {
    if (!max_clusters[ncta]) {
      cudaLaunchConfig_t q = cfg;
      q.gridDim = dim3(ncta, 8, 2);
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, gru_fwd_cluster_kernel, &q) != cudaSuccess) {
        cudaGetLastError();
        n = 8;
      }
      max_clusters[ncta] = n + 1;
    }
    if (need > max_clusters[ncta] - 1) return -3;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, gru_fwd_cluster_kernel, p);
  count_launch();
  if (e != cudaSuccess) return -21;
  return launch_status();
}

extern "C" int m3t_gru_bwd(const void* dout_bf16, const void* out_bf16, const float* saved, const void* w_hh_t_bf16,
                           void* dgi_bf16, void* dgh_bf16, void* hprev_bf16, unsigned* counters, float* dbias, int B,
                           int T, int H, void* stream) {
  GruParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.T = T; p.H = H;
  p.dout = reinterpret_cast<const __nv_bfloat16*>(dout_bf16);
  p.out = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(out_bf16));
  p.saved = const_cast<float*>(saved);
  p.w = reinterpret_cast<const __nv_bfloat16*>(w_hh_t_bf16);
  p.dgi = reinterpret_cast<__nv_bfloat16*>(dgi_bf16);
  p.dgh = reinterpret_cast<__nv_bfloat16*>(dgh_bf16);
  p.hprev = reinterpret_cast<__nv_bfloat16*>(hprev_bf16);
  p.counters = counters;
  p.dbias = dbias;
  return gru_launch(false, p, stream);
}

template <int CS>
static int gru_bwd_cluster_launch(GruParams& p, cudaStream_t st, int* query_max) {
  constexpr int H = CS * kJS;
  const size_t smem = (size_t)(H + 2 * kBRows) * kBLd * 2 + (size_t)3 * kBRows * (H + 8) * 4 +
                      (size_t)kClMaxSlices * kGruThreads * 8;
  auto kern = gru_bwd_cluster_kernel<CS>;
  static int max_clusters = -1;        // co-resident clusters of this size on the device (0: cannot launch)
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(kGruThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (max_clusters < 0) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -20;
    if (CS > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      cudaGetLastError();
      max_clusters = 0;
      return -23;
    }
    cfg.gridDim = dim3(CS, 8, 2);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
      cudaGetLastError();
      n = 0;
    }
    max_clusters = n;
  }
  if (query_max) { *query_max = max_clusters; return 0; }
  if (max_clusters < 2) return -23;
  p.nbslices = (p.B + kBRows - 1) / kBRows;
  int ncl = max_clusters / 2;          // per direction
  if (ncl > p.nbslices) ncl = p.nbslices;
  const int spc = (p.nbslices + ncl - 1) / ncl;
  if (spc > kClMaxSlices) return -3;
  ncl = (p.nbslices + spc - 1) / spc;
  cfg.gridDim = dim3(CS, ncl, 2);
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  count_launch();
  if (e != cudaSuccess) return -21;
  return launch_status();
}

// Cluster / DSMEM BPTT (see gru_bwd_cluster_kernel).  Same arguments and results as m3t_gru_bwd without the arrival
// counters; H must be 128, 256 or 512.  Returns -23 when the device cannot co-schedule a cluster of H/32 CTAs with
// this kernel's shared memory and -3 when the batch needs more than 6 slices per cluster (callers fall back to
// m3t_gru_bwd).
extern "C" int m3t_gru_bwd_cluster(const void* dout_bf16, const void* out_bf16, const float* saved,
                                   const void* w_hh_t_bf16, void* dgi_bf16, void* dgh_bf16, void* hprev_bf16,
                                   float* dbias, int B, int T, int H, void* stream) {
  if (B <= 0 || T <= 0) return -1;
  GruParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.T = T; p.H = H;
  p.dout = reinterpret_cast<const __nv_bfloat16*>(dout_bf16);
  p.out = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(out_bf16));
  p.saved = const_cast<float*>(saved);
  p.w = reinterpret_cast<const __nv_bfloat16*>(w_hh_t_bf16);
  p.dgi = reinterpret_cast<__nv_bfloat16*>(dgi_bf16);
  p.dgh = reinterpret_cast<__nv_bfloat16*>(dgh_bf16);
  p.hprev = reinterpret_cast<__nv_bfloat16*>(hprev_bf16);
  p.dbias = dbias;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (H == 128) return gru_bwd_cluster_launch<4>(p, st, nullptr);
  if (H == 256) return gru_bwd_cluster_launch<8>(p, st, nullptr);
  if (H == 512) return gru_bwd_cluster_launch<16>(p, st, nullptr);
  return -1;
}

// How many clusters of m3t_gru_bwd_cluster's kernel for hidden size H the device keeps resident at once
// (cudaOccupancyMaxActiveClusters; 0 = none, negative = error / unsupported H).
extern "C" int m3t_gru_bwd_cluster_max(int H) {
  GruParams p;
  memset(&p, 0, sizeof(p));
  int n = 0, rc = -1;
  if (H == 128) rc = gru_bwd_cluster_launch<4>(p, nullptr, &n);
  if (H == 256) rc = gru_bwd_cluster_launch<8>(p, nullptr, &n);
  if (H == 512) rc = gru_bwd_cluster_launch<16>(p, nullptr, &n);
  return rc ? rc : n;
}
