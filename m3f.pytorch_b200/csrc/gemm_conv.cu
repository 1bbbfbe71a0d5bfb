// Host launchers + C-ABI entry points for the tcgen05 pipeline (umma_kernel.cuh):
//   m3t_gemm_bf16, m3t_conv_fprop_bf16, m3t_conv_wgrad_bf16.
#include "../../include/m3t_b200.h"
#include "common.cuh"
#include "tmap.cuh"
#include "umma_kernel.cuh"
#include "umma_persist.cuh"

namespace m3t {

template <int BN, int MT, int STAGES, int AKIND, bool A_MN, bool B_MN, int EPI, int CK = 64>
static int launch_umma(const CUtensorMap& tmA, const CUtensorMap& tmB, const UmmaParams& p, int tiles_m, int splits,
                       cudaStream_t st) {
  auto kern = umma_kernel<BN, MT, STAGES, AKIND, A_MN, B_MN, EPI, CK>;
  constexpr int smem = umma_smem_bytes<BN, MT, STAGES>();
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return -20;
    attr_done = true;
  }
  dim3 grid((unsigned)(tiles_m * p.tiles_n), (unsigned)splits, 1);
  m3t::launch_k(kern, dim3(grid), dim3(kUmmaThreads), smem, st, tmA, tmB, p);
  count_launch();
  return launch_status();
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

template <int BN, int MT, int STAGES, int AKIND, bool TCN = false>
static int launch_persist(const CUtensorMap& tmA, const CUtensorMap& tmB, const UmmaParams& p, int tiles_m,
                          cudaStream_t st) {
  auto kern = umma_persist_kernel<BN, MT, STAGES, AKIND, TCN>;
  constexpr int smem = umma_persist_smem_bytes<BN, MT, STAGES>();
  static_assert(smem <= 227 * 1024, "shared memory");
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -20;
    attr_done = true;
  }
  const int sms = m3t::usable_sms();
  const int num_tiles = tiles_m * p.tiles_n;
  int grid = num_tiles < sms ? num_tiles : sms;
  if (p.tiles_n <= grid) grid -= grid % p.tiles_n;   // a CTA then keeps one column block (register-resident BN stats)
  m3t::launch_k(kern, dim3(grid), dim3(kUmmaThreads), smem, st, tmA, tmB, p, num_tiles);
  count_launch();
  return launch_status();
}

}  // namespace m3t

using namespace m3t;

extern "C" int m3t_gemm_bf16(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, void* D,
                             long long ldd, int d_f32, int M, int N, int K, const float* scale, const float* shift,
                             const void* residual, long long ldr, int relu, float* stats, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return -1;
  if (a_mn && !b_mn) return -2;  // (MN, K) never occurs on the hot path
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N;
  p.k_iters = ceil_div(K, kBlockK);
  p.out = D; p.ldc = ldd; p.out_f32 = d_f32;
  p.scale = scale; p.shift = shift;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual); p.ldr = ldr;
  p.relu = relu; p.stats = stats;
  const int tiles_m = ceil_div(M, 128);
  const int bn = (N <= 32 && !b_mn) ? 32 : (N <= 64 ? 64 : 128);
  p.tiles_n = ceil_div(N, bn);
  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = make_tmap_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, 128);
  else rc = make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, 64);
  if (rc) return rc;
  if (!b_mn) rc = make_tmap_2d_bf16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, (uint32_t)bn);
  else rc = make_tmap_2d_bf16(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, 64);
  if (rc) return rc;
  // K-major x K-major with more tiles than SMs: the persistent tile walker (epilogue of tile i overlaps tile i+1)
  if (!a_mn && !b_mn && p.k_iters >= 2 && (long long)tiles_m * p.tiles_n > 148) {
    if (bn == 32) return launch_persist<32, 1, 8, A_TILED>(tmA, tmB, p, tiles_m, st);
    if (bn == 64) return launch_persist<64, 1, 8, A_TILED>(tmA, tmB, p, tiles_m, st);
    return launch_persist<128, 1, 6, A_TILED>(tmA, tmB, p, tiles_m, st);
  }
#define GEMM_CASE(BN_, AMN_, BMN_)                                                                         \
  if (bn == BN_ && (bool)a_mn == AMN_ && (bool)b_mn == BMN_)                                               \
    return launch_umma<BN_, 1, (BN_ >= 128 ? 4 : 4), A_TILED, AMN_, BMN_, EPI_STORE>(tmA, tmB, p, tiles_m, 1, st);
  GEMM_CASE(32, false, false)
  GEMM_CASE(64, false, false)
  GEMM_CASE(128, false, false)
  GEMM_CASE(64, false, true)
  GEMM_CASE(128, false, true)
  GEMM_CASE(64, true, true)
  GEMM_CASE(128, true, true)
#undef GEMM_CASE
  return -3;
}

namespace {

struct ConvGeom {
  int nd;                 // spatial dims 1..3
  int N, D, H, W, Cin, Cout;
  int kd, kh, kw, sd, sh, sw, dd, dh, dw;
  int pdl, pdu, phl, phu, pwl, pwu;
  int Z, P, Q;
};

int conv_geom(ConvGeom& g, const int* v) {
  // v: nd, N, D, H, W, Cin, Cout, kd, kh, kw, sd, sh, sw, pdl, pdu, phl, phu, pwl, pwu, dd, dh, dw
  g.nd = v[0]; g.N = v[1]; g.D = v[2]; g.H = v[3]; g.W = v[4]; g.Cin = v[5]; g.Cout = v[6];
  g.kd = v[7]; g.kh = v[8]; g.kw = v[9]; g.sd = v[10]; g.sh = v[11]; g.sw = v[12];
  g.pdl = v[13]; g.pdu = v[14]; g.phl = v[15]; g.phu = v[16]; g.pwl = v[17]; g.pwu = v[18];
  g.dd = v[19]; g.dh = v[20]; g.dw = v[21];
  if (g.nd < 1 || g.nd > 3) return -1;
  if (g.nd < 3) { if (g.D != 1 || g.kd != 1) return -1; g.sd = 1; g.dd = 1; g.pdl = g.pdu = 0; }
  if (g.nd < 2) { if (g.H != 1 || g.kh != 1) return -1; g.sh = 1; g.dh = 1; g.phl = g.phu = 0; }
  if ((g.Cin % 64 != 0 && g.Cin != 16) || g.Cout % 8 != 0) return -4;
  if (g.Cin == 16 && (g.kd * g.kh * g.kw) % 4 != 0) return -4;  // 16-channel mode packs 4 taps per stage
  g.Q = (g.W + g.pwl + g.pwu - g.dw * (g.kw - 1) - 1) / g.sw + 1;
  g.P = (g.H + g.phl + g.phu - g.dh * (g.kh - 1) - 1) / g.sh + 1;
  g.Z = (g.D + g.pdl + g.pdu - g.dd * (g.kd - 1) - 1) / g.sd + 1;
  if (g.Q <= 0 || g.P <= 0 || g.Z <= 0) return -5;
  return 0;
}

int conv_ck(const ConvGeom& g) { return g.Cin == 16 ? 16 : 64; }

int conv_tmap(CUtensorMap* tm, const void* x, const ConvGeom& g, uint32_t pixels_per_column) {
  const int rank = g.nd + 2;
  uint64_t dims[5];
  int lower[3], upper[3], cs[3];
  dims[0] = (uint64_t)g.Cin;
  dims[1] = (uint64_t)g.W; lower[0] = -g.pwl; upper[0] = g.pwu - (g.kw - 1) * g.dw; cs[0] = g.sw;
  int i = 2;
  if (g.nd >= 2) { dims[i] = (uint64_t)g.H; lower[1] = -g.phl; upper[1] = g.phu - (g.kh - 1) * g.dh; cs[1] = g.sh; ++i; }
  if (g.nd >= 3) { dims[i] = (uint64_t)g.D; lower[2] = -g.pdl; upper[2] = g.pdu - (g.kd - 1) * g.dd; cs[2] = g.sd; ++i; }
  dims[i] = (uint64_t)g.N;
  const int ck = conv_ck(g);
  return make_tmap_im2col_bf16(tm, x, rank, dims, lower, upper, cs, (uint32_t)ck, pixels_per_column,
                               ck == 64 ? 128 : 32);
}

void conv_fill_params(UmmaParams& p, const ConvGeom& g) {
  p.rank = g.nd + 2;
  p.Q = g.Q; p.P = g.P; p.Z = g.Z;
  p.sw = g.sw; p.sh = g.sh; p.sd = g.sd;
  p.pw = g.pwl; p.ph = g.phl; p.pd = g.pdl;
  p.dw = g.dw; p.dh = g.dh; p.dd = g.dd;
  p.S = g.kw; p.R = g.kh; p.T = g.kd;
  p.Cin = g.Cin;
  p.cblocks = g.Cin / conv_ck(g);
}

}  // namespace

struct TcnEpilogue {      // fused TemporalBlock epilogue (UmmaParams::pre_act ...)
  float drop_p;
  unsigned long long seed;
  void* t_out;
  const unsigned long long* seed_dev;
};
static int conv_fprop_impl(const void* x, const void* w_packed, void* y, const int* geom, const float* scale,
                           const float* shift, const void* residual, int relu, float* stats, int tile_hint,
                           const long long* out_map, void* stream, const TcnEpilogue* tcn = nullptr);

// When the persistent kernel is the default (tests/tune_conv.py, 4096 frames): it wins whenever a tile has at least
// two k-iterations (28->14 64->128 s2 0.198 -> 0.166 ms, 7x7x256 0.214 -> 0.189, 4x4x512 0.255 -> 0.226 = 1367
// TFLOP/s, parity sub-convolutions 0.078 -> 0.071); with a single k-iteration (1x1, 64 channels) two co-resident
// one-tile CTAs hide the epilogue better (0.107 vs 0.146 ms).
static bool conv_use_persist(const ConvGeom& g, const UmmaParams& p, int bn, int mt) {
  (void)g; (void)bn; (void)mt;
  return p.k_iters >= 2;
}

extern "C" int m3t_conv_fprop_bf16(const void* x, const void* w_packed, void* y, const int* geom, const float* scale,
                                   const float* shift, const void* residual, int relu, float* stats, int tile_hint,
                                   void* stream) {
  return conv_fprop_impl(x, w_packed, y, geom, scale, shift, residual, relu, stats, tile_hint, nullptr, stream);
}

extern "C" int m3t_tcn_conv_bf16(const void* x, const void* w_packed, void* y, void* t_out, const int* geom,
                                 const float* scale, const float* shift, const void* residual, float drop_p,
                                 unsigned long long seed, void* stream) {
  if (!(drop_p >= 0.f) || !(drop_p < 1.f) || (t_out && !residual)) return -1;
  TcnEpilogue e;
  e.drop_p = drop_p; e.seed = seed; e.t_out = t_out; e.seed_dev = nullptr;
  return conv_fprop_impl(x, w_packed, y, geom, scale, shift, residual, residual ? 1 : 0, nullptr, 0, nullptr, stream, &e);
}

// The same with the mask seed = seed + *seed_dev (a 64-bit device-resident counter the caller advances once per
// training step): nothing about the mask is a launch argument that a captured CUDA graph would freeze.
extern "C" int m3t_tcn_conv_bf16_dseed(const void* x, const void* w_packed, void* y, void* t_out, const int* geom,
                                       const float* scale, const float* shift, const void* residual, float drop_p,
                                       unsigned long long seed, const unsigned long long* seed_dev, void* stream) {
  if (!(drop_p >= 0.f) || !(drop_p < 1.f) || (t_out && !residual)) return -1;
  TcnEpilogue e;
  e.drop_p = drop_p; e.seed = seed; e.t_out = t_out; e.seed_dev = seed_dev;
  return conv_fprop_impl(x, w_packed, y, geom, scale, shift, residual, residual ? 1 : 0, nullptr, 0, nullptr, stream, &e);
}

extern "C" int m3t_conv_fprop_scatter_bf16(const void* x, const void* w_packed, void* y, const int* geom,
                                           long long img_pitch, long long row_pitch, long long px_pitch,
                                           int accumulate, int tile_hint, void* stream) {
  const long long map[3] = {img_pitch, row_pitch, px_pitch};
  return conv_fprop_impl(x, w_packed, y, geom, nullptr, nullptr, accumulate ? y : nullptr, 0, nullptr, tile_hint, map,
                         stream);
}

static int conv_fprop_impl(const void* x, const void* w_packed, void* y, const int* geom, const float* scale,
                           const float* shift, const void* residual, int relu, float* stats, int tile_hint,
                           const long long* out_map, void* stream, const TcnEpilogue* tcn) {
  ConvGeom g;
  int rc = conv_geom(g, geom);
  if (rc) return rc;
  if (out_map && g.Z != 1) return -1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long Mpix = (long long)g.N * g.Z * g.P * g.Q;
  const int taps = g.kd * g.kh * g.kw;
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  conv_fill_params(p, g);
  p.M = (int)Mpix; p.N = g.Cout;
  const int ck = conv_ck(g);
  p.k_iters = taps * g.Cin / kBlockK;
  p.out = y; p.ldc = g.Cout; p.out_f32 = (tile_hint & 64) ? 1 : 0;
  if (p.out_f32 && residual) return -1;
  p.cb_major = (tile_hint & 128) ? 1 : 0;
  p.scale = scale; p.shift = shift;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual); p.ldr = g.Cout;
  p.relu = relu; p.stats = stats;
  const bool det = (tile_hint & 256) != 0 && stats != nullptr;   // deterministic BatchNorm statistics (slots + m3t_det_reduce)
  if (det) p.det_stride = 2LL * g.Cout;
  if (tcn) {
    p.pre_act = 1;
    if (tcn->drop_p > 0.f) {
      p.drop_thresh = dropout_threshold(tcn->drop_p);
      p.drop_scale = 1.f / (1.f - tcn->drop_p);
      p.drop_seed = tcn->seed;
      p.drop_seed_dev = tcn->seed_dev;
    }
    p.out2 = reinterpret_cast<__nv_bfloat16*>(tcn->t_out);
  }
  if (out_map) {
    p.oq = g.Q; p.op = g.P;
    p.o_img = out_map[0]; p.o_row = out_map[1]; p.o_px = out_map[2];
    p.res_mapped = 1;   // a residual here is the output itself (accumulate)
  }
  // tile selection: BN = 64 / 128 / 256 ; MT = 2 when there is enough M to still fill the machine
  // 256-column tiles halve the A re-reads of the wide layers (measured +3..7 % on the 256/512-channel convs);
  // tile_hint bit3 forces 128 columns.
  int bn = g.Cout <= 64 ? 64 : (g.Cout % 256 == 0 && !(tile_hint & 8) ? 256 : 128);
  int mt = 1;
  if (tile_hint & 2) mt = 2;
  else if (!(tile_hint & 1)) {
    const long long ctas_mt2 = (Mpix / 256) * ((g.Cout + bn - 1) / bn);
    if (ctas_mt2 >= 2 * 148 * 2) mt = 2;
  }
  if (bn == 256) mt = 1;
  p.tiles_n = ceil_div(g.Cout, bn);
  const int tiles_m = ceil_div(Mpix, 128 * mt);
  CUtensorMap tmA, tmB;
  rc = conv_tmap(&tmA, x, g, 128);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmB, w_packed, (uint64_t)taps * g.Cin, (uint64_t)g.Cout, (uint64_t)taps * g.Cin,
                         (uint32_t)ck, (uint32_t)bn, ck == 64 ? 128 : 32);
  if (rc) return rc;
  if (ck == 16) {
    if (bn != 64) return -4;
    if (det) return -9;      // the 16-channel im2col mode has no persistent variant, hence no slotted statistics
    if (mt == 2) return launch_umma<64, 2, 2, A_IM2COL, false, false, EPI_STORE, 16>(tmA, tmB, p, tiles_m, 1, st);
    return launch_umma<64, 1, 4, A_IM2COL, false, false, EPI_STORE, 16>(tmA, tmB, p, tiles_m, 1, st);
  }
  // persistent tile walker (umma_persist.cuh): tile_hint bit4 forces it, bit5 forbids it
  const bool persist = det || (tile_hint & 16) != 0 || (!(tile_hint & 32) && conv_use_persist(g, p, bn, mt));
  if (tcn) {       // fused TemporalBlock epilogue: its own instantiations of the persistent kernel
    if (!persist) return -5;
    if (bn == 64) return launch_persist<64, 1, 8, A_IM2COL, true>(tmA, tmB, p, ceil_div(Mpix, 128), st);
    if (bn == 128) return launch_persist<128, 1, 6, A_IM2COL, true>(tmA, tmB, p, ceil_div(Mpix, 128), st);
    return launch_persist<256, 1, 4, A_IM2COL, true>(tmA, tmB, p, ceil_div(Mpix, 128), st);
  }
  if (persist) {
    if (bn == 64 && mt == 1) return launch_persist<64, 1, 8, A_IM2COL>(tmA, tmB, p, tiles_m, st);
    if (bn == 64 && mt == 2) return launch_persist<64, 2, 5, A_IM2COL>(tmA, tmB, p, tiles_m, st);
    if (bn == 128 && mt == 1) return launch_persist<128, 1, 6, A_IM2COL>(tmA, tmB, p, tiles_m, st);
    if (bn == 128 && mt == 2) return launch_persist<128, 2, 4, A_IM2COL>(tmA, tmB, p, tiles_m, st);
    if (bn == 256) return launch_persist<256, 1, 4, A_IM2COL>(tmA, tmB, p, tiles_m, st);
  }
  if (bn == 64 && mt == 1) return launch_umma<64, 1, 4, A_IM2COL, false, false, EPI_STORE>(tmA, tmB, p, tiles_m, 1, st);
  if (bn == 64 && mt == 2) return launch_umma<64, 2, 2, A_IM2COL, false, false, EPI_STORE>(tmA, tmB, p, tiles_m, 1, st);
  if (bn == 128 && mt == 1) return launch_umma<128, 1, 3, A_IM2COL, false, false, EPI_STORE>(tmA, tmB, p, tiles_m, 1, st);
  if (bn == 128 && mt == 2) return launch_umma<128, 2, 2, A_IM2COL, false, false, EPI_STORE>(tmA, tmB, p, tiles_m, 1, st);
  if (bn == 256) return launch_umma<256, 1, 2, A_IM2COL, false, false, EPI_STORE>(tmA, tmB, p, tiles_m, 1, st);
  return -3;
}

static int conv_wgrad_impl(const void* x, const void* dy, float* dw_packed, const int* geom, int splits_hint,
                           void* stream, int* query_splits);

extern "C" int m3t_conv_wgrad_bf16(const void* x, const void* dy, float* dw_packed, const int* geom, int splits_hint,
                                   void* stream) {
  return conv_wgrad_impl(x, dy, dw_packed, geom, splits_hint, stream, nullptr);
}

extern "C" int m3t_conv_wgrad_splits(const int* geom, int splits_hint) {
  int splits = 0;
  const int rc = conv_wgrad_impl(nullptr, nullptr, nullptr, geom, splits_hint, nullptr, &splits);
  return rc ? rc : splits;
}

static int conv_wgrad_impl(const void* x, const void* dy, float* dw_packed, const int* geom, int splits_hint,
                           void* stream, int* query_splits) {
  ConvGeom g;
  int rc = conv_geom(g, geom);
  if (rc) return rc;
  if (g.Cout % 64 != 0) return -4;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long Mpix = (long long)g.N * g.Z * g.P * g.Q;
  const int taps = g.kd * g.kh * g.kw;
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  conv_fill_params(p, g);
  p.atoms = taps * p.cblocks;
  p.M = taps * g.Cin;  // rows of D = (tap, ci)
  p.N = g.Cout;
  const int bn = g.Cout <= 64 ? 64 : 128;
  const int ck = conv_ck(g);
  p.tiles_n = ceil_div(g.Cout, bn);
  // 256-row CTA tiles (two accumulators share every dY stage) whenever there are enough M atoms
  // (bit 30 of splits_hint forces 128-row tiles).  With 128 columns the 256-row tile is used when the M atoms pair
  // up evenly (measured, 4096 frames: 14x14x128 0.357 -> 0.328 ms, 7x7x256 0.368 -> 0.305, 4x4x512 0.472 -> 0.386;
  // 9 atoms (64 -> 128 stride 2) lose to the padding; 256x256 tiles and 3-stage / 1-CTA-per-SM variants were no faster)
  const bool force_mt1 = (splits_hint & (1 << 30)) != 0;
  const bool det = (splits_hint & (1 << 29)) != 0;    // every split accumulates into its own copy of dw_packed
  splits_hint &= ~(1 << 29);
  int mt = (ck == 64 && p.atoms >= 4 && bn == 64 && !force_mt1) ? 2 : 1;
  if (ck == 64 && bn == 128 && p.atoms % 4 == 0 && !force_mt1) mt = 2;
  splits_hint &= ~(1 << 30);
  const int tiles_m = ceil_div(p.atoms, mt * 128 / ck);
  const int kblocks = ceil_div(Mpix, kBlockK);
  int splits = splits_hint;
  if (splits <= 0) {
    // fill exactly two waves of CTAs (2 CTAs/SM resident): one CTA more than a wave costs a whole extra wave
    // (measured: 612 CTAs 0.504 ms vs 576 CTAs 0.366 ms on the 7x7x256 layer)
    const int sms = m3t::usable_sms();
    const int tiles = tiles_m * p.tiles_n;
    splits = (2 * 2 * sms) / tiles;
    const int max_splits = kblocks / 4 > 0 ? kblocks / 4 : 1;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  p.k_iters = ceil_div(kblocks, splits);
  splits = ceil_div(kblocks, p.k_iters);   // never more than asked for: the wave count is preserved
  p.out = dw_packed; p.ldc = (long long)taps * g.Cin; p.out_f32 = 1;
  if (det) p.det_stride = (long long)g.Cout * taps * g.Cin;
  if (query_splits) { *query_splits = splits; return 0; }
  CUtensorMap tmA, tmB;
  rc = conv_tmap(&tmA, x, g, 64);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmB, dy, (uint64_t)g.Cout, (uint64_t)Mpix, (uint64_t)g.Cout, 64, 64);
  if (rc) return rc;
  if (ck == 16) {
    if (bn != 64) return -4;
    return launch_umma<64, 1, 4, A_WGRAD, false, true, EPI_ATOMIC_T, 16>(tmA, tmB, p, tiles_m, splits, st);
  }
  if (bn == 64 && mt == 2)
    return launch_umma<64, 2, 2, A_WGRAD, false, true, EPI_ATOMIC_T>(tmA, tmB, p, tiles_m, splits, st);
  if (bn == 64) return launch_umma<64, 1, 4, A_WGRAD, false, true, EPI_ATOMIC_T>(tmA, tmB, p, tiles_m, splits, st);
  if (mt == 2) return launch_umma<128, 2, 2, A_WGRAD, false, true, EPI_ATOMIC_T>(tmA, tmB, p, tiles_m, splits, st);
  return launch_umma<128, 1, 3, A_WGRAD, false, true, EPI_ATOMIC_T>(tmA, tmB, p, tiles_m, splits, st);
}
