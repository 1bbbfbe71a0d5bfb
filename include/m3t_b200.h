/* m3t_b200 — C ABI of the B200-native M3T audio-visual backbone hot path.
 *
 * The reference (sailordiary/m3f.pytorch) has no FFI: its hot path is a stack of torch.nn modules
 * (models/backbone.py, resnet.py, tcn.py, rnn.py, att_fusion.py, model.py) whose device work PyTorch hands to
 * cuDNN / cuBLAS / ATen.  This header is the boundary a replacement sits behind: every entry point states the
 * reference call site (file:line under /root/reference) whose implicit library kernel it replaces.  The Python
 * host side (m3f.pytorch_b200/ops.py) binds these with ctypes and keeps the reference's nn.Module API.
 *
 * Conventions
 *   - plain pointers + sizes only; all pointers are DEVICE pointers unless the name ends in _host;
 *   - activations are channels-last bf16 (N[,D],H,W,C); parameters/gradients are fp32 in PyTorch layout unless a
 *     "packed" bf16 copy is asked for; statistics and reductions are fp32;
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises, allocates or frees;
 *   - return 0 on success, negative on error (bad shape, tensor-map failure, launch failure) — never throws;
 *   - thread-safe and re-entrant across streams (no mutable global state beyond one-time function attributes).
 */
#ifndef M3T_B200_H_
#define M3T_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

/* ABI version; bumps on any signature change. */
int m3t_abi_version(void);
/* Number of kernels launched by this library since load (bench.py reports it as gpu_launches). */
long long m3t_launch_count(void);
/* Programmatic dependent launch (csrc/common.cuh): every kernel of the library is launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization and orders itself behind the previous kernel of its stream with
 * griddepcontrol.wait, so launch latency and CTA scheduling overlap the previous kernel's tail.  on = 0 / 1 sets the
 * switch (default: environment M3T_PDL == "1"; m3t_b200.engine turns it on while it captures a small-shard training
 * step), on < 0 only queries; returns the previous setting. */
int m3t_set_pdl(int on);
/* The persistent one-CTA-per-SM kernels (conv / GEMM tile walkers, halo kernels) size their grids for the device's SM
 * count minus this reserve (default 0).  m3t_b200.engine sets it while a bucket of the gradient all-reduce overlaps
 * the backward pass, so that NCCL's CTAs find free SMs instead of turning one-wave launches into two-wave ones.
 * n < 0 only queries; returns the previous value. */
int m3t_set_sm_reserve(int n);

/* ------------------------------------------------------------------------------------------------------------
 * Tensor-core GEMM (tcgen05 / TMEM / TMA).  D[M,N] = act( (A . B^T) * scale[n] + shift[n] + residual[m,n] )
 *   A: a_mn == 0 -> stored [M][lda] (K contiguous)   a_mn == 1 -> stored [K][lda] (M contiguous)
 *   B: b_mn == 0 -> stored [N][ldb] (K contiguous)   b_mn == 1 -> stored [K][ldb] (N contiguous)
 *   D: bf16 (d_f32 == 0) or fp32 (d_f32 == 1), row stride ldd.  scale/shift/residual/stats may be NULL.
 *   stats: fp32 [2][N]; column sums and sums of squares of the raw accumulator are atomically added.
 * Replaces: nn.Linear (models/rnn.py:20-55, models/model.py:88), the W_ih x projection inside nn.GRU
 * (models/rnn.py:17,75) and their backward GEMMs (dX = dY.W, dW = dY^T.X).
 * Leading dimensions must be multiples of 8 elements (16 bytes, a TMA requirement). */
int m3t_gemm_bf16(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, void* D,
                  long long ldd, int d_f32, int M, int N, int K, const float* scale, const float* shift,
                  const void* residual, long long ldr, int relu, float* stats, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution, forward (and stride-1 dgrad when called on dY with the flipped/transposed filter).
 *   geom[22] = { nd, N, D, H, W, Cin, Cout, kd, kh, kw, sd, sh, sw, pdl, pdu, phl, phu, pwl, pwu, dd, dh, dw }
 *     nd = number of spatial dims (1: Conv1d over W, 2: Conv2d, 3: Conv3d); unused dims are 1 / pads 0.
 *     p?l / p?u = lower / upper zero padding (they differ for the causal TCN convolution).
 *   x        bf16 channels-last [N][D][H][W][Cin]      (Cin % 64 == 0)
 *   w_packed bf16 [Cout][kd*kh*kw*Cin]                 (tap-major, channel-minor)
 *   y        bf16 channels-last [N][Z][P][Q][Cout]
 *   y = act( conv(x,w) * scale[c] + shift[c] + residual ), all optional; stats as in m3t_gemm_bf16.
 *   tile_hint: 0 = auto; bit0 force 128-row tiles, bit1 force 256-row tiles, bit3 force 128-column tiles,
 *     bit4 / bit5 force / forbid the persistent tile walker, bit6: y is float32 (no residual; fp32-parity mode),
 *     bit7: contraction runs channel-block-major (all taps of block 0, then block 1, ...) instead of tap-major.
 * Replaces: nn.Conv2d 3x3 / 1x1 in BasicBlock (models/resnet.py:7-15,24-27,40-54,98-101), nn.Conv3d 3x3x3 in
 * VA_3DVGGM(_Split) (models/backbone.py:73-103,179-195,243-271), weight-normed dilated causal nn.Conv1d in
 * TemporalBlock (models/tcn.py:19-33) and Conv1d k5 in tcn_simple (models/backbone.py:214-231), with the
 * BatchNorm / residual / ReLU that follow them folded into the epilogue (eval) or their statistics (train). */
int m3t_conv_fprop_bf16(const void* x, const void* w_packed, void* y, const int* geom, const float* scale,
                        const float* shift, const void* residual, int relu, float* stats, int tile_hint,
                        void* stream);

/* One weight-normed dilated causal Conv1d of a TemporalBlock with everything that follows it in the block fused into
 * the epilogue (models/tcn.py:19-33 conv -> Chomp1d -> ReLU -> Dropout, :43-46 relu(net(x) + res)):
 *   t = dropout_p(relu(conv(x) * scale[c] + shift[c]))         scale = weight_g / ||weight_v|| (weight-norm), shift = bias
 *   residual == NULL:  y = t                                    (first conv of the block)
 *   residual != NULL:  y = relu(t + residual), t_out = t        (second conv; t_out may be NULL at inference)
 * x channels-last bf16 [B,T,Cin], w_packed = bf16 pack of weight_v, geom as m3t_conv_fprop_bf16 (nd = 1, left-only
 * padding (k-1)*dilation: Chomp1d is index math).  Dropout mask = the counter-based generator of m3t_dropout_bf16 over
 * the element index of y (row * Cout + c).  Replaces cuDNN conv1d + 2 `.contiguous()` copies + ATen relu / dropout /
 * add / relu per block in the reference. */
int m3t_tcn_conv_bf16(const void* x, const void* w_packed, void* y, void* t_out, const int* geom, const float* scale,
                      const float* shift, const void* residual, float drop_p, unsigned long long seed, void* stream);
/* The same with mask seed = seed + *seed_dev, seed_dev a 64-bit counter in device memory that the caller advances once
 * per training step: a step captured into a CUDA graph then draws a new mask on every replay (the launch arguments,
 * frozen by the capture, carry only the per-layer offset). */
int m3t_tcn_conv_bf16_dseed(const void* x, const void* w_packed, void* y, void* t_out, const int* geom,
                            const float* scale, const float* shift, const void* residual, float drop_p,
                            unsigned long long seed, const unsigned long long* seed_dev, void* stream);
/* Backward of that epilogue in one pass: dsum = dy * [y > 0] (residual case: y, dsum non-NULL; dsum is also the
 * residual's gradient), da = dsum * scale * [t > 0] (t = the stored pre-residual tensor, or y itself without residual;
 * scale = 1 / (1 - p)). */
int m3t_tcn_epilogue_bwd_bf16(const void* dy, const void* y, const void* t, void* dsum, void* da, float scale,
                              long long n, void* stream);
/* m3t_conv_fprop_bf16 (2-D, no epilogue arithmetic) whose output pixel (n, p, q) is stored at pixel offset
 * n*img_pitch + p*row_pitch + q*px_pitch (in units of Cout-element pixels) from y instead of densely: the data
 * gradient of a stride-2 convolution is evaluated as one small stride-1 convolution of dY per output parity
 * (h%2, w%2), each writing its quarter of dX in place - no zero-inserted copy of dY, no multiplications by zero.
 * accumulate != 0 adds to the stored values instead (the 1x1 downsample gradient joins the 3x3 one at the block input).
 * Replaces autograd's conv_backward_input for the stride-2 nn.Conv2d 3x3 / 1x1 of the first BasicBlock of
 * layer2..4 (models/resnet.py:7-15,24-27 via :98-101). */
int m3t_conv_fprop_scatter_bf16(const void* x, const void* w_packed, void* y, const int* geom, long long img_pitch,
                                long long row_pitch, long long px_pitch, int accumulate, int tile_hint, void* stream);

/* 3x3 / stride 1 / pad 1 convolution for Cin = Cout = 64 (ResNet layer1 fprop; its dgrad with the flipped filter):
 * persistent CTAs, the 9 filter taps resident in shared memory, ONE halo box per 128-position tile (zero padding by
 * TMA out-of-bounds fill), the taps read as row-shifted views of that box, double-buffered TMEM accumulators.
 * Same epilogue contract as m3t_conv_fprop_bf16.  x, y: bf16 [F][H][W][64]; w_packed: bf16 [64][9*64].
 * Replaces the same call sites as m3t_conv_fprop_bf16 for models/resnet.py layer1 (4 convs of 64->64 at 28x28). */
int m3t_conv3x3_c64_halo(const void* x, const void* w_packed, void* y, int F, int H, int W, const float* scale,
                         const float* shift, const void* residual, int relu, float* stats, void* stream);

/* 3x3 / stride 1 / pad 1 convolution for Cin = Cout = 128 with one whole image per tile (ResNet layer2 at 14x14; its
 * dgrad with the flipped filter): 128 <= H*(W+2) <= 256 and (H+2)*(W+2) <= 256, a multiple of 8.  The image's halo is
 * two 64-channel TMA boxes, the filter streams through a shared-memory ring as [128][64] blocks, taps are row-shifted
 * views of the boxes, N = 128 MMAs into double-buffered TMEM.  Epilogue contract and replaced call sites as
 * m3t_conv_fprop_bf16 (models/resnet.py layer2: 3 convs 128->128 at 14x14 and their data gradients).
 * x, y: bf16 [F][H][W][128]; w_packed: bf16 [128][9*128]. */
int m3t_conv3x3_c128_halo(const void* x, const void* w_packed, void* y, int F, int H, int W, const float* scale,
                          const float* shift, const void* residual, int relu, float* stats, void* stream);

/* Stem forward as a halo-tile kernel over the W-unrolled space-to-depth image xs [B][T][H2][W2][64] with the packed
 * (5,4,1)x64 filter [64][20*64]: per temporal tap one box of TR+3 rows plus that tap's filter slices, the 4 vertical
 * taps as row-shifted views; y bf16 [B*T][H2][W2][64]; epilogue contract as m3t_conv_fprop_bf16 (no residual).
 * Channels 48..63 of every pixel of xs are structural zeros of that layout and are NOT multiplied (3 K steps per tap).
 * Replaces nn.Conv3d(3,64,(5,7,7),(1,2,2),(2,3,3)) at models/backbone.py:328. */
int m3t_stem_fprop_halo(const void* xs, const void* w_packed, void* y, int B, int T, int H2, int W2,
                        const float* scale, const float* shift, int relu, float* stats, void* stream);

/* Halo-tile weight gradients for the 64-channel layers (wgrad_halo.cu): persistent CTAs, one activation halo box and
 * one dY box per K-block of whole image rows, filter taps as row-shifted MN-major views, accumulators resident in
 * TMEM, one atomic flush.  dw_packed is fp32, caller-zeroed, same packed layout as m3t_conv_wgrad_bf16.
 *   m3t_wgrad3x3_c64_halo : 3x3/s1/p1, 64->64, x/dy bf16 [F][H][W][64], dw_packed [64][9*64]   (ResNet layer1)
 *   m3t_wgrad_stem_halo   : stem over the W-unrolled s2d image xs [B][T][H2][W2][64], dy [B*T][H2][W2][64],
 *                           dw_packed [64][20*64] with tap = kt*4 + jh                         (models/backbone.py:328);
 *                           the activation box of an input frame stays in shared memory while the dY boxes of every
 *                           output frame that reads it stream through (two passes over the temporal taps) */
int m3t_wgrad3x3_c64_halo(const void* x, const void* dy, float* dw_packed, int F, int H, int W, void* stream);
int m3t_wgrad_stem_halo(const void* xs, const void* dy, float* dw_packed, int B, int T, int H2, int W2, void* stream);

/* Convolution weight gradient: dw_packed[Cout][taps*Cin] (fp32) += sum over output pixels of dy (x) patch(x).
 * The caller zero-fills dw_packed; split-K partial tiles are combined with fp32 atomics.
 * Replaces: the cuDNN wgrad autograd runs for every conv listed above (loss.backward(), models/model.py:146). */
int m3t_conv_wgrad_bf16(const void* x, const void* dy, float* dw_packed, const int* geom, int splits_hint,
                        void* stream);


/* ------------------------------------------------------------------------------------------------------------
 * HBM-bound passes of the visual stream (elementwise.cu).  Algorithmic bytes per element are given per entry; all
 * tensors are channels-last bf16 unless noted, C % 8 == 0.
 * ---------------------------------------------------------------------------------------------------------- */

/* Input pass of the stem: video (B,3,T,H,W) fp32 or uint8 -> out bf16 (B,T,H/2,W/2,16), 2x2 space-to-depth with
 * channel (ph*2+pw)*3+c (12..15 zero) and value x*mul+add.  mul=1/127.5, add=-1 reproduces
 * `(batch['video'] - 127.5) / 127.5` (models/model.py:106); the reference's transpose(1,2).contiguous()
 * (models/backbone.py:349-350) disappears because the layout is already frame-major.  Bytes/pixel: 12|3 in, 8 out. */
int m3t_video_prep_s2d(const void* video, int is_u8, void* out, int B, int T, int H, int W, float mul, float add,
                       void* stream);

/* Same, additionally unrolled over the stem's four horizontal taps: out bf16 (B,T,H/2,W/2,64) with
 * out[..][w2][jw*12+ch] = s2d[..][w2+jw-2][ch] for the 12 real channels ch = (ph*2+pw)*3+c (zero outside the image),
 * channels 48..63 = 0.  Every pixel is then one 128-byte row, and the stem
 * Conv3d(3,64,(5,7,7),s(1,2,2),p(2,3,3)) (models/backbone.py:328) becomes a (5,4,1) filter over 64 channels. */
int m3t_video_prep_s2d_w4(const void* video, int is_u8, void* out, int B, int T, int H, int W, float mul, float add,
                          void* stream);

/* Input pipeline on the device: the tensor m3t_video_prep_s2d_w4 produces, straight from DECODED uint8
 * frames [B][T][Hs][Ws][3] (HWC, channel order as decoded) with the reference's per-clip augmentations applied on the
 * fly: params int32 [B][8] = {crop_x, crop_y, flip, cut_y1, cut_y2, cut_x1, cut_x2, 0}; clip pixel (y,x) = frame pixel
 * (crop_y+y, crop_x+(flip ? W-1-x : x)); the cutout rectangle [cut_y1,cut_y2) x [cut_x1,cut_x2) of the clip is 127.5.
 * Replaces the crop / cv2.flip / stack / sequence_cutout steps of models/dataset.py:46-80 (load_video, no-resize case
 * crop_size == 112) and :16-31 (sequence_cutout), the float32 CTHW copy and the H2D transfer of 4-byte pixels. */
int m3t_video_augment_prep_s2d_w4(const void* frames_u8, const int* params, void* out, int B, int T, int Hs, int Ws,
                                  int H, int W, float mul, float add, void* stream);

/* Train-mode BatchNorm statistics -> per-channel scale/shift (+ running-stat update, momentum, unbiased variance).
 * stats = [2][C] column (sum, sum of squares) produced by the conv epilogue.  Replaces the statistics half of
 * nn.BatchNorm2d/3d in training (models/resnet.py:25,28; models/backbone.py:329). */
int m3t_bn_finalize(const float* stats, int C, double count, const float* gamma, const float* beta, float eps,
                    float momentum, float* running_mean, float* running_var, float* mean, float* invstd, float* scale,
                    float* shift, void* stream);
/* Eval-mode BatchNorm folded to scale/shift (optionally absorbing a conv bias). */
int m3t_bn_fold(int C, const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                const float* conv_bias, float eps, float* scale, float* shift, void* stream);
/* out = act(y*scale + shift + (res*res_scale + res_shift)) over [rows][C]: BN apply + residual add + ReLU of a
 * BasicBlock in ONE pass (models/resnet.py:41-54: bn -> (+identity) -> relu).  Bytes/element: 2 (+2) in, 2 out. */
int m3t_bn_act(const void* y, const float* scale, const float* shift, const void* res, const float* res_scale,
               const float* res_shift, int relu, void* out, long long rows, int C, void* stream);
/* Backward of m3t_bn_act with batch statistics, two passes: reduce (sum dz, sum dz*xhat; optionally materialises
 * dz for the residual branch) and apply (dy = scale*(dz - s0/n - xhat*s1/n)).  relu: 0 = none, 1 = dz = dout*(out>0)
 * from the stored activation, 2 = mask recomputed as (y*scale+shift > 0) (no residual: `out` may be NULL). */
int m3t_bn_bwd_reduce(const void* dout, const void* out, const void* y, const float* mean, const float* invstd,
                      const float* scale, const float* shift, int relu, void* dz_out, float* sums, long long rows,
                      int C, void* stream);
int m3t_bn_bwd_apply(const void* dout, const void* out, const void* y, const float* mean, const float* invstd,
                     const float* scale, const float* shift, const float* sums, double count, int relu, void* dy,
                     long long rows, int C, void* stream);
/* BN apply + ReLU + spatial max-pool in one pass: out = maxpool_{KxK, stride S, pad PAD}( relu(y*scale+shift) ) over
 * (H,W) of [F][H][W][C]; idx (uint8, optional) is the arg-max tap kh*K+kw.  Replaces BatchNorm3d apply + ReLU +
 * MaxPool3d((1,3,3),(1,2,2),(0,1,1)) of the ResNet stem (models/backbone.py:329-331) and + MaxPool3d((1,2,2))
 * of the VGG-M stacks (models/backbone.py:74-92,180-183,218-231). */
int m3t_bn_relu_maxpool(const void* y, const float* scale, const float* shift, void* out, void* idx, int F, int H,
                        int W, int C, int K, int S, int PAD, void* stream);
/* As m3t_bn_relu_maxpool, additionally writing ymax[f][p][q][c] = the RAW conv output y at the window's arg-max.  With
 * it the BatchNorm-backward sums of the unit (sum dz, sum dz*xhat) are a streaming pass over the POOLED tensors only —
 * m3t_bn_bwd_reduce(dout_pooled, NULL, ymax, ..., relu = 2) — instead of a pass over y and the indices (each window
 * sends its gradient to exactly one position, whose y is ymax): 0.8 GB instead of 2.3 GB for the stem at 4096 frames. */
int m3t_bn_relu_maxpool_ymax(const void* y, const float* scale, const float* shift, void* out, void* idx, void* ymax,
                             int F, int H, int W, int C, int K, int S, int PAD, void* stream);
/* Backward of the stem tail; mode 0 accumulates (sum dz, sum dz*xhat) into sums, mode 1 writes dy. */
int m3t_maxpool_bn_bwd(int mode, const void* dout, const void* idx, const void* y, const float* mean,
                       const float* invstd, const float* scale, const float* shift, float* sums, double count,
                       void* dy, int F, int H, int W, int C, int K, int S, int PAD, void* stream);
/* AdaptiveAvgPool2d(1)+flatten (models/resnet.py:117-119) over [F][HW][C] and its backward. */
int m3t_avgpool(const void* x, void* out_bf16, float* out_f32, int F, int HW, int C, void* stream);
int m3t_avgpool_bwd(const void* dout, int dout_f32, void* dx, int F, int HW, int C, void* stream);
/* Module-boundary layout changes: fp32 [N][C][S] <-> channels-last [N][S][Cpad] (bf16, or fp32 on the way back). */
int m3t_ncs_f32_to_nsc_bf16(const float* in, void* out, int N, int C, int S, int Cpad, void* stream);
int m3t_nsc_to_ncs_f32(const void* in, int in_f32, float* out, int N, int C, int S, int Cpad, void* stream);
/* Row-wise dtype casts with leading dimensions (pad columns of the bf16 side are zero-filled). */
int m3t_cast_f32_bf16(const float* in, long long ld_in, void* out, long long ld_out, long long rows, int cols,
                      void* stream);
int m3t_cast_bf16_f32(const void* in, long long ld_in, float* out, long long ld_out, long long rows, int cols,
                      void* stream);
/* Every convolution filter of a model re-packed by ONE launch after the optimizer step (the per-filter m3t_pack_filter
 * calls were 19 launches + 15 index_select per step).  `table_dev` = n entries in device memory, ordered by `start`
 * (prefix sum of the entries' tile counts, m3t_pack_entry_tiles); outputs as m3t_pack_filter, plus - when has_parity - the four parity sub-filters of
 * a stride-2 convolution's data gradient: tap t goes to par[par_of_tap[t]] at position pos_of_tap[t] of its
 * ntaps_par[.] taps, layout [Cin][ntaps][Cout] (par_of_tap[t] < 0: tap unused).  taps <= 27.  A matrix is a filter
 * with taps = 1 (wf = bf16 copy, wd = bf16 transpose: W_hh / W_hh^T of a GRU layer, Linear weights); has_parity < 0
 * marks a plain fp32 copy of Cout*Cin*taps values into wf (stacked bias vectors). */
typedef struct m3t_pack_entry {
  const void* src;      /* fp32 [Cout][Cin][taps] */
  void* wf;             /* bf16 [Cout][taps][Cin] or NULL */
  void* wd;             /* bf16 [Cin][taps flipped][Cout] or NULL */
  void* par[4];         /* bf16 parity sub-filters or NULL */
  long long start;      /* index of the entry's first tile (m3t_pack_entry_tiles) */
  int Cout, Cin, taps, has_parity;
  int ntaps_par[4];
  int par_of_tap[27];
  int pos_of_tap[27];
} m3t_pack_entry;
int m3t_pack_filters_batched(const m3t_pack_entry* table_dev, int n, long long total_tiles, void* stream);
/* Tiles (= thread blocks) one entry occupies: 32 output channels x min(64, 288 / taps) input channels x all taps per
 * tile, 2048 values per tile for a plain copy.  entry.start = sum of the tiles of the entries before it,
 * total_tiles = the sum over all entries. */
long long m3t_pack_entry_tiles(int Cout, int Cin, int taps, int has_parity);
/* Filter packing fp32 [Cout][Cin][taps] -> bf16 [Cout][taps][Cin] (fprop) and [Cin][taps reversed][Cout] (dgrad);
 * and the inverse re-layout of a packed fp32 weight gradient. */
int m3t_pack_filter(const float* w, void* w_fprop, void* w_dgrad, int Cout, int Cin, int taps, void* stream);
int m3t_unpack_filter_grad(const float* dw_packed, float* dw, int Cout, int Cin, int taps, void* stream);
/* Index-driven filter re-layout (space-to-depth stem): out[r][k] = idx[k] >= 0 ? w[r*row_stride+idx[k]] : 0, and
 * the scatter of a packed gradient back. */
int m3t_gather_pack_bf16(const float* w, const int* idx, void* out, int rows, long long row_stride, int K,
                         void* stream);
int m3t_scatter_unpack_f32(const float* dwp, const int* idx, float* dw, int rows, long long row_stride, int K,
                           void* stream);
/* Stride-2 dgrad helper: up[n][2p][2q][c] = dy[n][p][q][c], zeros elsewhere. */
int m3t_zero_insert2(const void* dy, void* up, int N, int P, int Q, int Hup, int Wup, int C, void* stream);
/* Small helpers: bf16 add (gradient fan-in), bias gradient (column sums), ReLU backward. */
int m3t_add_bf16(const void* a, const void* b, void* out, long long n, void* stream);
int m3t_colsum_bf16(const void* x, long long ld, long long rows, int cols, float* out, void* stream);
int m3t_relu_bwd_bf16(const void* dy, const void* out, void* dz, long long n, void* stream);
/* Inverted dropout of nn.Dropout(p) in the temporal blocks (models/tcn.py:23,29,35-36), bf16, n % 8 == 0:
 * y[i] = u_i >= floor(p * 2^32) ? x[i] / (1 - p) : 0 with u_i = top 32 bits of splitmix64(seed + (i+1) * 0x9E3779B97F4A7C15).
 * Counter-based: calling it on the gradient with the same seed applies the same mask (nothing is stored).  The mask
 * stream is this library's own (PyTorch's Philox stream is not reproduced); oracle/dropout.py restates it bit for bit. */
int m3t_dropout_bf16(const void* x, void* y, long long n, float p, unsigned long long seed, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Bidirectional GRU layer recurrence (gru.cu), persistent kernel, both directions in one launch.
 *   gi        fp32 [B*T][2][3H]  x . W_ih^T + b_ih (one m3t_gemm_bf16), gate order r,z,n
 *   w_hh_bf16 bf16 [2][3H][H];  b_hh fp32 [2][3H]
 *   out_bf16  bf16 [B][T][2H] (forward | reverse halves);  out_f32 optional fp32 copy
 *   saved     fp32 [B*T][2][4][H] (r, z, n, W_hn h + b_hn) or NULL (inference)
 *   counters  uint32 scratch, >= 2*ceil(B/32) entries (zeroed by the call)
 * Replaces nn.GRU(batch_first=True, bidirectional=True) (models/rnn.py:17,72-75) = cuDNN RNN in the reference. */
int m3t_gru_fwd(const float* gi, const void* w_hh_bf16, const float* b_hh, void* out_bf16, float* out_f32,
                float* saved, unsigned* counters, int B, int T, int H, void* stream);
/* Small-batch variant of m3t_gru_fwd (B <= 64 always, up to the number of 16-row clusters the device keeps resident
 * beyond that - returns -3 when they would not all be resident; H in {128, 256, 512}; `saved` as in m3t_gru_fwd or NULL): per direction and group
 * of 16 batch rows one thread-block cluster (H/64 CTAs), W_hh stays in shared memory and the hidden state is exchanged over
 * distributed shared memory behind one hardware cluster barrier per step instead of through L2 + an arrival counter.
 * Same operands, k order and gate arithmetic as m3t_gru_fwd.  Returns -1 for shapes it does not take. */
int m3t_gru_fwd_cluster(const float* gi, const void* w_hh_bf16, const float* b_hh, void* out_bf16, float* out_f32,
                        float* saved, int B, int T, int H, void* stream);
/* The bf16 weight copies of one bidirectional layer in one launch (rebuilt after every optimizer step):
 * wih bf16 [6H][Ipad] = [W_ih ; W_ih_reverse] zero-padded to Ipad columns, whh bf16 [2][3H][H], whht bf16 [2][H][3H]
 * (W_hh transposed, for m3t_gru_bwd; may be NULL); bias fp32 [2][6H] = [b_ih ; b_ih_reverse], [b_hh ; b_hh_reverse]
 * (optional, with the four bias vectors).  Inputs are the nn.GRU parameters (models/rnn.py:17). */
int m3t_gru_pack_weights(const float* w_ih, const float* w_ih_r, const float* w_hh, const float* w_hh_r, void* wih,
                         void* whh, void* whht, int I, int Ipad, int H, const float* b_ih, const float* b_ih_r,
                         const float* b_hh, const float* b_hh_r, float* bias, void* stream);
/* BPTT: dgi, dgh bf16 [B*T][2][3H] (gradients wrt the input / hidden pre-activations) and hprev bf16 [B*T][2][H]
 * (h_{t-1}, zero at the sequence start); w_hh_t_bf16 = bf16 [2][H][3H] (W_hh transposed).  dbias (optional, caller-
 * zeroed) fp32 [2][2][3H] += the bias gradients (b_ih | b_hh) x direction = column sums of the fp32 gate gradients over
 * (b, t).  The weight / input gradients follow as GEMMs over dgi / dgh. */
int m3t_gru_bwd(const void* dout_bf16, const void* out_bf16, const float* saved, const void* w_hh_t_bf16,
                void* dgi_bf16, void* dgh_bf16, void* hprev_bf16, unsigned* counters, float* dbias, int B, int T, int H,
                void* stream);
/* BPTT on thread-block clusters of H/32 CTAs (H in {128, 256, 512}; 16-CTA clusters for H = 512): each CTA multiplies
 * the gate gradients of ITS 32 hidden units (just computed, in its own shared memory) with its K-slice of W_hh^T and
 * the partial products are reduce-scattered over distributed shared memory behind one hardware cluster barrier per
 * step, added in rank order (deterministic).  Same operands and results as m3t_gru_bwd up to the fp32 summation
 * order of the recurrent product; no arrival counters.  Returns -23 when the device cannot co-schedule such a cluster
 * and -3 / -1 for batches / shapes it does not take: callers fall back to m3t_gru_bwd. */
int m3t_gru_bwd_cluster(const void* dout_bf16, const void* out_bf16, const float* saved, const void* w_hh_t_bf16,
                        void* dgi_bf16, void* dgh_bf16, void* hprev_bf16, float* dbias, int B, int T, int H,
                        void* stream);
/* Clusters of that kernel (hidden size H) the device keeps resident at once; <= 0: it cannot run here. */
int m3t_gru_bwd_cluster_max(int H);

/* ------------------------------------------------------------------------------------------------------------
 * Attention-fusion mix (fusion.cu): f = softmax(sigmoid(s_v), sigmoid(s_a)) . (x_v, x_a) per (b,t) row of C
 * channels, and its backward.  Replaces sigmoid/cat/softmax/mul/add (models/att_fusion.py:21-25, six ATen kernels). */
int m3t_att_mix_fwd(const void* x_a, const void* x_v, const float* s_a, const float* s_v, void* f, float* w_v,
                    long long rows, int C, void* stream);
int m3t_att_mix_bwd(const void* df, const void* x_a, const void* x_v, const float* s_a, const float* s_v, void* dx_a,
                    void* dx_v, float* ds_a, float* ds_v, long long rows, int C, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused log-Mel front end (melspec.cu).  wav: fp32 mono 16 kHz [n_samples]; out_db: fp32 [1 + n_samples/hop][n_mels].
 * Framing (centre padding n_fft/2, pad_mode 0 = zeros / 1 = reflect), periodic Hann(win_length) zero-padded to
 * n_fft = 512, FFT, power, mel projection with the caller's filterbank mel_fb [n_mels][257], 10*log10(max(S,1e-10)),
 * clip to (global max - top_db) when top_db > 0.  scratch: one int.  Replaces librosa.feature.melspectrogram +
 * librosa.power_to_db at process/extract_melspec.py:15-19 (CPU, offline in the reference). */
int m3t_logmel(const float* wav, long long n_samples, int hop, int win_length, int n_mels, const float* mel_fb,
               int pad_mode, float top_db, float* out_db, int* scratch, void* stream);
/* 200-d stacked audio features: out[t][j*n_mels+m] = mel[3*(start+t)+j][m] for j < 5, zero past the end
 * (models/dataset.py:83-95 `load_audio`). */
int m3t_mel_stack(const float* mel, long long n_frames, int n_mels, long long start, int w_len, float* out,
                  void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Optimiser pass over the flat fp32 parameter / gradient arena (optim.cu).
 * m3t_sumsq_f32: out[0] = sum g^2, computed in two passes with a fixed summation order (block partials in
 * `workspace`, m3t_sumsq_workspace_floats() floats, then one block): the result is bit-reproducible, so data-parallel
 * replicas derive identical clip coefficients from their all-reduced gradients.  m3t_adam_clip_step: g' = g*grad_scale*min(1, max_norm /
 * (sqrt(gnorm_sq)*grad_scale + 1e-6)) (skipped when max_norm <= 0), then torch.optim.Adam semantics with coupled
 * weight decay.  Replaces clip_grad_norm_(1.0) (train.py:35 via Lightning) + Adam(lr, weight_decay=1e-4)
 * (models/model.py:388-390); grad_scale = 1/world_size folds the data-parallel mean. */
long long m3t_sumsq_workspace_floats(void);
int m3t_sumsq_f32(const float* g, long long n, float* out, float* workspace, void* stream);
int m3t_adam_clip_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                       float eps, float weight_decay, int step, float max_norm, float grad_scale,
                       const float* gnorm_sq, void* stream);
/* The same step with the optimiser scalars resident on the device (CUDA-graph replays of the whole training step hold
 * no host value that changes between steps): hyper = float[8] {lr, weight_decay, step, bc1, sqrt(bc2), -, -, -}; the
 * call first advances hyper[2] by one and derives the two bias corrections on the device, then applies clip + Adam.
 * The host rewrites hyper[0..1] when a scheduler changes them (models/model.py:393-407) and hyper[2] when a
 * checkpoint restores the step count. */
int m3t_adam_clip_step_dev(float* p, const float* g, float* m, float* v, long long n, float* hyper, float beta1,
                           float beta2, float eps, float max_norm, float grad_scale, const float* gnorm_sq,
                           void* stream);

/* 3x3/pad-1 patches of a 1-channel fp32 image [N][H][W] as bf16 GEMM rows [N*H*W][16] (9 taps + 7 zero columns):
 * the 1->64 channel stem of the builder-declared audio ResNet composition (BASELINE config 2; no reference symbol,
 * SURVEY F6) then is one m3t_gemm_bf16 with K = 16. */
int m3t_patch3x3_c1(const float* x, void* out, int N, int H, int W, void* stream);

/* Hardware probe (debug.cu), not on the product path: out[128][64] = A[shift_rows : shift_rows+128] . B^T with the
 * K-major SWIZZLE_128B A operand starting at an arbitrary 128-byte row of a TMA-written [256][64] tile;
 * mode 1 additionally sets the descriptor base_offset to (addr >> 7) & 7. */
int m3t_debug_rowshift(const void* A, const void* B, float* out, int shift_rows, int mode, void* stream);

/* ---- evaluation post-processing on the device (csrc/postproc.cu) --------------------------------
 * All videos at once as ragged sequences: seq_off[v]..seq_off[v+1] (V+1 int64, device) are the frames of video v in
 * one flat [total_frames][C] array. */

/* Overlap-add of window predictions into per-video frame tracks, then halving of frames >= window/2 of every video
 * (the half-stride windows cover them twice).  pred f32 [S][L][C]; seg_start/seg_len int32 [S] (start frame inside
 * the video, valid frames of the segment); seg_base int64 [S] = seq_off[video of segment]; out f32
 * [total_frames][C] (zeroed here).  Replaces the Python loops of models/model.py:281-297 (validation_end) and
 * :358-366 (test_end).  With at most two windows per frame the fp32 sums are order-independent (bit-exact). */
int m3t_overlap_add_f32(const float* pred, const int* seg_start, const int* seg_len, const long long* seg_base,
                        const long long* seq_off, float* out, long long S, int L, int C, int V,
                        long long total_frames, int window, void* stream);

/* scipy.signal.wiener(x, window) along the frames of every (video, channel) sequence, float64 arithmetic on the
 * float32 tracks exactly as scipy promotes them (x**2 in float32, sums and the filter in float64).
 * lmean, lvar: f64 [total_frames][C] workspaces; noise_sum: f64 [V][C] workspace; out f64 [total_frames][C];
 * max_len = longest sequence.  Replaces models/utils.py:29-33 smooth_predictions(mode='wiener') as called with
 * window 35 by get_smoothed_ccc.py:15-16. */
int m3t_wiener1d_f64(const float* x, const long long* seq_off, int V, int C, int window, long long max_len,
                     double* lmean, double* lvar, double* noise_sum, double* out, void* stream);

/* Moments for the masked concordance correlation: moments f64 [V][C][6] = {n, sum a, sum b, sum a^2, sum b^2,
 * sum ab} over the frames whose ground truth is >= -1 in EVERY channel (a = pred f64, b = gt f32).  Per-video CCC
 * and, by adding the moments over videos, the global CCC follow in closed form.  Replaces
 * models/utils.py:19-21 concordance_cc2_np + the mask / concatenation of get_smoothed_ccc.py:17-30. */
int m3t_ccc_moments_f64(const double* pred, const float* gt, const long long* seq_off, int V, int C, double* moments,
                        void* stream);

/* ---- fp32-parity inference mode (csrc/fp32mode.cu) ----------------------------------------------------------------
 * North-star tolerance "fp32 max error <= 1e-4 on the per-frame V/A predictions".  Activations stay float32; the
 * tensor-core launches are the SAME bf16 kernels (m3t_gemm_bf16 / m3t_conv_fprop_bf16 with fp32 output) over 3x the
 * contraction: x = hi + lo with hi = bf16(x), lo = bf16(x - hi), and a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi is
 * obtained by concatenating A' = [a_hi | a_hi | a_lo], B' = [b_hi | b_lo | b_hi] along K (per channel block); bf16
 * products are exact in fp32 and accumulate in the fp32 TMEM accumulator.  Forward / eval only. */

/* y = x (+ res) (ReLU);  out_f32 = y (optional, may alias x);  out3 bf16 [rows][3C] = [hi | hi | lo] of y.
 * Replaces, in this mode, the residual add + ReLU of BasicBlock.forward (models/resnet.py:52-54) and feeds the next
 * convolution / Linear. */
int m3t_split3_bf16(const float* x, const float* res, int relu, float* out_f32, void* out3, long long rows, int C,
                    int nterms, void* stream);
/* nterms = 3 (above, error 2^-16) or 6: three pieces per value, x = p0 + p1 + p2, and the products a0b0 + a0b1 + a1b0 +
 * a0b2 + a1b1 + a2b0 (error 2^-24): activations [a0 a0 a1 a0 a1 a2], weights [b0 b1 b0 b2 b1 b0], 6x the contraction. */
/* Weights: f32 [N][G][C] (tap_minor 0) or [N][C][G] (tap_minor 1, nn.Conv layout, G = taps) -> bf16 [N][G][3C] =
 * [hi | lo | hi] per group (nn.Linear: G = 1). */
int m3t_pack_split3_bf16(const float* w, void* out, long long N, int G, int C, int tap_minor, int cpad, int nterms,
                         void* stream);
/* cpad (0 = nterms*C): elements per group in `out`; a tail beyond nterms*C is zero-filled (first VGG-M conv: 3*16 -> 64).
 * m3t_video_prep_s2d with split output for that conv: bf16 (B,T,H/2,W/2,cpad) = nterms blocks of 16 channels + zeros. */
int m3t_video_prep_s2d_split3(const void* video, int is_u8, void* out, int B, int T, int H, int W, float mul, float add,
                              int nterms, int cpad, void* stream);
/* nn.MaxPool3d((1,2,2),(1,2,2)) of the VGG-M groups (models/backbone.py:76-96) on float32 channels-last tensors. */
int m3t_maxpool2x2_f32(const float* x, float* out, int F, int H, int W, int C, void* stream);
/* m3t_video_prep_s2d_w4 with split output: bf16 (B,T,H/2,W/2,64*nterms), nterms blocks of 64 channels per pixel. */
int m3t_video_prep_s2d_w4_split3(const void* video, int is_u8, void* out, int B, int T, int H, int W, float mul,
                                 float add, int nterms, void* stream);
/* nn.MaxPool3d((1,3,3),(1,2,2),(0,1,1)) (models/backbone.py:331) / AdaptiveAvgPool2d(1) (models/resnet.py:117) on
 * float32 channels-last tensors. */
int m3t_maxpool3s2_f32(const float* x, float* out, int F, int H, int W, int C, void* stream);
int m3t_avgpool_f32(const float* x, float* out, int F, int HW, int C, void* stream);
/* models/att_fusion.py:21-25 in float32: f = softmax(sigmoid(s_v), sigmoid(s_a)) . (x_v, x_a). */
int m3t_att_mix_f32(const float* x_a, const float* x_v, const float* s_a, const float* s_v, float* f, long long rows,
                    int C, void* stream);
/* Bidirectional GRU layer recurrence in float32 FFMA (one launch per time step; gi f32 [B*T][2][3H] from the split
 * GEMM, w_hh f32 [2][3H][H], b_hh f32 [2][3H], out f32 [B][T][2H]).  Replaces nn.GRU (models/rnn.py:17,72-75). */
int m3t_gru_fwd_f32(const float* gi, const float* w_hh, const float* b_hh, float* out, int B, int T, int H,
                    void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * CBAM gates (cbam.cu) on channels-last bf16 feature maps x[F][S][C], S = H*W.  Replace the ATen mean / max / mul /
 * expand_as chains of models/cbam.py:46-53 (ChannelGate), :62-66,82-87 (ChannelPool + SpatialGate) and its cuDNN
 * Conv2d(2,1,5,padding=2) (:79).  The shared MLP on (F,C) and BatchNorm2d(1) on (F,1,H,W) stay on the host side.
 *   pool_hw:       avg / max (first maximum) over S per (f,c);   _bwd: dx = davg/S + [s == arg] dmx
 *   scale_c:       y = x * sc[f][c];                              _bwd: dx = dy * sc, dsc[f][c] = sum_s dy * x
 *   pool_c:        comp[f][0][s] = max_c, comp[f][1][s] = mean_c; _bwd: dx = dcomp[1]/C + [c == carg] dcomp[0]
 *   scale_s:       y = x * ss[f][s];                              _bwd: dx = dy * ss, dss[f][s] = sum_c dy * x
 *   conv5:         out[f][y][x] = sum w[c][kh][kw] in[f][c][y+kh-2][x+kw-2] (fp32);  _bwd: din and dw[2][5][5] */
int m3t_cbam_pool_hw(const void* x, float* avg, float* mx, int* arg, int F, int S, int C, void* stream);
int m3t_cbam_pool_hw_bwd(const float* davg, const float* dmx, const int* arg, void* dx, int F, int S, int C,
                         void* stream);
int m3t_cbam_scale_c(const void* x, const float* sc, void* y, int F, int S, int C, void* stream);
int m3t_cbam_scale_c_bwd(const void* dy, const void* x, const float* sc, void* dx, float* dsc, int F, int S, int C,
                         void* stream);
int m3t_cbam_pool_c(const void* x, float* comp, int* carg, int F, int S, int C, void* stream);
int m3t_cbam_pool_c_bwd(const float* dcomp, const int* carg, void* dx, int F, int S, int C, void* stream);
int m3t_cbam_scale_s(const void* x, const float* ss, void* y, long long rows, int C, void* stream);
int m3t_cbam_scale_s_bwd(const void* dy, const void* x, const float* ss, void* dx, float* dss, long long rows, int C,
                         void* stream);
int m3t_cbam_conv5(const float* in, const float* w, float* out, int F, int H, int W, void* stream);
int m3t_cbam_conv5_bwd(const float* dout, const float* in, const float* w, float* din, float* dw, int F, int H, int W,
                       void* stream);

/* AvgPool3d((1,2,2), stride (1,2,2)) of the DenseNet transitions (models/densenet.py:38) on channels-last bf16
 * [F][H][W][C] -> [F][H/2][W/2][C] (floor), and its backward (dx = dy/4 inside the windows, 0 on an odd last row / column). */
int m3t_avgpool2x2(const void* x, void* y, int F, int H, int W, int C, void* stream);
int m3t_avgpool2x2_bwd(const void* dy, void* dx, int F, int H, int W, int C, void* stream);

/* Training loss of the task module and its gradient in one launch (fusion.cu): L = lambda (1 - CCC(v_hat, v)) +
 * (1 - lambda)(1 - CCC(a_hat, a)) [+ w_ce * mean_i(valid_i * CE(logits_i, class_i)) when n_logits > 0], CCC with the
 * reference's mixed estimators (biased covariance, unbiased variances; models/utils.py:6-17).  y_hat fp32 [N][C],
 * idx_v / idx_a = columns of the valence / arousal predictions, n_logits = leading expression logits (0: none).
 * out4 = {L, L_v, L_a, CE mean}; dy = dL/dy_hat [N][C].  One CTA, fixed-order reductions (bit-reproducible).  Replaces
 * ~90 ATen launches of concordance_cc2 / cross_entropy forward + backward (models/model.py:132-144,162-182). */
int m3t_av_loss(const float* y_hat, const float* label_v, const float* label_a, const long long* cls,
                const unsigned char* valid, int N, int C, int idx_v, int idx_a, int n_logits, float lambda, float w_ce,
                float* out4, float* dy, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Deterministic training (the reference asks cuDNN for deterministic algorithms, train.py:17).  Every fp32 atomic
 * accumulation of the training step has a slotted form: the destination passed by the caller is the FIRST of
 * (1 + nslots) consecutive copies (all zero-filled by the caller); CTA / warp / split s accumulates into copy 1 + s,
 * where it is the only writer, and m3t_det_reduce(buf, len, nslots) adds the copies to copy 0 in index order.
 *   BatchNorm statistics of the conv kernels   m3t_conv_fprop_bf16 with tile_hint bit 8 (forces the persistent kernel);
 *                                              m3t_conv3x3_c64_halo / _c128_halo / m3t_stem_fprop_halo with relu bit 8
 *                                              nslots = m3t_det_stats_slots(), len = 2 * Cout
 *   split-K weight gradients                   m3t_conv_wgrad_bf16 with splits_hint bit 29,
 *                                              nslots = m3t_conv_wgrad_splits(geom, splits_hint), len = Cout*taps*Cin
 *   halo-tile weight gradients                 m3t_wgrad3x3_c64_halo_det / m3t_wgrad_stem_halo_det,
 *                                              nslots = m3t_det_cta_slots(), len = 64*576 / 64*1280
 *   BatchNorm-backward sums                    m3t_bn_bwd_reduce with relu bit 8 (fixed-order block reduction),
 *                                              nslots = m3t_det_stats_slots(), len = 2 * C
 *   bias-gradient column sums                  m3t_colsum_bf16 with cols bit 30 (one row block per column strip; no slots)
 * The GRU bias gradients are taken as deterministic column sums of dgi / dgh (dbias = NULL in m3t_gru_bwd); the loss,
 * the gradient norm, clip and Adam are fixed-order already. */
int m3t_det_stats_slots(void);
int m3t_det_cta_slots(void);
int m3t_det_reduce(float* buf, long long len, int nslots, void* stream);
int m3t_conv_wgrad_splits(const int* geom, int splits_hint);
int m3t_wgrad3x3_c64_halo_det(const void* x, const void* dy, float* dw_packed, int F, int H, int W, void* stream);
int m3t_wgrad_stem_halo_det(const void* xs, const void* dy, float* dw_packed, int B, int T, int H2, int W2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* M3T_B200_H_ */
