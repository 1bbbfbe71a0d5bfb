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
 *   tile_hint: 0 = auto; bit0 force 128-row tiles, bit1 force 256-row tiles, bit2 allow 256-column tiles.
 * Replaces: nn.Conv2d 3x3 / 1x1 in BasicBlock (models/resnet.py:7-15,24-27,40-54,98-101), nn.Conv3d 3x3x3 in
 * VA_3DVGGM(_Split) (models/backbone.py:73-103,179-195,243-271), weight-normed dilated causal nn.Conv1d in
 * TemporalBlock (models/tcn.py:19-33) and Conv1d k5 in tcn_simple (models/backbone.py:214-231), with the
 * BatchNorm / residual / ReLU that follow them folded into the epilogue (eval) or their statistics (train). */
int m3t_conv_fprop_bf16(const void* x, const void* w_packed, void* y, const int* geom, const float* scale,
                        const float* shift, const void* residual, int relu, float* stats, int tile_hint,
                        void* stream);

/* Convolution weight gradient: dw_packed[Cout][taps*Cin] (fp32) += sum over output pixels of dy (x) patch(x).
 * The caller zero-fills dw_packed; split-K partial tiles are combined with fp32 atomics.
 * Replaces: the cuDNN wgrad autograd runs for every conv listed above (loss.backward(), models/model.py:146). */
int m3t_conv_wgrad_bf16(const void* x, const void* dy, float* dw_packed, const int* geom, int splits_hint,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* M3T_B200_H_ */
