"""`torch.library` registration of the C-ABI kernels: `torch.ops.m3t.*`.

The north star asks for the kernels to be callable "as torch custom ops and autograd.Functions through a thin C-ABI
layer".  The `torch.autograd.Function`s of ops.py are what the modules call (they carry Python-side context such as
packed-weight caches, and their dispatch costs a few microseconds less per launch, which matters in the launch-bound
regimes).  This module registers the same entry points with the dispatcher, so that they are visible to
`torch.ops`, carry schemas and fake (meta) implementations — FakeTensor / `torch.export` tracing, `opcheck` — and, for
the differentiable ones, `register_autograd` formulas built from the same raw calls:

    torch.ops.m3t.gemm            D = act((A . B^T) * scale + shift + residual)            m3t_gemm_bf16
    torch.ops.m3t.conv_fprop      implicit-GEMM convolution (1-3 D) with the fused epilogue  m3t_conv_fprop_bf16 (+ halo routes)
    torch.ops.m3t.conv_wgrad      packed fp32 weight gradient                                 m3t_conv_wgrad_bf16 (+ halo routes)
    torch.ops.m3t.bn_act          y * scale + shift (+ residual) (ReLU)                       m3t_bn_act
    torch.ops.m3t.linear          act(x W^T + b)                     [autograd]               m3t_gemm_bf16 x3 + m3t_colsum
    torch.ops.m3t.att_mix         softmax(sigmoid(s_v), sigmoid(s_a)) . (x_v, x_a) [autograd] m3t_att_mix_fwd / _bwd
    torch.ops.m3t.gru_layer       one bidirectional GRU layer        [autograd]               m3t_gemm_bf16 + m3t_gru_fwd / _bwd
    torch.ops.m3t.logmel          log-Mel spectrogram in dB                                   m3t_logmel

Importing this module performs the registration (idempotent).  `M3T_TORCH_OPS=1` makes ops.linear / ops.AttMixFn's
callers go through the dispatcher instead of the autograd.Functions (tests/gpu_cases.py::case_torch_library compares
the two paths bit for bit).
"""
from typing import List, Optional

import torch

from . import ops, raw

_LIB = "m3t"
_registered = False


def _out_dtype(f32):
    return torch.float32 if f32 else torch.bfloat16


def register():
    global _registered
    if _registered:
        return
    _registered = True
    custom_op = torch.library.custom_op

    # ---------------------------------------------------------------- GEMM
    @custom_op(_LIB + "::gemm", mutates_args=())
    def gemm(A: torch.Tensor, B: torch.Tensor, a_mn: bool, b_mn: bool, out_f32: bool,
             scale: Optional[torch.Tensor], shift: Optional[torch.Tensor], residual: Optional[torch.Tensor],
             relu: bool) -> torch.Tensor:
        return raw.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out_dtype=_out_dtype(out_f32), scale=scale, shift=shift,
                        residual=residual, relu=relu)

    @gemm.register_fake
    def _(A, B, a_mn, b_mn, out_f32, scale, shift, residual, relu):
        M = A.shape[1] if a_mn else A.shape[0]
        N = B.shape[1] if b_mn else B.shape[0]
        return A.new_empty((M, N), dtype=_out_dtype(out_f32))

    # ---------------------------------------------------------------- convolution
    @custom_op(_LIB + "::conv_fprop", mutates_args=())
    def conv_fprop(x: torch.Tensor, w_packed: torch.Tensor, geom: List[int], scale: Optional[torch.Tensor],
                   shift: Optional[torch.Tensor], residual: Optional[torch.Tensor], relu: bool) -> torch.Tensor:
        return raw.conv_fprop(x, w_packed, list(geom), scale=scale, shift=shift, residual=residual, relu=relu)

    @conv_fprop.register_fake
    def _(x, w_packed, geom, scale, shift, residual, relu):
        Z, P, Q = raw.conv_out_dims(list(geom))
        return x.new_empty((geom[1], Z, P, Q, geom[6]), dtype=torch.bfloat16)

    @custom_op(_LIB + "::conv_wgrad", mutates_args=())
    def conv_wgrad(x: torch.Tensor, dy: torch.Tensor, geom: List[int]) -> torch.Tensor:
        return raw.conv_wgrad(x, dy, list(geom)).clone()

    @conv_wgrad.register_fake
    def _(x, dy, geom):
        return x.new_empty((geom[6], geom[7] * geom[8] * geom[9] * geom[5]), dtype=torch.float32)

    @custom_op(_LIB + "::bn_act", mutates_args=())
    def bn_act(y: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, residual: Optional[torch.Tensor],
               relu: bool) -> torch.Tensor:
        return raw.bn_act(y, scale, shift, res=residual, relu=relu)

    @bn_act.register_fake
    def _(y, scale, shift, residual, relu):
        return torch.empty_like(y)

    # ---------------------------------------------------------------- Linear (differentiable)
    @custom_op(_LIB + "::linear", mutates_args=())
    def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], relu: bool, out_f32: bool) -> torch.Tensor:
        with torch.no_grad():
            return ops.LinearFn.apply(ops.as_bf16(x), w, b, relu, out_f32)

    @linear.register_fake
    def _(x, w, b, relu, out_f32):
        return x.new_empty(tuple(x.shape[:-1]) + (w.shape[0],), dtype=_out_dtype(out_f32))

    @custom_op(_LIB + "::linear_bwd", mutates_args=())
    def linear_bwd(dout: torch.Tensor, x: torch.Tensor, w: torch.Tensor, out: torch.Tensor, relu: bool,
                   has_bias: bool) -> List[torch.Tensor]:
        N, K = w.shape
        x2 = ops.as_bf16(x).contiguous().view(-1, x.shape[-1])
        M = x2.shape[0]
        d2 = dout.contiguous().view(M, N)
        npad = (N + 7) // 8 * 8
        if d2.dtype == torch.float32:
            d2 = raw.cast_bf16(d2, npad)
        elif npad != N:
            d2 = raw.cast_bf16(raw.cast_f32(d2), npad)
        if relu:
            o2 = out.contiguous().view(M, N)
            d2 = raw.relu_bwd(d2, o2 if o2.dtype == torch.bfloat16 else raw.cast_bf16(o2))
        dd = d2[:, :N] if npad != N else d2
        wb = raw.cast_bf16(w.detach().view(N, -1))
        dx = raw.gemm(dd, wb, b_mn=True, out_dtype=torch.bfloat16).view(x.shape)
        dw = raw.gemm(dd, x2, a_mn=True, b_mn=True, out_dtype=torch.float32)
        db = raw.colsum(d2, N).clone() if has_bias else dw.new_zeros((N,))
        return [dx if x.dtype == torch.bfloat16 else raw.cast_f32(dx.view(M, K)).view(x.shape), dw, db]

    @linear_bwd.register_fake
    def _(dout, x, w, out, relu, has_bias):
        return [torch.empty_like(x), w.new_empty(w.shape, dtype=torch.float32), w.new_empty((w.shape[0],),
                                                                                            dtype=torch.float32)]

    def _linear_setup(ctx, inputs, output):
        x, w, b, relu, _ = inputs
        ctx.save_for_backward(x, w, output)
        ctx.relu, ctx.has_bias = relu, b is not None

    def _linear_backward(ctx, dout):
        x, w, out = ctx.saved_tensors
        dx, dw, db = torch.ops.m3t.linear_bwd(dout, x, w, out, ctx.relu, ctx.has_bias)
        return dx, dw, (db if ctx.has_bias else None), None, None

    linear.register_autograd(_linear_backward, setup_context=_linear_setup)

    # ---------------------------------------------------------------- attention mix (differentiable)
    @custom_op(_LIB + "::att_mix", mutates_args=())
    def att_mix(x_a: torch.Tensor, x_v: torch.Tensor, s_a: torch.Tensor, s_v: torch.Tensor) -> torch.Tensor:
        return raw.att_mix_fwd(x_a.contiguous(), x_v.contiguous(), s_a.contiguous().float(), s_v.contiguous().float())

    @att_mix.register_fake
    def _(x_a, x_v, s_a, s_v):
        return torch.empty_like(x_a)

    @custom_op(_LIB + "::att_mix_bwd", mutates_args=())
    def att_mix_bwd(df: torch.Tensor, x_a: torch.Tensor, x_v: torch.Tensor, s_a: torch.Tensor,
                    s_v: torch.Tensor) -> List[torch.Tensor]:
        return list(raw.att_mix_bwd(df.contiguous(), x_a.contiguous(), x_v.contiguous(), s_a.contiguous().float(),
                                    s_v.contiguous().float()))

    @att_mix_bwd.register_fake
    def _(df, x_a, x_v, s_a, s_v):
        return [torch.empty_like(x_a), torch.empty_like(x_v), s_a.new_empty(s_a.shape, dtype=torch.float32),
                s_v.new_empty(s_v.shape, dtype=torch.float32)]

    def _att_setup(ctx, inputs, output):
        ctx.save_for_backward(*inputs)

    def _att_backward(ctx, df):
        return tuple(torch.ops.m3t.att_mix_bwd(df, *ctx.saved_tensors))

    att_mix.register_autograd(_att_backward, setup_context=_att_setup)

    # ---------------------------------------------------------------- one bidirectional GRU layer (differentiable)
    @custom_op(_LIB + "::gru_layer", mutates_args=())
    def gru_layer(x: torch.Tensor, w_ih: torch.Tensor, w_hh: torch.Tensor, b_ih: torch.Tensor, b_hh: torch.Tensor,
                  w_ih_r: torch.Tensor, w_hh_r: torch.Tensor, b_ih_r: torch.Tensor,
                  b_hh_r: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            return ops.GRULayerFn.apply(ops.as_bf16(x), w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r, False)

    @gru_layer.register_fake
    def _(x, w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r):
        return x.new_empty((x.shape[0], x.shape[1], 2 * w_hh.shape[1]), dtype=torch.bfloat16)

    @custom_op(_LIB + "::gru_layer_bwd", mutates_args=())
    def gru_layer_bwd(dout: torch.Tensor, x: torch.Tensor, w_ih: torch.Tensor, w_hh: torch.Tensor,
                      b_ih: torch.Tensor, b_hh: torch.Tensor, w_ih_r: torch.Tensor, w_hh_r: torch.Tensor,
                      b_ih_r: torch.Tensor, b_hh_r: torch.Tensor) -> List[torch.Tensor]:
        # recompute-forward backward: the dispatcher-level op keeps no Python context, so the saved gates are rebuilt
        with torch.enable_grad():
            leaves = [t.detach().requires_grad_(True) for t in (x if x.dtype == torch.bfloat16 else x.float(), w_ih,
                                                                 w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r)]
            out = ops.GRULayerFn.apply(ops.as_bf16(leaves[0]), *leaves[1:], True)
            grads = torch.autograd.grad(out, leaves, dout.to(out.dtype))
        return [g.clone() for g in grads]

    @gru_layer_bwd.register_fake
    def _(dout, x, *params):
        return [torch.empty_like(x)] + [p.new_empty(p.shape, dtype=torch.float32) for p in params]

    def _gru_setup(ctx, inputs, output):
        ctx.save_for_backward(*inputs)

    def _gru_backward(ctx, dout):
        return tuple(torch.ops.m3t.gru_layer_bwd(dout, *ctx.saved_tensors))

    gru_layer.register_autograd(_gru_backward, setup_context=_gru_setup)

    # ---------------------------------------------------------------- log-Mel
    @custom_op(_LIB + "::logmel", mutates_args=())
    def logmel(wav: torch.Tensor, fps: float, reflect_pad: bool, top_db: float) -> torch.Tensor:
        from .process.extract_melspec import melspectrogram_db
        return melspectrogram_db(wav, fps, pad_mode="reflect" if reflect_pad else "constant", top_db=top_db)

    @logmel.register_fake
    def _(wav, fps, reflect_pad, top_db):
        hop = int(1 / 3 * 1 / fps * 16000)
        return wav.new_empty((1 + wav.numel() // hop, 40), dtype=torch.float32)


OP_NAMES = ("gemm", "conv_fprop", "conv_wgrad", "bn_act", "linear", "linear_bwd", "att_mix", "att_mix_bwd", "gru_layer",
            "gru_layer_bwd", "logmel")

register()
