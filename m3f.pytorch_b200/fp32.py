"""fp32-parity inference mode (north-star tolerance: max error <= 1e-4 on the per-frame V/A predictions).

    with m3t_b200.fp32.parity_mode():
        y = model(batch)                      # AffWild2VA (resnet backbone), VA_3DResNet, ResNet, GRU, AttFusion; eval

Activations stay float32 channels-last; every GEMM / convolution is ONE launch of the ordinary bf16 tcgen05 kernel over
3x the contraction (split operands, csrc/fp32mode.cu): x = hi + lo, a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi, fp32
accumulation in TMEM.  BatchNorm is folded into the epilogue (eval), residual add + ReLU ride on the split pass of the
next layer, pooling / attention mix / the GRU recurrence are float32 kernels.  Forward only: training runs the bf16
path.  Everything below calls the C ABI; there is no PyTorch-math fallback.
"""
import contextlib

import torch

from . import lib as L
from . import ops, raw

_ON = False
NT = 3                 # product terms per contraction: 3 (two bf16 pieces per value, 2^-16) or 6 (three pieces, 2^-24)
BF16_OUT_F32 = 64 | 128   # tile_hint bits 6, 7 of m3t_conv_fprop_bf16: float32 output, channel-block-major K order


def enabled():
    return _ON


@contextlib.contextmanager
def parity_mode(on=True, terms=3):
    """terms=3: x = hi + lo, three bf16 products per fp32 product (2^-16 relative; 3x the tensor work).
    terms=6: three pieces per value and the six leading products (2^-24, fp32-grade; 6x the tensor work)."""
    global _ON, NT
    assert terms in (3, 6)
    old, _ON, NT = (_ON, NT), bool(on), terms
    try:
        yield
    finally:
        _ON, NT = old


def _lib():
    return L.load()


# ------------------------------------------------------------------------------------------------------------
# split passes
# ------------------------------------------------------------------------------------------------------------
def split3(x, res=None, relu=False, want_f32=False):
    """x: f32 [..., C] contiguous -> (bf16 [..., 3C] = [hi|hi|lo] of y, y f32 or None),  y = relu?(x + res)."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
    C = x.shape[-1]
    rows = x.numel() // C
    out3 = torch.empty(x.shape[:-1] + (NT * C,), device=x.device, dtype=torch.bfloat16)
    y = torch.empty_like(x) if want_f32 else None
    if res is not None:
        assert res.shape == x.shape and res.dtype == torch.float32 and res.is_contiguous()
    L.check(_lib().m3t_split3_bf16(L.ptr(x), L.ptr(res), L.i32(1 if relu else 0), L.ptr(y), L.ptr(out3), L.i64(rows),
                                   L.i32(C), L.i32(NT), L.stream_ptr()), "m3t_split3_bf16")
    return out3, y


def _pack_raw(s, N, G, C, tap_minor, cpad=0):
    out = torch.empty((N, G * (cpad or NT * C)), device=s.device, dtype=torch.bfloat16)
    L.check(_lib().m3t_pack_split3_bf16(L.ptr(s), L.ptr(out), L.i64(N), L.i32(G), L.i32(C), L.i32(tap_minor),
                                        L.i32(cpad), L.i32(NT), L.stream_ptr()), "m3t_pack_split3_bf16")
    return out


def _pack_w(w, tag, N, G, C, tap_minor, src=None, cpad=0):
    """bf16 [N, G*3C] split copy of a float32 parameter, cached until the parameter changes."""
    def make():
        return _pack_raw((w.detach() if src is None else src()).contiguous().float(), N, G, C, tap_minor, cpad)
    return ops._cached(w, "split%d:%s" % (NT, tag), make)


def linear(x, weight, bias, relu=False):
    """x: f32 [..., K] -> f32 [..., N] = act(x W^T + b) at fp32-grade accuracy on the bf16 tensor-core kernel."""
    K = x.shape[-1]
    N = weight.shape[0]
    x3, _ = split3(x.contiguous().view(-1, K))
    w3 = _pack_w(weight, "lin", N, 1, K, 0)
    out = raw.gemm(x3, w3, out_dtype=torch.float32, shift=bias.detach().float().contiguous() if bias is not None else None,
                   relu=relu)
    return out.view(x.shape[:-1] + (N,))


def conv2d_bn(x3, x_shape, conv, bn, relu):
    """x3: bf16 [N,H,W,3Cin] split input; eval-mode BN folded into the epilogue; returns f32 [N,P,Q,Cout]."""
    N, H, W, Cin = x_shape
    w = conv.weight
    Cout, _, kh, kw = w.shape
    stride, pad = conv.stride[0], conv.padding[0]
    w3 = _pack_w(w, "conv", Cout, kh * kw, Cin, 1)
    ss = raw.bn_fold(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, None, ops.BN_EPS)
    geom = raw.conv_geom(2, N, 1, H, W, NT * Cin, Cout, (1, kh, kw), (1, stride, stride), (0, pad, pad), (0, pad, pad),
                         (1, 1, 1))
    Z, P, Q = raw.conv_out_dims(geom)
    y = torch.empty((N, P, Q, Cout), device=x3.device, dtype=torch.float32)
    rc = _lib().m3t_conv_fprop_bf16(L.ptr(x3), L.ptr(w3), L.ptr(y), L.int_array(geom), L.ptr(ss[0]), L.ptr(ss[1]),
                                    L.ptr(None), L.i32(1 if relu else 0), L.ptr(None), L.i32(BF16_OUT_F32),
                                    L.stream_ptr())
    L.check(rc, "m3t_conv_fprop_bf16 (fp32 mode)")
    return y


def basic_block(blk, x, x3):
    """models/resnet.py:37-56 (eval).  x: f32 [N,H,W,C] block input, x3 its split copy -> (out f32, out3)."""
    h = conv2d_bn(x3, tuple(x.shape), blk.conv1, blk.bn1, relu=True)
    h3, _ = split3(h)
    y2 = conv2d_bn(h3, tuple(h.shape), blk.conv2, blk.bn2, relu=False)
    idt = x if blk.downsample is None else conv2d_bn(x3, tuple(x.shape), blk.downsample[0], blk.downsample[1], False)
    out3, out = split3(y2, res=idt, relu=True, want_f32=True)
    return out, out3


def resnet_trunk(resnet, x):
    """x: f32 [F,H,W,64] -> f32 [F,512] (agg_mode 'ap')."""
    if resnet.agg_mode != 'ap':
        raise NotImplementedError("fp32 parity mode covers agg_mode 'ap' (the hot path)")
    x3, _ = split3(x)
    for layer in (resnet.layer1, resnet.layer2, resnet.layer3, resnet.layer4):
        for blk in layer:
            x, x3 = basic_block(blk, x, x3)
    F_, H, W, C = x.shape
    out = torch.empty((F_, C), device=x.device, dtype=torch.float32)
    L.check(_lib().m3t_avgpool_f32(L.ptr(x), L.ptr(out), L.i32(F_), L.i32(H * W), L.i32(C), L.stream_ptr()),
            "m3t_avgpool_f32")
    return out


def stem3d(video, conv, bn, normalise):
    """models/backbone.py:327-332 (eval): video (B,3,T,H,W) f32/u8 -> f32 [B*T, H/4, W/4, 64]."""
    video = video.contiguous()
    B, _, T, H, W = video.shape
    xs3 = torch.empty((B, T, H // 2, W // 2, 64 * NT), device=video.device, dtype=torch.bfloat16)
    mul, add = (1.0 / 127.5, -1.0) if normalise else (1.0, 0.0)
    L.check(_lib().m3t_video_prep_s2d_w4_split3(L.ptr(video), L.i32(video.dtype == torch.uint8), L.ptr(xs3), L.i32(B),
                                                L.i32(T), L.i32(H), L.i32(W), L.f32(mul), L.f32(add), L.i32(NT),
                                                L.stream_ptr()),
            "m3t_video_prep_s2d_w4_split3")
    w = conv.weight
    idx = ops.stem_s2d_index(w.device).long()

    def gathered():   # f32 [64][20 taps][64 channels] in the W-unrolled space-to-depth order, zeros where idx < 0
        flat = w.detach().view(64, -1)
        return (flat[:, idx.clamp(min=0)] * (idx >= 0).to(flat.dtype)).contiguous()

    w3 = _pack_w(w, "stem", 64, 20, 64, 0, src=gathered)
    ss = raw.bn_fold(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, None, ops.BN_EPS)
    geom = raw.conv_geom(3, B, T, H // 2, W // 2, 64 * NT, 64, (5, 4, 1), (1, 1, 1), (2, 2, 0), (2, 1, 0), (1, 1, 1))
    y = torch.empty((B * T, H // 2, W // 2, 64), device=video.device, dtype=torch.float32)
    rc = _lib().m3t_conv_fprop_bf16(L.ptr(xs3), L.ptr(w3), L.ptr(y), L.int_array(geom), L.ptr(ss[0]), L.ptr(ss[1]),
                                    L.ptr(None), L.i32(1), L.ptr(None), L.i32(BF16_OUT_F32), L.stream_ptr())
    L.check(rc, "m3t_conv_fprop_bf16 (fp32 stem)")
    P, Q = (H // 2 - 1) // 2 + 1, (W // 2 - 1) // 2 + 1
    out = torch.empty((B * T, P, Q, 64), device=video.device, dtype=torch.float32)
    L.check(_lib().m3t_maxpool3s2_f32(L.ptr(y), L.ptr(out), L.i32(B * T), L.i32(H // 2), L.i32(W // 2), L.i32(64),
                                      L.stream_ptr()), "m3t_maxpool3s2_f32")
    return out


def gru_module(m, x):
    """models/rnn.py:71-81 (eval): x f32 [B,T,I] -> head output f32 (or BiGRU features when num_classes <= 0)."""
    g = m.gru
    B, T, _ = x.shape
    H = m.hidden_size
    h = x.contiguous()
    for l in range(m.num_layers):
        sfx = '_l%d' % l
        w_ih = torch.cat([getattr(g, 'weight_ih' + sfx), getattr(g, 'weight_ih' + sfx + '_reverse')]).detach()
        b_ih = torch.cat([getattr(g, 'bias_ih' + sfx), getattr(g, 'bias_ih' + sfx + '_reverse')]).detach()
        w_hh = torch.stack([getattr(g, 'weight_hh' + sfx), getattr(g, 'weight_hh' + sfx + '_reverse')]).detach()
        b_hh = torch.stack([getattr(g, 'bias_hh' + sfx), getattr(g, 'bias_hh' + sfx + '_reverse')]).detach()
        # input projection of all steps, both directions: [B*T, 2*3H] = x . [W_ih ; W_ih_reverse]^T + b_ih
        K = h.shape[-1]
        x3, _ = split3(h.view(B * T, K))
        w3 = _pack_raw(w_ih.contiguous().float(), 6 * H, 1, K, 0)
        gi = raw.gemm(x3, w3, out_dtype=torch.float32, shift=b_ih.float().contiguous())
        out = torch.empty((B, T, 2 * H), device=h.device, dtype=torch.float32)
        L.check(_lib().m3t_gru_fwd_f32(L.ptr(gi), L.ptr(w_hh.contiguous().float()), L.ptr(b_hh.contiguous().float()),
                                       L.ptr(out), L.i32(B), L.i32(T), L.i32(H), L.stream_ptr()), "m3t_gru_fwd_f32")
        h = out
    if m.num_classes <= 0:
        return h
    if isinstance(m.fc, torch.nn.Linear):
        return linear(h, m.fc.weight, m.fc.bias)
    lins = [mod for mod in m.fc if isinstance(mod, torch.nn.Linear)]
    for i, lin in enumerate(lins):
        h = linear(h, lin.weight, lin.bias, relu=i + 1 < len(lins))
    return h


def att_fusion(m, x_a, x_v):
    """models/att_fusion.py:18-27 (eval) on f32 [B,T,512] streams."""
    if m.use_proj:
        x_v = linear(x_v, m.proj_v.weight, m.proj_v.bias)
    s_v = gru_module(m.scorer_v, x_v).contiguous()
    s_a = gru_module(m.scorer_a, x_a).contiguous()
    x_a, x_v = x_a.contiguous(), x_v.contiguous()
    f = torch.empty_like(x_a)
    C = x_a.shape[-1]
    L.check(_lib().m3t_att_mix_f32(L.ptr(x_a), L.ptr(x_v), L.ptr(s_a), L.ptr(s_v), L.ptr(f), L.i64(x_a.numel() // C),
                                   L.i32(C), L.stream_ptr()), "m3t_att_mix_f32")
    return f


def va_3dresnet(m, video, normalise=False):
    """models/backbone.py:347-355 (eval): video -> GRU head output (or (B,T,512) features without a backend)."""
    conv, bn = m.c3d[0], m.c3d[1]
    B, T = video.shape[0], video.shape[2]
    x = stem3d(video, conv, bn, normalise)
    f = resnet_trunk(m.resnet, x).view(B, T, -1)
    if f.shape[1] != m.frameLen:
        raise RuntimeError("VA_3DResNet: T (%d) must equal frameLen (%d)" % (f.shape[1], m.frameLen))
    return gru_module(m.gru, f) if m.backend == 'gru' else f


def convnd_bias_bn(x3, x_shape, nd, w, w3, bias, bn, k, pad_lo, pad_hi, relu, cin3, dil=(1, 1, 1)):
    """Generic N-d stride-1 conv on a split input with conv bias + eval BN folded into the epilogue -> f32."""
    if nd == 3:
        N, D, H, W = x_shape
    else:
        N, W = x_shape
        D = H = 1
    Cout = w.shape[0]
    if bn is not None:
        ss = raw.bn_fold(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var,
                         bias.detach() if bias is not None else None, ops.BN_EPS)
        scale, shift = ss[0], ss[1]
    else:
        scale, shift = None, (bias.detach().float().contiguous() if bias is not None else None)
    geom = raw.conv_geom(nd, N, D, H, W, cin3, Cout, k, (1, 1, 1), pad_lo, pad_hi, dil)
    Z, P, Q = raw.conv_out_dims(geom)
    y = torch.empty((N, Z, P, Q, Cout), device=x3.device, dtype=torch.float32)
    rc = _lib().m3t_conv_fprop_bf16(L.ptr(x3), L.ptr(w3), L.ptr(y), L.int_array(geom), L.ptr(scale), L.ptr(shift),
                                    L.ptr(None), L.i32(1 if relu else 0), L.ptr(None), L.i32(BF16_OUT_F32),
                                    L.stream_ptr())
    L.check(rc, "m3t_conv_fprop_bf16 (fp32 mode)")
    return y


def vggm_stack(seq, x, video_first, normalise=False):
    """A Sequential of [Conv3d 3x3x3 pad(1,0,0) + BN3d + ReLU (+ MaxPool3d(1,2,2))] groups (models/backbone.py:73-103)
    in float32.  x: the raw video (B,3,T,H,W) when video_first, else f32 [B,T,H,W,C]."""
    mods = list(seq)
    i = 0
    while i < len(mods):
        conv, bn = mods[i], mods[i + 1]
        pool = i + 3 < len(mods) and isinstance(mods[i + 3], torch.nn.MaxPool3d)
        w = conv.weight
        Cout = w.shape[0]
        if video_first and i == 0:
            video = x.contiguous()
            B, _, T, H, W = video.shape
            cpad = (16 * NT + 63) // 64 * 64
            xs3 = torch.empty((B, T, H // 2, W // 2, cpad), device=video.device, dtype=torch.bfloat16)
            mul, add = (1.0 / 127.5, -1.0) if normalise else (1.0, 0.0)
            L.check(_lib().m3t_video_prep_s2d_split3(L.ptr(video), L.i32(video.dtype == torch.uint8), L.ptr(xs3),
                                                     L.i32(B), L.i32(T), L.i32(H), L.i32(W), L.f32(mul), L.f32(add),
                                                     L.i32(NT), L.i32(cpad), L.stream_ptr()), "m3t_video_prep_s2d_split3")
            idx = ops.vggm_s2d_index(w.device).long()

            def gathered(w=w, idx=idx, Cout=Cout):
                flat = w.detach().view(Cout, -1)
                return (flat[:, idx.clamp(min=0)] * (idx >= 0).to(flat.dtype)).contiguous()

            w3 = _pack_w(w, "vggm1", Cout, 12, 16, 0, src=gathered, cpad=cpad)
            y = convnd_bias_bn(xs3, (B, T, H // 2, W // 2), 3, w, w3, conv.bias, bn, (3, 2, 2), (1, 0, 0), (1, 0, 0),
                               True, cpad)
        else:
            B, T, H, W, C = x.shape
            x3, _ = split3(x)
            w3 = _pack_w(w, "conv3d", Cout, 27, C, 1)
            y = convnd_bias_bn(x3, (B, T, H, W), 3, w, w3, conv.bias, bn, (3, 3, 3), (1, 0, 0), (1, 0, 0), True, NT * C)
        if pool:
            Bq, Z, P, Q, Cq = y.shape
            out = torch.empty((Bq, Z, P // 2, Q // 2, Cq), device=y.device, dtype=torch.float32)
            L.check(_lib().m3t_maxpool2x2_f32(L.ptr(y), L.ptr(out), L.i32(Bq * Z), L.i32(P), L.i32(Q), L.i32(Cq),
                                              L.stream_ptr()), "m3t_maxpool2x2_f32")
            y = out
        x = y
        i += 4 if pool else 3
    return x


def tcn_simple(mlist, x):
    """[Conv1d(k, p=(k-1)/2) + BN1d + ReLU] x2 (+ Linear) on f32 (B,T,C) (models/backbone.py:214-231,284-289)."""
    seq = mlist[0]
    for ci in (0, 3):
        conv, bn = seq[ci], seq[ci + 1]
        k, p = conv.kernel_size[0], conv.padding[0]
        B, T, C = x.shape
        x3, _ = split3(x.contiguous())
        w3 = _pack_w(conv.weight, "conv1d", conv.weight.shape[0], k, C, 1)
        x = convnd_bias_bn(x3, (B, T), 1, conv.weight, w3, conv.bias, bn, (1, 1, k), (0, 0, p), (0, 0, p), True, NT * C)
        x = x.view(B, T, -1)
    if len(mlist) > 1:
        x = linear(x, mlist[1].weight, mlist[1].bias)
    return x


def temporal_conv_net(net, x):
    """models/tcn.py:43-64 (eval, dropout off): weight-normed dilated causal Conv1d pairs + residual, f32 (B,T,C)."""
    for blk in net.network:
        B, T, C = x.shape
        h = x
        for conv in (blk.conv1, blk.conv2):
            w = torch._weight_norm(conv.weight_v, conv.weight_g, 0).detach().contiguous()   # g * v / ||v|| per channel
            k, Ci = w.shape[2], w.shape[1]
            h3, _ = split3(h.contiguous())
            w3 = _pack_raw(w.float(), w.shape[0], k, Ci, 1)
            h = convnd_bias_bn(h3, (B, T), 1, w, w3, conv.bias, None, (1, 1, k), (0, 0, blk.padding), (0, 0, 0), True,
                               NT * Ci, dil=(1, 1, blk.dilation)).view(B, T, -1)
        if blk.downsample is None:
            res = x.contiguous()
        else:
            x3, _ = split3(x.contiguous())
            wd = blk.downsample.weight
            w3 = _pack_w(wd, "conv1d", wd.shape[0], 1, C, 1)
            res = convnd_bias_bn(x3, (B, T), 1, wd, w3, blk.downsample.bias, None, (1, 1, 1), (0, 0, 0), (0, 0, 0),
                                 False, NT * C).view(B, T, -1)
        _, x = split3(h.contiguous(), res=res, relu=True, want_f32=True)
    return x


def require_eval(module):
    if module.training:
        raise NotImplementedError("fp32 parity mode is forward / eval only; training runs the bf16 path")
