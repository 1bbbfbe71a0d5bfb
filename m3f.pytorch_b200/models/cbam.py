"""Convolutional block attention (reference: models/cbam.py:1-112 `CBAM` = `ChannelGate` + `SpatialGate`).

Same module tree and state_dict keys as the reference (`ChannelGate.mlp.{1,3}.*`, `SpatialGate.spatial.conv.weight`,
`SpatialGate.spatial.bn.*`).  Feature maps stay channels-last bf16; every pass over them is a kernel of csrc/cbam.cu
(pooling over pixels / over channels with arg-max bookkeeping, the two broadcast scalings with their reductions, the
2-channel 5x5 gate convolution).  What remains on the host side acts on O(F*C) or O(F*H*W) values: the shared MLP runs
on the tcgen05 GEMM (ops.linear), the sigmoid and the one-channel BatchNorm2d (momentum 0.01) are fp32 tensor glue.
"""
import torch
import torch.nn as nn

from .. import ops, raw


class Flatten(nn.Module):
    def forward(self, x):
        return x.view(x.size(0), -1)


class BasicConv(nn.Module):
    """Parameter container of the reference's BasicConv (models/cbam.py:15-29); run by SpatialGate.forward_cl."""

    def __init__(self, in_planes, out_planes, kernel_size, stride=1, padding=0, dilation=1, groups=1, relu=True,
                 bn=True, bias=False):
        super().__init__()
        self.out_channels = out_planes
        self.conv = nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding,
                              dilation=dilation, groups=groups, bias=bias)
        self.bn = nn.BatchNorm2d(out_planes, eps=1e-5, momentum=0.01, affine=True) if bn else None
        self.relu = nn.ReLU() if relu else None


class _PoolHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        avg, mx, arg = raw.cbam_pool_hw(x)
        ctx.save_for_backward(arg)
        ctx.shape = tuple(x.shape)
        ctx.mark_non_differentiable(arg)
        return avg, mx, arg

    @staticmethod
    def backward(ctx, davg, dmx, _):
        (arg,) = ctx.saved_tensors
        return raw.cbam_pool_hw_bwd(davg.contiguous(), dmx.contiguous(), arg, ctx.shape)


class _ScaleC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, sc):
        ctx.save_for_backward(x, sc)
        return raw.cbam_scale_c(x, sc)

    @staticmethod
    def backward(ctx, dy):
        x, sc = ctx.saved_tensors
        return raw.cbam_scale_c_bwd(dy.contiguous(), x, sc)


class _PoolC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        comp, carg = raw.cbam_pool_c(x)
        ctx.save_for_backward(carg)
        ctx.shape = tuple(x.shape)
        return comp

    @staticmethod
    def backward(ctx, dcomp):
        (carg,) = ctx.saved_tensors
        return raw.cbam_pool_c_bwd(dcomp.contiguous(), carg, ctx.shape)


class _ScaleS(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ss):
        ctx.save_for_backward(x, ss)
        return raw.cbam_scale_s(x, ss)

    @staticmethod
    def backward(ctx, dy):
        x, ss = ctx.saved_tensors
        return raw.cbam_scale_s_bwd(dy.contiguous(), x, ss)


class _Conv5(torch.autograd.Function):
    @staticmethod
    def forward(ctx, comp, w):
        ctx.save_for_backward(comp, w)
        return raw.cbam_conv5(comp, w)

    @staticmethod
    def backward(ctx, dout):
        comp, w = ctx.saved_tensors
        return raw.cbam_conv5_bwd(dout.contiguous(), comp, w)


class ChannelGate(nn.Module):
    def __init__(self, gate_channels, reduction_ratio=16):
        super().__init__()
        self.gate_channels = gate_channels
        self.mlp = nn.Sequential(Flatten(), nn.Linear(gate_channels, gate_channels // reduction_ratio), nn.ReLU(),
                                 nn.Linear(gate_channels // reduction_ratio, gate_channels))

    def _mlp(self, v):
        """(F,C) fp32 -> (F,C) fp32 on the tcgen05 GEMM; the hidden width C/16 is zero-padded to a multiple of 8."""
        l1, l2 = self.mlp[1], self.mlp[3]
        hid = l1.out_features
        hpad = (hid + 7) // 8 * 8
        w1, b1, w2 = l1.weight, l1.bias, l2.weight
        if hpad != hid:
            w1 = torch.cat((w1, w1.new_zeros(hpad - hid, w1.shape[1])), dim=0)
            b1 = torch.cat((b1, b1.new_zeros(hpad - hid)))
            w2 = torch.cat((w2, w2.new_zeros(w2.shape[0], hpad - hid)), dim=1)
        h = ops.linear(v, w1, b1, relu=True)
        return ops.linear(h, w2, l2.bias, out_f32=True)

    def forward_cl(self, x):
        avg, mx, _ = _PoolHW.apply(x)
        scale = torch.sigmoid(self._mlp(avg) + self._mlp(mx))
        return _ScaleC.apply(x, scale.contiguous())


class ChannelPool(nn.Module):
    def forward_cl(self, x):
        return _PoolC.apply(x)          # fp32 (F,2,H,W): channel max, channel mean


class SpatialGate(nn.Module):
    def __init__(self):
        super().__init__()
        kernel_size = 5
        self.compress = ChannelPool()
        self.spatial = BasicConv(2, 1, kernel_size, stride=1, padding=(kernel_size - 1) // 2, relu=False)

    def forward_cl(self, x):
        comp = self.compress.forward_cl(x)
        g = _Conv5.apply(comp, self.spatial.conv.weight)                     # (F,1,H,W) fp32
        bn = self.spatial.bn
        if bn.training:                 # BatchNorm2d(1), momentum 0.01: scalars over the whole map
            mean, var = g.mean(), g.var(unbiased=False)
            with torch.no_grad():
                n = g.numel()
                bn.running_mean.mul_(1 - bn.momentum).add_(bn.momentum * mean.detach())
                bn.running_var.mul_(1 - bn.momentum).add_(bn.momentum * var.detach() * n / max(n - 1, 1))
            ops.bump_num_batches_tracked(bn)
        else:
            mean, var = bn.running_mean[0], bn.running_var[0]
        g = (g - mean) / torch.sqrt(var + bn.eps) * bn.weight[0] + bn.bias[0]
        return _ScaleS.apply(x, torch.sigmoid(g).contiguous())


class CBAM(nn.Module):
    def __init__(self, gate_channels, reduction_ratio=16):
        super().__init__()
        self.ChannelGate = ChannelGate(gate_channels, reduction_ratio)
        self.SpatialGate = SpatialGate()

    def forward_cl(self, x):
        """x: bf16 (F,H,W,C) -> bf16 (F,H,W,C)."""
        return self.SpatialGate.forward_cl(self.ChannelGate.forward_cl(x.contiguous()))

    def forward(self, x):
        return ops.FromCL.apply(self.forward_cl(ops.ToCL.apply(x)))
