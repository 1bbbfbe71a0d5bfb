"""Temporal convolutional network (reference: models/tcn.py).

Same module tree and state_dict keys as the reference (`network.{i}.conv{1,2}.{bias,weight_g,weight_v}`, their
aliases under `net.{0,4}`, `downsample.*`).  Each weight-normed dilated causal Conv1d runs as an implicit-GEMM
conv over channels-last (B,T,C) with left-only zero padding -- Chomp1d becomes index math -- and the rest of the block
rides in the conv epilogues (ops.TCNConvFn / m3t_tcn_conv_bf16): weight-norm scale, bias, ReLU, dropout, and for the
second conv the residual add + final ReLU.  M3T_TCN_FUSED=0 keeps the round-1 path (conv + separate dropout / add).
"""
import os
import warnings

import torch
import torch.nn as nn
from torch.nn.utils import weight_norm

from .. import fp32, ops


class Chomp1d(nn.Module):
    def __init__(self, chomp_size):
        super().__init__()
        self.chomp_size = chomp_size

    def forward(self, x):
        return x[:, :, :-self.chomp_size].contiguous()


def _wn_weight(conv):
    return torch._weight_norm(conv.weight_v, conv.weight_g, 0)


def _causal_conv(c_in, c_out, kernel_size, dilation, padding):
    """Weight-normed Conv1d holding the reference's parameters (`bias`, `weight_g`, `weight_v`); never called as a
    module here — forward_cl feeds its weights to the implicit-GEMM kernel."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # torch deprecates this weight_norm; the key names depend on it
        return weight_norm(nn.Conv1d(c_in, c_out, kernel_size, stride=1, padding=padding, dilation=dilation))


class TemporalBlock(nn.Module):
    def __init__(self, n_inputs, n_outputs, kernel_size, stride, dilation, padding, dropout=0.2):
        super().__init__()
        assert stride == 1, "the reference only builds stride-1 temporal blocks"
        self.dilation, self.padding = dilation, padding
        stages = []
        for i, c_in in ((1, n_inputs), (2, n_outputs)):
            stage = (("conv", _causal_conv(c_in, n_outputs, kernel_size, dilation, padding)),
                     ("chomp", Chomp1d(padding)), ("relu", nn.ReLU()), ("dropout", nn.Dropout(dropout)))
            for name, mod in stage:
                setattr(self, "%s%d" % (name, i), mod)          # conv1, chomp1, relu1, dropout1, conv2, ...
                stages.append(mod)
        self.net = nn.Sequential(*stages)                       # the reference's `net.{0,4}` aliases of conv1 / conv2
        self.downsample = None if n_inputs == n_outputs else nn.Conv1d(n_inputs, n_outputs, 1)
        self.relu = nn.ReLU()
        self.init_weights()

    def init_weights(self):
        # N(0, 0.01) lands in the DERIVED `.weight` of the weight-normed convs, which weight_norm recomputes from
        # weight_g / weight_v at the next forward (so only the plain 1x1 `downsample` really gets it)
        for conv in (self.conv1, self.conv2, self.downsample):
            if conv is not None:
                conv.weight.data.normal_(0, 0.01)

    def forward_cl(self, x):
        """Two launches per block (three when a 1x1 `downsample` exists): each conv carries weight-norm scale, bias,
        ReLU, dropout and — the second one — the residual add + final ReLU in its epilogue (ops.TCNConvFn)."""
        if os.environ.get("M3T_TCN_FUSED", "1") == "1":
            if self.downsample is None:
                res = x
            else:
                res = ops.Conv1dBiasAct.apply(x, self.downsample.weight, self.downsample.bias, 1, 0, 0, False)
            h = ops.TCNConvFn.apply(x, self.conv1.weight_v, self.conv1.weight_g, self.conv1.bias, None, self.dilation,
                                    self.padding, self.dropout1.p, self.training)
            return ops.TCNConvFn.apply(h, self.conv2.weight_v, self.conv2.weight_g, self.conv2.bias, res,
                                       self.dilation, self.padding, self.dropout2.p, self.training)
        h = x
        for conv, drop in ((self.conv1, self.dropout1), (self.conv2, self.dropout2)):
            h = ops.Conv1dBiasAct.apply(h, _wn_weight(conv), conv.bias, self.dilation, self.padding, 0, True)
            if self.training and drop.p > 0:
                if ops.fused_dropout_enabled():
                    h = ops.DropoutFn.apply(h, drop.p)
                else:
                    h = torch.nn.functional.dropout(h, drop.p, True)
        if self.downsample is None:
            res = x
        else:
            res = ops.Conv1dBiasAct.apply(x, self.downsample.weight, self.downsample.bias, 1, 0, 0, False)
        return ops.AddReLU.apply(h, res)

    def forward(self, x):
        return ops.FromCL.apply(self.forward_cl(ops.ToCL.apply(x)))


class TemporalConvNet(nn.Module):
    def __init__(self, num_inputs, num_channels, kernel_size=2, dropout=0.2):
        super().__init__()
        layers = []
        for i, out_ch in enumerate(num_channels):
            d = 2 ** i
            in_ch = num_inputs if i == 0 else num_channels[i - 1]
            layers.append(TemporalBlock(in_ch, out_ch, kernel_size, stride=1, dilation=d,
                                        padding=(kernel_size - 1) * d, dropout=dropout))
        self.network = nn.Sequential(*layers)

    def forward_cl(self, x):
        if fp32.enabled():       # fp32-parity inference mode (m3t_b200.fp32)
            fp32.require_eval(self)
            return fp32.temporal_conv_net(self, ops.as_f32(x))
        for blk in self.network:
            x = blk.forward_cl(x)
        return x

    def forward(self, x):
        if fp32.enabled():
            return self.forward_cl(x.float().transpose(1, 2).contiguous()).transpose(1, 2).contiguous()
        return ops.FromCL.apply(self.forward_cl(ops.ToCL.apply(x)))
