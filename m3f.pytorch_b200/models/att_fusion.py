"""Attention fusion of the audio and visual streams (reference: models/att_fusion.py:8-27)."""
import torch.nn as nn

from .. import fp32, ops, streams
from .rnn import GRU


class AttFusion(nn.Module):
    def __init__(self, input_dim=[512, 512], hidden_dim=128):
        super().__init__()
        self.use_proj = input_dim[1] != input_dim[0]
        if self.use_proj:
            self.proj_v = nn.Linear(input_dim[1], input_dim[0])
        self.scorer_a = GRU(input_dim[0], hidden_dim, 1, 1, 1)
        self.scorer_v = GRU(input_dim[0], hidden_dim, 1, 1, 1)

    def forward_bf16(self, x_a, x_v):
        if fp32.enabled():
            fp32.require_eval(self)
            return fp32.att_fusion(self, ops.as_f32(x_a), ops.as_f32(x_v))
        x_a, x_v = ops.as_bf16(x_a), ops.as_bf16(x_v)
        if self.use_proj:
            x_v = ops.linear(x_v, self.proj_v.weight, self.proj_v.bias)
        if streams.overlap_ok(x_a):
            # inference: the two scorers are independent -> one of them runs on a side stream (streams.py)
            s_a = streams.run_on_side(0, x_a.device, self.scorer_a.forward_bf16, x_a)
            s_v = self.scorer_v.forward_bf16(x_v)
            streams.join(x_a.device, 0)
        else:
            s_v = self.scorer_v.forward_bf16(x_v)     # (B,T,1) fp32 logits
            s_a = self.scorer_a.forward_bf16(x_a)
        return ops.att_mix(x_a, x_v, s_a, s_v)

    def forward(self, x_a, x_v):
        return ops.as_f32(self.forward_bf16(x_a, x_v))
