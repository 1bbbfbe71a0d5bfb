"""Recurrent heads (reference: models/rnn.py:11-81 `GRU`).

`self.gru` is a stock nn.GRU used purely as the parameter container (state_dict keys `gru.weight_ih_l0[_reverse]`,
...); the forward pass is: one tcgen05 GEMM per layer for the input projection of all time steps, the persistent
recurrence kernel (gru.cu), and GEMM(+bias+ReLU) epilogues for the FC head.
"""
import math

import torch
import torch.nn as nn

from .. import fp32, ops


class GRU(nn.Module):
    def __init__(self, input_size, hidden_size, num_layers, num_classes, num_fcs=1, dropout=False, return_h=False):
        super().__init__()
        self.hidden_size = hidden_size
        self.num_layers = num_layers
        self.num_classes = num_classes
        self.return_h = return_h
        if return_h:
            raise NotImplementedError("return_h is only used by AttEncDec (fusion_type 'att_dec'), out of scope")
        self.gru = nn.GRU(input_size, hidden_size, num_layers, batch_first=True, bidirectional=True)
        if num_classes > 0:
            dims = {1: [hidden_size * 2, num_classes], 2: [hidden_size * 2, hidden_size, num_classes],
                    3: [hidden_size * 2, hidden_size, hidden_size, num_classes]}[num_fcs]
            if num_fcs == 1:
                self.fc = nn.Linear(dims[0], dims[1])
            else:
                mods = []
                for i in range(len(dims) - 1):
                    mods.append(nn.Linear(dims[i], dims[i + 1]))
                    if i + 2 < len(dims):
                        mods.append(nn.ReLU(True))
                        if dropout:
                            mods.append(nn.Dropout(0.5))
                self.fc = nn.Sequential(*mods)
        bound = math.sqrt(3.0) * math.sqrt(2.0 / (input_size + hidden_size))
        H = hidden_size
        for name, prm in self.gru.named_parameters():
            for g in range(3):
                blk = prm.data[g * H:(g + 1) * H]
                if 'weight_ih' in name:
                    nn.init.uniform_(blk, -bound, bound)
                elif 'weight_hh' in name:
                    nn.init.orthogonal_(blk)
                else:
                    nn.init.zeros_(blk)

    def _head(self, h):
        if self.num_classes <= 0:
            return h
        if isinstance(self.fc, nn.Linear):
            return ops.linear(h, self.fc.weight, self.fc.bias, relu=False, out_f32=True)
        lins = [m for m in self.fc if isinstance(m, nn.Linear)]
        drops = [m for m in self.fc if isinstance(m, nn.Dropout)]
        if drops and self.training:
            raise NotImplementedError("FC-head dropout (GRU(dropout=True)) is not used by the reference's hot path")
        for i, lin in enumerate(lins):
            last = i + 1 == len(lins)
            h = ops.linear(h, lin.weight, lin.bias, relu=not last, out_f32=last)
        return h

    def forward_bf16(self, x):
        """x: [B,T,I] (fp32 or bf16) -> head output (fp32) or raw BiGRU features (bf16) when num_classes <= 0."""
        if fp32.enabled():
            fp32.require_eval(self)
            return fp32.gru_module(self, ops.as_f32(x))
        h = ops.as_bf16(x)
        g = self.gru
        for l in range(self.num_layers):
            sfx = '_l%d' % l
            h = ops.GRULayerFn.apply(
                h, getattr(g, 'weight_ih' + sfx), getattr(g, 'weight_hh' + sfx), getattr(g, 'bias_ih' + sfx),
                getattr(g, 'bias_hh' + sfx), getattr(g, 'weight_ih' + sfx + '_reverse'),
                getattr(g, 'weight_hh' + sfx + '_reverse'), getattr(g, 'bias_ih' + sfx + '_reverse'),
                getattr(g, 'bias_hh' + sfx + '_reverse'), torch.is_grad_enabled())
        return self._head(h)

    def forward(self, x):
        return ops.as_f32(self.forward_bf16(x))
