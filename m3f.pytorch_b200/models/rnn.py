"""Recurrent heads (reference: models/rnn.py:11-81 `GRU`; :84-165 `Attention` / `Decoder` / `AttEncDec`).

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
        finals = []
        for l in range(self.num_layers):
            sfx = '_l%d' % l
            h = ops.GRULayerFn.apply(
                h, getattr(g, 'weight_ih' + sfx), getattr(g, 'weight_hh' + sfx), getattr(g, 'bias_ih' + sfx),
                getattr(g, 'bias_hh' + sfx), getattr(g, 'weight_ih' + sfx + '_reverse'),
                getattr(g, 'weight_hh' + sfx + '_reverse'), getattr(g, 'bias_ih' + sfx + '_reverse'),
                getattr(g, 'bias_hh' + sfx + '_reverse'), torch.is_grad_enabled())
            if self.return_h:       # nn.GRU's h_n: final state of every (layer, direction)
                H = self.hidden_size
                finals += [h[:, -1, :H], h[:, 0, H:]]
        if self.return_h:
            return self._head(h), torch.stack(finals)          # (B,T,..), (2*num_layers, B, H)
        return self._head(h)

    def forward(self, x):
        if self.return_h:
            out, h_n = self.forward_bf16(x)
            return ops.as_f32(out), ops.as_f32(h_n)
        return ops.as_f32(self.forward_bf16(x))


class Attention(nn.Module):
    """Additive attention of the decoder (reference models/rnn.py:84-109): energy_t = v . relu(W [h ; enc_t] + b),
    weights = softmax over t.  W is split into its hidden and encoder halves, so the encoder half is ONE GEMM per
    sequence (`precompute`) and a decoding step only projects the (B,H) hidden state."""

    def __init__(self, hidden_size):
        super().__init__()
        self.hidden_size = hidden_size
        self.attn = nn.Linear(hidden_size * 2, hidden_size)
        self.v = nn.Parameter(torch.rand(hidden_size))
        stdv = 1. / math.sqrt(self.v.size(0))
        self.v.data.uniform_(-stdv, stdv)

    def precompute(self, encoder_outputs):
        """(B,T,H) bf16 -> fp32 (B,T,H): W_enc enc_t + b."""
        H = self.hidden_size
        return ops.linear(encoder_outputs, self.attn.weight[:, H:].contiguous(), self.attn.bias, out_f32=True)

    def forward(self, hidden, encoder_outputs, enc_proj=None):
        """hidden (B,H), encoder_outputs (B,T,H) -> (B,1,T) attention weights."""
        H = self.hidden_size
        if enc_proj is None:
            enc_proj = self.precompute(ops.as_bf16(encoder_outputs))
        h_proj = ops.linear(hidden, self.attn.weight[:, :H].contiguous(), None, out_f32=True)      # (B,H)
        energy = torch.relu(enc_proj + h_proj.unsqueeze(1))                                        # (B,T,H)
        return torch.softmax((energy * self.v).sum(-1), dim=1).unsqueeze(1)


class Decoder(nn.Module):
    """One decoding step (reference models/rnn.py:112-140): attention context + previous output -> GRU cell -> Linear.
    `self.gru` is an nn.GRU used as the parameter container; the cell runs as two tcgen05 GEMMs (input and hidden
    projections, ops.linear) and fp32 gate arithmetic on the (B,H) state."""

    def __init__(self, embed_size=128, hidden_size=512, output_size=2, n_layers=1):
        super().__init__()
        self.embed_size, self.hidden_size = embed_size, hidden_size
        self.output_size, self.n_layers = output_size, n_layers
        if n_layers != 1:
            raise NotImplementedError("the reference builds a one-layer decoder (models/rnn.py:146)")
        self.attention = Attention(hidden_size)
        self.gru = nn.GRU(hidden_size + embed_size, hidden_size, n_layers, batch_first=True)
        self.out = nn.Linear(hidden_size * 2, output_size)

    def _cell(self, x, h):
        """nn.GRU single step: x (B,I) fp32, h (B,H) fp32 -> h' (B,H) fp32; gate order [r, z, n]."""
        H = self.hidden_size
        I = x.shape[1]
        ipad = (I + 7) // 8 * 8
        if ipad != I:       # the GEMM takes K in multiples of 8: zero columns on both operands
            x = torch.cat((x, x.new_zeros(x.shape[0], ipad - I)), dim=1)
            w_ih = torch.cat((self.gru.weight_ih_l0, self.gru.weight_ih_l0.new_zeros(3 * H, ipad - I)), dim=1)
        else:
            w_ih = self.gru.weight_ih_l0
        gi = ops.linear(x, w_ih, self.gru.bias_ih_l0, out_f32=True)
        gh = ops.linear(h, self.gru.weight_hh_l0, self.gru.bias_hh_l0, out_f32=True)
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        return (1 - z) * n + z * h

    def forward(self, inputs, last_hidden, encoder_outputs, enc_proj=None):
        """inputs (B,E), last_hidden (1,B,H), encoder_outputs (B,T,H) -> output (B,O), hidden (1,B,H), weights (B,1,T)."""
        enc = ops.as_f32(encoder_outputs)
        h = ops.as_f32(last_hidden[-1])
        attn_weights = self.attention(h, encoder_outputs, enc_proj)
        context = (attn_weights.transpose(1, 2) * enc).sum(dim=1)          # (B,H): bmm of the reference as a reduction
        h_new = self._cell(torch.cat((ops.as_f32(inputs), context), dim=1), h)
        output = ops.linear(torch.cat((h_new, context), dim=1), self.out.weight, self.out.bias, out_f32=True)
        return output, h_new.unsqueeze(0), attn_weights


class AttEncDec(nn.Module):
    """Attention encoder-decoder fusion head (reference models/rnn.py:143-165; `--fusion_type att_dec`): a 2-layer
    BiGRU encoder over the concatenated audio-visual features, then T-1 sequential decoding steps that feed the
    previous valence / arousal prediction (or, with probability `teacher_forcing_ratio`, the label) back in."""

    def __init__(self):
        super().__init__()
        self.encoder = GRU(1024, 512, 2, -1, return_h=True)
        self.decoder = Decoder(2, 512, 2, 1)

    def forward(self, src, trg=None, teacher_forcing_ratio=0.5):
        import random
        batch_size = src.size(0)
        max_len = trg.size(1) if trg is not None else src.size(1)
        encoder_output, hidden = self.encoder.forward_bf16(src)
        H = self.decoder.hidden_size
        encoder_output = ops.as_f32(encoder_output)
        encoder_output = encoder_output[:, :, :H] + encoder_output[:, :, H:]      # sum the two directions
        hidden = ops.as_f32(hidden[:self.decoder.n_layers])
        enc_proj = self.decoder.attention.precompute(ops.as_bf16(encoder_output))
        output = encoder_output.new_zeros(batch_size, 2)
        steps = [encoder_output.new_zeros(batch_size, 2)]
        for t in range(1, max_len):
            output, hidden, _ = self.decoder(output, hidden, encoder_output, enc_proj)
            steps.append(output)
            is_teacher = trg is not None and random.random() < teacher_forcing_ratio
            output = trg.data[:, t].to(output) if is_teacher else output
        return torch.stack(steps, dim=1)

