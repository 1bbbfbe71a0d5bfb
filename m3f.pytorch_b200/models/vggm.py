"""VGG-M style 3-D visual streams (reference: models/backbone.py:62-161 `VA_3DVGGM`, :164-311 `VA_3DVGGM_Split`).

Module trees and state_dict keys follow the reference (`v2p.*` / `shared.*`, `v_private.*`, `a_private.*`, `gru*`,
`tcn*`).  Every [Conv3d 3x3x3 pad(1,0,0) (+bias) -> BatchNorm3d -> ReLU (-> MaxPool3d (1,2,2))] group runs as one
ops.ConvNdBNAct unit on channels-last (B,T,H,W,C) bf16; the first conv (3 input channels, stride (1,2,2)) is
evaluated over the 2x2 space-to-depth image, like the ResNet stem.
"""
import torch
import torch.nn as nn

from .. import fp32, ops, raw, streams
from .backbone import _init_like_reference
from .rnn import GRU
from .tcn import TemporalConvNet


def _norm(norm_layer, ch):
    if norm_layer != 'bn':
        raise NotImplementedError("GroupNorm variant of VGG-M is not used by the reference's defaults")
    return nn.BatchNorm3d(ch)


def _conv_group(cin, cout, first, pool, norm_layer):
    mods = [nn.Conv3d(cin, cout, 3, stride=(1, 2, 2) if first else 1, padding=(1, 0, 0)), _norm(norm_layer, cout),
            nn.ReLU(True)]
    if pool:
        mods.append(nn.MaxPool3d(kernel_size=(1, 2, 2), stride=(1, 2, 2)))
    return mods


_CHANNELS = {1: (3, 64), 2: (64, 128), 3: (128, 256), 4: (256, 512), 5: (512, 512)}


def _run_stack(seq, x, first_is_s2d):
    """Run a Sequential of conv groups on a CL tensor (the s2d image for the very first conv)."""
    mods = list(seq)
    i = 0
    while i < len(mods):
        conv, bn = mods[i], mods[i + 1]
        assert isinstance(conv, nn.Conv3d) and isinstance(bn, nn.BatchNorm3d) and isinstance(mods[i + 2], nn.ReLU)
        pool = i + 3 < len(mods) and isinstance(mods[i + 3], nn.MaxPool3d)
        s2d = first_is_s2d and i == 0
        cfg = dict(nd=3, k=(3, 2, 2) if s2d else (3, 3, 3), pad_lo=(1, 0, 0), pad_hi=(1, 0, 0), relu=True,
                   pool=(2, 2, 0) if pool else None, s2d_first=s2d)
        x = ops.ConvNdBNAct.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, cfg,
                                  bn.training)
        if bn.training:
            ops.bump_num_batches_tracked(bn)
        i += 4 if pool else 3
    return x


def _tcn_simple(mlist, x):
    """[Conv1d(k,p=(k-1)/2) + BN1d + ReLU] x2 (+ Linear) on CL (B,T,C) (reference :214-231,284-289)."""
    seq = mlist[0]
    for ci in (0, 3):
        conv, bn = seq[ci], seq[ci + 1]
        k, p = conv.kernel_size[0], conv.padding[0]
        cfg = dict(nd=1, k=(1, 1, k), pad_lo=(0, 0, p), pad_hi=(0, 0, p), relu=True, pool=None)
        x = ops.ConvNdBNAct.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, cfg,
                                  bn.training)
        if bn.training:
            ops.bump_num_batches_tracked(bn)
    if len(mlist) > 1:
        x = ops.linear(x, mlist[1].weight, mlist[1].bias, out_f32=True)
    return x


def _tcn_simple_modules(cin, hidden, k):
    return nn.Sequential(nn.Conv1d(cin, hidden, k, 1, (k - 1) // 2), nn.BatchNorm1d(512), nn.ReLU(True),
                         nn.Conv1d(hidden, hidden, k, 1, (k - 1) // 2), nn.BatchNorm1d(512), nn.ReLU(True))


def _squeeze_features(x):
    """(B,T,1,1,C) -> (B,T,C); mirrors `.squeeze()` of the reference without its B==1 / T==1 hazard."""
    B, T = x.shape[0], x.shape[1]
    if x.shape[2] != 1 or x.shape[3] != 1:
        raise RuntimeError("VGG-M towers expect 112x112 inputs (final feature map must be 1x1)")
    return x.view(B, T, x.shape[-1])


def _cat_cl(x, feats):
    """cat((x, se), dim=channels) with se given as fp32 (B,C,T) like the reference."""
    return torch.cat((x, ops.ToCL.apply(feats)), dim=-1)


class VA_3DVGGM(nn.Module):
    def __init__(self, inputDim=512, hiddenDim=512, nLayers=2, nClasses=2, frameLen=16, backend='gru',
                 norm_layer='bn', nFCs=1):
        super().__init__()
        self.inputDim, self.hiddenDim, self.nClasses = inputDim, hiddenDim, nClasses
        self.frameLen, self.nLayers, self.backend, self.nFCs = frameLen, nLayers, backend, nFCs
        mods = []
        for c in range(1, 6):
            mods += _conv_group(*_CHANNELS[c], first=(c == 1), pool=(c <= 3), norm_layer=norm_layer)
        self.v2p = nn.Sequential(*mods)
        if backend == 'gru':
            self.gru = GRU(inputDim, hiddenDim, nLayers, nClasses, nFCs)
        elif backend == 'tcn':
            self.tcn = nn.ModuleList([TemporalConvNet(inputDim, [hiddenDim] * nLayers, 3), nn.Linear(hiddenDim, 2)])
        elif backend == 'tcn_simple':
            self.tcn = nn.ModuleList([_tcn_simple_modules(inputDim, hiddenDim, 3), nn.Linear(hiddenDim, 2)])
        elif backend == 'fc':
            self.fc = nn.Sequential(nn.Linear(512, hiddenDim), nn.ReLU(True), nn.Linear(hiddenDim, nClasses))
        _init_like_reference(self)

    def forward_bf16(self, video, *unused, normalise=False, after_features=None):
        if fp32.enabled():       # fp32-parity inference mode (m3t_b200.fp32)
            fp32.require_eval(self)
            f = _squeeze_features(fp32.vggm_stack(self.v2p, video, True, normalise))
            if self.backend == 'gru':
                return self.gru.forward_bf16(f)
            if self.backend == 'tcn':
                return fp32.linear(self.tcn[0].forward_cl(f), self.tcn[1].weight, self.tcn[1].bias)
            if self.backend == 'tcn_simple':
                return fp32.tcn_simple(self.tcn, f)
            h = fp32.linear(f, self.fc[0].weight, self.fc[0].bias, relu=True)
            return fp32.linear(h, self.fc[2].weight, self.fc[2].bias).mean(dim=1)
        xs = raw.video_prep_s2d(video.contiguous(), normalise)
        f = _squeeze_features(_run_stack(self.v2p, xs, True))           # (B,T,512)
        if after_features is not None:
            after_features()
        if self.backend == 'gru':
            return self.gru.forward_bf16(f)
        if self.backend == 'tcn':
            h = self.tcn[0].forward_cl(f)
            return ops.linear(h, self.tcn[1].weight, self.tcn[1].bias, out_f32=True)
        if self.backend == 'tcn_simple':
            return _tcn_simple(self.tcn, f)
        h = ops.linear(f, self.fc[0].weight, self.fc[0].bias, relu=True)
        return ops.linear(h, self.fc[2].weight, self.fc[2].bias, out_f32=True).mean(dim=1)

    def forward(self, x, *unused):
        return ops.as_f32(self.forward_bf16(x))


class VA_3DVGGM_Split(nn.Module):
    def __init__(self, inputDim=512, hiddenDim=512, nLayers=2, frameLen=16, nClasses=2, backend='gru',
                 norm_layer='bn', split_layer=5, nFCs=1, use_mtl=False):
        super().__init__()
        self.inputDim, self.hiddenDim, self.frameLen, self.nLayers = inputDim, hiddenDim, frameLen, nLayers
        self.nClasses, self.backend, self.split_layer, self.norm_layer = nClasses, backend, split_layer, norm_layer
        self.nFCs, self.use_mtl = nFCs, use_mtl
        assert split_layer >= 2, 'degenerate multi-tower structure'
        shared, v_priv, a_priv = [], [], []
        for c in range(1, 6):
            grp = dict(first=(c == 1), pool=(c <= 3), norm_layer=norm_layer)
            if c <= split_layer:
                shared += _conv_group(*_CHANNELS[c], **grp)
            else:
                v_priv += _conv_group(*_CHANNELS[c], **grp)
                a_priv += _conv_group(*_CHANNELS[c], **grp)
        self.shared = nn.Sequential(*shared)
        if split_layer != 5:
            self.v_private = nn.Sequential(*v_priv)
            self.a_private = nn.Sequential(*a_priv)
        if backend == 'gru':
            if split_layer == 5:
                self.gru = GRU(inputDim + 512 + 512, hiddenDim, nLayers, nClasses, nFCs)
            else:
                self.gru_v = GRU(inputDim + 512, hiddenDim, nLayers, nClasses - 1, nFCs)
                self.gru_a = GRU(inputDim + 512, hiddenDim, nLayers, min(nClasses, 1), nFCs)
        elif backend == 'tcn_simple' and split_layer != 5:
            self.tcn_v = nn.ModuleList([_tcn_simple_modules(inputDim + 512, hiddenDim, 5)])
            self.tcn_a = nn.ModuleList([_tcn_simple_modules(inputDim + 512, hiddenDim, 5)])
            if use_mtl:
                if nClasses > 0:
                    self.tcn_v.append(nn.Linear(hiddenDim, nClasses - 1))
                    self.tcn_a.append(nn.Linear(hiddenDim, 1))
            else:
                self.tcn_v.append(nn.Linear(hiddenDim, 1))
                self.tcn_a.append(nn.Linear(hiddenDim, 1))
        _init_like_reference(self)

    def forward_bf16(self, video, se, au, normalise=False, after_features=None):
        if fp32.enabled():       # fp32-parity inference mode (m3t_b200.fp32)
            fp32.require_eval(self)
            cl = lambda f: f.float().transpose(1, 2).contiguous()        # noqa: E731  (B,C,T) -> (B,T,C)
            x = fp32.vggm_stack(self.shared, video, True, normalise)
            if self.split_layer == 5:
                f = torch.cat((_squeeze_features(x), cl(se), cl(au)), dim=-1)
                return self.gru.forward_bf16(f) if self.backend == 'gru' else f
            x_v = torch.cat((_squeeze_features(fp32.vggm_stack(self.v_private, x, False)), cl(se)), dim=-1)
            x_a = torch.cat((_squeeze_features(fp32.vggm_stack(self.a_private, x, False)), cl(au)), dim=-1)
            if self.backend == 'gru':
                o_v, o_a = self.gru_v.forward_bf16(x_v), self.gru_a.forward_bf16(x_a)
            else:
                o_v, o_a = fp32.tcn_simple(self.tcn_v, x_v), fp32.tcn_simple(self.tcn_a, x_a)
            return torch.cat((o_v, o_a), dim=-1)
        xs = raw.video_prep_s2d(video.contiguous(), normalise)
        x = _run_stack(self.shared, xs, True)
        if self.split_layer == 5:
            f = torch.cat((_squeeze_features(x), ops.ToCL.apply(se), ops.ToCL.apply(au)), dim=-1)
            if after_features is not None:
                after_features()
            return self.gru.forward_bf16(f) if self.backend == 'gru' else f
        x_v = _cat_cl(_squeeze_features(_run_stack(self.v_private, x, False)), se)
        x_a = _cat_cl(_squeeze_features(_run_stack(self.a_private, x, False)), au)
        if after_features is not None:
            after_features()
        if self.backend == 'gru' and streams.overlap_ok(x_a):
            # inference: the two private BiGRU heads are independent -> the arousal head runs on a side stream
            o_a = streams.run_on_side(1, x_a.device, self.gru_a.forward_bf16, x_a)
            o_v = self.gru_v.forward_bf16(x_v)
            streams.join(x_a.device, 1)
        elif self.backend == 'gru':
            o_v, o_a = self.gru_v.forward_bf16(x_v), self.gru_a.forward_bf16(x_a)
        else:
            o_v, o_a = _tcn_simple(self.tcn_v, x_v), _tcn_simple(self.tcn_a, x_a)
        if o_v.dtype != o_a.dtype:
            o_v, o_a = ops.as_f32(o_v), ops.as_f32(o_a)
        return torch.cat((o_v, o_a), dim=-1)

    def forward(self, x, se, au):
        return ops.as_f32(self.forward_bf16(x, se, au))
