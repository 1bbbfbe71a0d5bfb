"""Visual backbones (reference: models/backbone.py).

VA_3DResNet  (reference :314-372): Conv3d spatio-temporal stem + per-frame ResNet-18 trunk + BiGRU head.
The NCDHW -> (B*T)CHW transpose/copy of the reference (:349-350) does not exist here: the stem writes
channels-last (B*T, 28, 28, 64) bf16, which *is* the trunk's input layout.
"""
import math

import torch.nn as nn

from .. import fp32, ops
from .resnet import BasicBlock, BasicBlockV2, ResNet, ResNetV2
from .rnn import GRU


def _init_like_reference(module):
    """Conv3d ~ N(0, sqrt(2/(kt*kh*kw*Cout))), Conv1d kaiming-normal, BN3d/BN1d gamma=1 beta=0
    (reference: models/backbone.py:357-372)."""
    for m in module.modules():
        if isinstance(m, nn.Conv3d):
            fan = m.kernel_size[0] * m.kernel_size[1] * m.kernel_size[2] * m.out_channels
            m.weight.data.normal_(0, math.sqrt(2.0 / fan))
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.Conv1d):
            nn.init.kaiming_normal_(m.weight)
        elif isinstance(m, (nn.BatchNorm3d, nn.BatchNorm1d)):
            m.weight.data.fill_(1)
            m.bias.data.zero_()


class VA_3DResNet(nn.Module):
    def __init__(self, inputDim=512, hiddenDim=512, nLayers=2, nClasses=2, frameLen=16, backend='gru', use_cbam=False,
                 resnet_ver='v2', resnet_depth=18, frontend_agg_mode='ap', nFCs=1):
        super().__init__()
        self.inputDim, self.hiddenDim, self.nClasses = inputDim, hiddenDim, nClasses
        self.frameLen, self.nLayers, self.backend, self.nFCs = frameLen, nLayers, backend, nFCs
        assert resnet_depth in (18, 34) and resnet_ver in ('v1', 'v2'), \
            'unsupported ResNet configuration: {}, {}'.format(resnet_depth, resnet_ver)
        self.c3d = nn.Sequential(
            nn.Conv3d(3, 64, kernel_size=(5, 7, 7), stride=(1, 2, 2), padding=(2, 3, 3), bias=False),
            nn.BatchNorm3d(64),
            nn.ReLU(True),
            nn.MaxPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1)))
        blocks = [2, 2, 2, 2] if resnet_depth == 18 else [3, 4, 6, 3]
        if resnet_ver == 'v2':      # reference :336-337 (the constructor's default; AffWild2VA passes 'v1')
            self.resnet = ResNetV2(BasicBlockV2, blocks, inputDim, zero_init_residual=False,
                                   agg_mode=frontend_agg_mode, fmap_out_size=3, use_cbam=use_cbam)
        else:
            self.resnet = ResNet(BasicBlock, blocks, inputDim, zero_init_residual=True, agg_mode=frontend_agg_mode,
                                 fmap_out_size=3, use_cbam=use_cbam)
        if backend == 'gru':
            self.gru = GRU(inputDim, hiddenDim, nLayers, nClasses, nFCs)
        _init_like_reference(self)

    def features_cl(self, video, normalise):
        """video (B,3,T,112,112) fp32/uint8 -> per-frame trunk features bf16 (B, T, 512)."""
        conv, bn = self.c3d[0], self.c3d[1]
        B, T = video.shape[0], video.shape[2]
        x = ops.Stem3D.apply(video, conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, normalise,
                             bn.training)
        if bn.training:
            ops.bump_num_batches_tracked(bn)
        f = self.resnet.forward_cl(x)                     # (B*T, 512)
        return f.view(B, T, -1)

    def forward_bf16(self, x, normalise=False, after_features=None):
        """after_features: optional callable invoked once the convolutions are enqueued and before the recurrent head
        (AffWild2VA forks the audio stream there, streams.py)."""
        if fp32.enabled():       # fp32-parity inference mode: float32 activations, split-operand tensor-core launches
            fp32.require_eval(self)
            return fp32.va_3dresnet(self, x, normalise)
        f = self.features_cl(x, normalise)
        if after_features is not None:
            after_features()
        if f.shape[1] != self.frameLen:
            raise RuntimeError("VA_3DResNet: T (%d) must equal frameLen (%d)" % (f.shape[1], self.frameLen))
        if self.backend == 'gru':
            return self.gru.forward_bf16(f)
        return f

    def forward(self, x, *unused):
        # `*unused` tolerates AffWild2VA.forward's 3-argument call (reference models/model.py:111,130)
        return ops.as_f32(self.forward_bf16(x))


class VA_VGGFace(nn.Module):
    """Per-frame VGG-Face descriptor + BiGRU head (reference models/backbone.py:16-58; `--backbone vggface`)."""

    def __init__(self, inputDim=4096, hiddenDim=512, nLayers=2, nClasses=2, frameLen=16, backend='gru', nFCs=1):
        super().__init__()
        from .vggface import VGGFace
        self.inputDim, self.hiddenDim, self.nClasses = inputDim, hiddenDim, nClasses
        self.frameLen, self.nLayers, self.backend, self.nFCs = frameLen, nLayers, backend, nFCs
        self.vgg = VGGFace()
        if backend == 'gru':
            self.gru = GRU(inputDim, hiddenDim, nLayers, nClasses, nFCs)
        for m in self.modules():            # reference :45-58
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_normal_(m.weight.data)
                if m.bias is not None:
                    nn.init.normal_(m.bias.data)
            elif isinstance(m, nn.Linear):
                nn.init.xavier_normal_(m.weight)
                nn.init.constant_(m.bias, 0)

    def forward_bf16(self, x, *unused, normalise=False, after_features=None):
        B, T = x.shape[0], x.shape[2]
        if normalise:
            x = (x.float() - 127.5) / 127.5
        frames = x.transpose(1, 2).contiguous().view(-1, x.shape[1], x.shape[3], x.shape[4])
        f = self.vgg.forward_bf16(frames).view(B, T, -1)
        if after_features is not None:
            after_features()
        return self.gru.forward_bf16(f) if self.backend == 'gru' else f

    def forward(self, x, *unused):
        return ops.as_f32(self.forward_bf16(x))


class VA_3DDenseNet(nn.Module):
    """Conv3d stem + 3-D DenseNet-52 + BiGRU head (reference models/backbone.py:375-420; `--backbone densenet`)."""

    def __init__(self, inputDim=392, hiddenDim=512, nLayers=2, nClasses=2, frameLen=16, backend='gru',
                 frontend_agg_mode='ap', nFCs=1):
        super().__init__()
        from .densenet import DenseNet52_3D
        self.inputDim, self.hiddenDim, self.nClasses = inputDim, hiddenDim, nClasses
        self.frameLen, self.nLayers, self.backend, self.nFCs = frameLen, nLayers, backend, nFCs
        self.c3d = nn.Sequential(
            nn.Conv3d(3, 64, kernel_size=(5, 7, 7), stride=(1, 2, 2), padding=(2, 3, 3), bias=False),
            nn.BatchNorm3d(64),
            nn.ReLU(True),
            nn.MaxPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1)))
        self.densenet = DenseNet52_3D(inputDim, agg_mode=frontend_agg_mode, fmap_out_size=3)
        if backend == 'gru':
            self.gru = GRU(inputDim, hiddenDim, nLayers, nClasses, nFCs)
        _init_like_reference(self)

    def forward_bf16(self, x, *unused, normalise=False, after_features=None):
        conv, bn = self.c3d[0], self.c3d[1]
        B, T = x.shape[0], x.shape[2]
        h = ops.Stem3D.apply(x, conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, normalise,
                             bn.training)
        if bn.training:
            ops.bump_num_batches_tracked(bn)
        f = self.densenet.forward_cl(h.view(B, T, h.shape[1], h.shape[2], h.shape[3]))
        if after_features is not None:
            after_features()
        return self.gru.forward_bf16(f) if self.backend == 'gru' else f

    def forward(self, x, *unused):
        return ops.as_f32(self.forward_bf16(x))
