"""Audio branch of BASELINE config 2: "ResNet over log-Mel windows + TCN head".

NOT a reference class: the reference's audio stream is a BiGRU over 200-d stacked log-Mel vectors
(models/model.py:86).  This is the builder-declared composition of reference pieces that config 2
names, and its parity oracle is the same composition in plain PyTorch (tests/gpu_cases.py::case_audio_resnet):

    audio (B,T,200) = T windows of 5 log-Mel frames x 40 bins (models/dataset.py:83-95)
      -> view (B*T, 1, 5, 40) -> Conv2d(1,64,3,1,1,bias=False) + BatchNorm2d(64) + ReLU          [new 2-D stem]
      -> ResNet(BasicBlock,[2,2,2,2], agg_mode='ap')  (models/resnet.py:59)       -> (B, T, 512)
      -> TemporalConvNet(512,[512,512],3) over time   (models/tcn.py:49)          -> (B, T, 512)
      -> Linear(512, 2)                                                            -> (B, T, 2)  valence / arousal
"""
import torch.nn as nn

from .. import ops
from .resnet import BasicBlock, ResNet
from .tcn import TemporalConvNet


class AudioResNetTCN(nn.Module):
    def __init__(self, n_mels=40, context=5, hidden=512, levels=2, kernel_size=3, dropout=0.2, n_out=2):
        super().__init__()
        self.n_mels, self.context = n_mels, context
        self.stem = nn.Sequential(nn.Conv2d(1, 64, 3, 1, 1, bias=False), nn.BatchNorm2d(64), nn.ReLU(True))
        nn.init.kaiming_normal_(self.stem[0].weight, mode='fan_out', nonlinearity='relu')
        self.resnet = ResNet(BasicBlock, [2, 2, 2, 2], hidden, zero_init_residual=True, agg_mode='ap')
        self.tcn = TemporalConvNet(hidden, [hidden] * levels, kernel_size, dropout=dropout)
        self.fc = nn.Linear(hidden, n_out)

    def forward(self, audio):
        B, T, D = audio.shape
        assert D == self.n_mels * self.context
        x = audio.reshape(B * T, self.context, self.n_mels).float()
        conv, bn = self.stem[0], self.stem[1]
        h = ops.Conv3x3C1BNReLU.apply(x, conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.training)
        if bn.training:
            ops.bump_num_batches_tracked(bn)
        f = self.resnet.forward_cl(h).view(B, T, -1)          # (B,T,512) bf16
        f = self.tcn.forward_cl(f)
        return ops.linear(f, self.fc.weight, self.fc.bias, out_f32=True)
