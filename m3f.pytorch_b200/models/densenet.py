"""3-D DenseNet-52 trunk (reference: models/densenet.py:5-93; `--backbone densenet` through VA_3DDenseNet).

Same module tree and state_dict keys (`features.denseblock{b}.denselayer{l}.{norm1,conv1,norm2,conv2}.*`,
`features.transition{b}.{norm,conv}.*`, `features.norm5.*`, `fc.*`).  On channels-last bf16 (N,D,H,W,C):
  * BatchNorm3d + ReLU ahead of every convolution = ops.BNActFn (stand-alone statistics pass + one apply pass);
  * the 1x1x1 bottleneck / transition convolutions are plain GEMMs over the pixels (ops.linear: K = 64 + 32 i input
    channels need no padding, the tensor map zero-fills the last K block);
  * the 3x3x3 growth convolution (128 -> 32) runs on the implicit-GEMM kernel with its filter bank zero-padded to 64
    output channels (ops.Conv3dPlainFn); the 32 new feature maps are the first half of its output;
  * feature concatenation is torch.cat on the channel axis; the transitions' AvgPool3d((1,2,2)) = m3t_avgpool2x2.
"""
import torch
import torch.nn as nn

from .. import ops
from .resnet import _bn_act


class _DenseLayer_3D(nn.Sequential):
    def __init__(self, num_input_features, growth_rate, bn_size, drop_rate):
        super().__init__()
        self.add_module('norm1', nn.BatchNorm3d(num_input_features))
        self.add_module('relu1', nn.ReLU(inplace=True))
        self.add_module('conv1', nn.Conv3d(num_input_features, bn_size * growth_rate, kernel_size=1, stride=1,
                                           bias=False))
        self.add_module('norm2', nn.BatchNorm3d(bn_size * growth_rate))
        self.add_module('relu2', nn.ReLU(inplace=True))
        self.add_module('conv2', nn.Conv3d(bn_size * growth_rate, growth_rate, kernel_size=3, stride=1, padding=1,
                                           bias=False))
        self.add_module('dp', nn.Dropout3d(p=drop_rate))

    def forward_cl(self, x):
        if self.training and self.dp.p > 0:
            raise NotImplementedError("Dropout3d > 0 (the reference builds DenseNet52_3D with dp = 0)")
        N, D, H, W, Cin = x.shape
        h = _bn_act(x, self.norm1)
        mid = self.conv1.out_channels
        h = ops.linear(h.view(-1, Cin), self.conv1.weight.view(mid, Cin), None).view(N, D, H, W, mid)
        h = _bn_act(h, self.norm2)
        w2 = self.conv2.weight
        grow = w2.shape[0]
        gpad = (grow + 63) // 64 * 64
        if gpad != grow:
            w2 = torch.cat((w2, w2.new_zeros((gpad - grow,) + tuple(w2.shape[1:]))), dim=0)
        new = ops.Conv3dPlainFn.apply(h, w2, 1)
        return torch.cat((x, new[..., :grow]), dim=-1)


class _DenseBlock_3D(nn.Sequential):
    def __init__(self, num_layers, num_input_features, bn_size, growth_rate, drop_rate):
        super().__init__()
        for i in range(num_layers):
            self.add_module('denselayer%d' % (i + 1),
                            _DenseLayer_3D(num_input_features + i * growth_rate, growth_rate, bn_size, drop_rate))

    def forward_cl(self, x):
        for layer in self:
            x = layer.forward_cl(x)
        return x


class _Transition_3D(nn.Sequential):
    def __init__(self, num_input_features, num_output_features, pooling=True):
        super().__init__()
        self.add_module('norm', nn.BatchNorm3d(num_input_features))
        self.add_module('relu', nn.ReLU(inplace=True))
        self.add_module('conv', nn.Conv3d(num_input_features, num_output_features, kernel_size=1, stride=1, bias=False))
        if pooling:
            self.add_module('pool', nn.AvgPool3d(kernel_size=(1, 2, 2), stride=(1, 2, 2)))

    def forward_cl(self, x):
        N, D, H, W, Cin = x.shape
        h = _bn_act(x, self.norm)
        Cout = self.conv.out_channels
        h = ops.linear(h.view(-1, Cin), self.conv.weight.view(Cout, Cin), None).view(N, D, H, W, Cout)
        if hasattr(self, 'pool'):
            h = ops.AvgPool2x2Fn.apply(h.view(N * D, H, W, Cout))
            h = h.view(N, D, h.shape[1], h.shape[2], Cout)
        return h


class DenseNet52_3D(nn.Module):
    def __init__(self, num_classes=256, num_init_features=64, bn_size=4, block_config=(4, 6, 8, 6), growth_rate=32,
                 dp=0., agg_mode='ap', fmap_out_size=3):
        super().__init__()
        num_features = num_init_features
        self.agg_mode = agg_mode
        self.features = nn.Sequential()
        for i, num_layers in enumerate(block_config):
            self.features.add_module('denseblock%d' % (i + 1),
                                     _DenseBlock_3D(num_layers, num_features, bn_size, growth_rate, dp))
            num_features = num_features + num_layers * growth_rate
            if i != len(block_config) - 1:
                pooling = not (len(block_config) > 4 and i == 1)
                self.features.add_module('transition%d' % (i + 1),
                                         _Transition_3D(num_features, num_features // 2, pooling=pooling))
                num_features = num_features // 2
        self.features.add_module('norm5', nn.BatchNorm3d(num_features))
        self.features.add_module('relu5', nn.ReLU())
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(num_features * fmap_out_size * fmap_out_size, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight)
            elif isinstance(m, nn.BatchNorm3d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                nn.init.constant_(m.bias, 0)

    def forward_cl(self, x):
        """x: CL bf16 (B,T,H,W,64) -> bf16 (B,T,C_out) for agg_mode 'ap'."""
        for name, mod in self.features.named_children():
            if name == 'norm5':
                x = _bn_act(x, mod)
            elif name != 'relu5':
                x = mod.forward_cl(x)
        if self.agg_mode != 'ap':
            raise NotImplementedError("agg_mode 'fc' flattens a 3x3 map the 112-pixel pipeline never produces")
        B, T, H, W, C = x.shape
        return ops.AvgPoolCL.apply(x.view(B * T, H, W, C)).view(B, T, C)

    def forward(self, x):
        # reference layout: (B,C,T,H,W) fp32 -> (B,T,C_out) fp32
        return ops.as_f32(self.forward_cl(ops.ToCL.apply(x)))
