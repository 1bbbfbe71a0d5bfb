"""Stem-less ResNet trunk (reference: models/resnet.py:18-124, v1 BasicBlock / ResNet).

Parameters live in ordinary nn.Conv2d / nn.BatchNorm2d containers so that init distributions and state_dict keys
(`layer{1-4}.{i}.conv{1,2}.weight`, `.bn{1,2}.*`, `.downsample.{0.weight,1.*}`, `fc.*`) are the reference's; the
arithmetic runs as fused conv+BN(+residual)+ReLU units on channels-last bf16 activations (ops.ConvBNAct).
"""
import torch
import torch.nn as nn

from .. import fp32, ops


def _conv_bn_act(x, conv, bn, residual, relu):
    training = bn.training
    out = ops.ConvBNAct.apply(x, conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, residual,
                              conv.stride[0], conv.padding[0], relu, training)
    if training:
        ops.bump_num_batches_tracked(bn)
    return out


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, use_cbam=False):
        super().__init__()
        if use_cbam:
            raise NotImplementedError("CBAM is off at every call site of the reference (models/backbone.py:315)")
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride
        self.cbam = None

    def forward_cl(self, x):
        if self.training and torch.is_grad_enabled():
            # one autograd node per block: the gradients meeting at the block input are summed in a dgrad epilogue
            ds = self.downsample
            dargs = ((ds[0].weight, ds[1].weight, ds[1].bias, ds[1].running_mean, ds[1].running_var)
                     if ds is not None else (None,) * 5)
            out = ops.BasicBlockFn.apply(x, self.stride, self.conv1.weight, self.bn1.weight, self.bn1.bias,
                                         self.bn1.running_mean, self.bn1.running_var, self.conv2.weight,
                                         self.bn2.weight, self.bn2.bias, self.bn2.running_mean, self.bn2.running_var,
                                         *dargs)
            for bn in (self.bn1, self.bn2) + ((ds[1],) if ds is not None else ()):
                ops.bump_num_batches_tracked(bn)
            return out
        h = _conv_bn_act(x, self.conv1, self.bn1, None, True)
        idt = x if self.downsample is None else _conv_bn_act(x, self.downsample[0], self.downsample[1], None, False)
        return _conv_bn_act(h, self.conv2, self.bn2, idt, True)

    def forward(self, x):
        return ops.FromCL.apply(self.forward_cl(ops.ToCL.apply(x)))


class ResNet(nn.Module):
    def __init__(self, block, layers, num_classes=256, zero_init_residual=True, agg_mode='ap', fmap_out_size=3,
                 use_cbam=False):
        super().__init__()
        self.inplanes = 64
        self.agg_mode = agg_mode
        widths, strides = (64, 128, 256, 512), (1, 2, 2, 2)
        for i, (wd, st, n) in enumerate(zip(widths, strides, layers), start=1):
            setattr(self, 'layer%d' % i, self._make_layer(block, wd, n, st, use_cbam))
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(512 * fmap_out_size * fmap_out_size, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        if zero_init_residual:
            for m in self.modules():
                if isinstance(m, BasicBlock):
                    nn.init.zeros_(m.bn2.weight)

    def _make_layer(self, block, planes, blocks, stride=1, use_cbam=False):
        out_ch = planes * block.expansion
        down = None
        if stride != 1 or self.inplanes != out_ch:
            down = nn.Sequential(nn.Conv2d(self.inplanes, out_ch, 1, stride, bias=False), nn.BatchNorm2d(out_ch))
        stack = [block(self.inplanes, planes, stride, down, use_cbam=use_cbam)]
        self.inplanes = out_ch
        stack += [block(out_ch, planes, use_cbam=use_cbam) for _ in range(1, blocks)]
        return nn.Sequential(*stack)

    def forward_cl(self, x):
        """x: bf16 [F,H,W,64] -> bf16 [F,512] (agg_mode 'ap')."""
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in layer:
                x = blk.forward_cl(x)
        if self.agg_mode == 'ap':
            return ops.AvgPoolCL.apply(x)
        if self.agg_mode == 'fc':
            flat = ops.FromCL.apply(x).flatten(1)
            return ops.linear(flat, self.fc.weight, self.fc.bias)
        return x

    def forward(self, x):
        if fp32.enabled():
            fp32.require_eval(self)
            return fp32.resnet_trunk(self, x.permute(0, 2, 3, 1).contiguous().float())
        out = self.forward_cl(ops.ToCL.apply(x))
        if out.dim() == 2:
            return ops.as_f32(out)
        return ops.FromCL.apply(out)
