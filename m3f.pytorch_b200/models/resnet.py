"""Stem-less ResNet trunks (reference: models/resnet.py:18-124 v1 BasicBlock / ResNet; :127-251 pre-activation
BasicBlockV2 / ResNetV2).

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
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride
        if use_cbam:
            from .cbam import CBAM
            self.cbam = CBAM(planes)
        else:
            self.cbam = None

    def forward_cl(self, x):
        if self.cbam is not None:
            # reference :48-49: the attention block sits between bn2 and the residual add
            h = _conv_bn_act(x, self.conv1, self.bn1, None, True)
            idt = x if self.downsample is None else _conv_bn_act(x, self.downsample[0], self.downsample[1], None, False)
            out = self.cbam.forward_cl(_conv_bn_act(h, self.conv2, self.bn2, None, False))
            return ops.AddReLU.apply(out, idt)
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


def _bn_act(x, bn, relu=True):
    out = ops.BNActFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, relu, bn.training, bn.momentum)
    if bn.training:
        ops.bump_num_batches_tracked(bn)
    return out


class BasicBlockV2(nn.Module):
    """Pre-activation block (reference models/resnet.py:127-179): [bn1 -> relu] -> conv1 -> bn2 -> relu -> conv2, plus the
    identity (or a bare 1x1 `downsample` conv of the PRE-ACTIVATED input).  Units: ops.BNActFn (stand-alone BN + ReLU),
    ops.ConvBNAct (conv1 + bn2 + relu, statistics from the conv epilogue), ops.ConvPlainFn (conv2 with the residual
    add in its epilogue; the downsample)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, is_first_block_of_first_layer=False,
                 use_cbam=False):
        super().__init__()
        self.is_first_block_of_first_layer = is_first_block_of_first_layer
        if not is_first_block_of_first_layer:
            self.bn1 = nn.BatchNorm2d(inplanes)
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.relu = nn.ReLU(True)
        self.downsample = downsample
        self.stride = stride
        if use_cbam:
            from .cbam import CBAM
            self.cbam = CBAM(planes)
        else:
            self.cbam = None

    def forward_cl(self, x):
        out = x if self.is_first_block_of_first_layer else _bn_act(x, self.bn1)
        if self.downsample is not None:
            identity = ops.ConvPlainFn.apply(out, self.downsample.weight, None, self.downsample.stride[0], 0)
        else:
            identity = x
        h = _conv_bn_act(out, self.conv1, self.bn2, None, True)
        if self.cbam is None:
            return ops.ConvPlainFn.apply(h, self.conv2.weight, identity, 1, 1)
        y = ops.ConvPlainFn.apply(h, self.conv2.weight, None, 1, 1)
        return ops.AddFn.apply(self.cbam.forward_cl(y), identity)

    def forward(self, x):
        return ops.FromCL.apply(self.forward_cl(ops.ToCL.apply(x)))


class ResNetV2(nn.Module):
    """Stem-less pre-activation ResNet (reference models/resnet.py:182-251): four stages, bn5 + relu5, pooling."""

    def __init__(self, block, layers, num_classes=256, zero_init_residual=False, agg_mode='ap', fmap_out_size=3,
                 use_cbam=False):
        super().__init__()
        self.inplanes = 64
        self.agg_mode = agg_mode
        self.layer1 = self._make_layer(block, 64, layers[0], use_cbam=use_cbam)
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2, use_cbam=use_cbam)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2, use_cbam=use_cbam)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2, use_cbam=use_cbam)
        self.bn5 = nn.BatchNorm2d(self.inplanes)
        self.relu5 = nn.ReLU(True)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(512 * fmap_out_size * fmap_out_size, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        with torch.no_grad():      # the reference zeroes the BatchNorm weight inside every CBAM spatial gate (:206-210)
            for key, val in self.state_dict().items():
                if key.split('.')[-1] == 'weight' and 'bn' in key and 'SpatialGate' in key:
                    val.zero_()
        if zero_init_residual:     # as the reference (:215-218); its first block has no bn1, so this raises there too
            for m in self.modules():
                if isinstance(m, BasicBlockV2):
                    nn.init.zeros_(m.bn1.weight)

    def _make_layer(self, block, planes, blocks, stride=1, use_cbam=False):
        down = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            down = nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False)
        stack = [block(self.inplanes, planes, stride, down, stride == 1, use_cbam=use_cbam)]
        self.inplanes = planes * block.expansion
        stack += [block(self.inplanes, planes, use_cbam=use_cbam) for _ in range(1, blocks)]
        return nn.Sequential(*stack)

    def forward_cl(self, x):
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in layer:
                x = blk.forward_cl(x)
        x = _bn_act(x, self.bn5)
        # the reference pools twice when agg_mode == 'ap' (:244-246): AdaptiveAvgPool2d(1) is idempotent
        f = ops.AvgPoolCL.apply(x)
        if self.agg_mode == 'fc':
            return ops.linear(f, self.fc.weight, self.fc.bias)
        return f

    def forward(self, x):
        if fp32.enabled():
            raise NotImplementedError("the fp32-parity mode covers the v1 trunk (the only one AffWild2VA builds)")
        return ops.as_f32(self.forward_cl(ops.ToCL.apply(x)))
