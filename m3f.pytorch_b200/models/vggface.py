"""VGG-Face descriptor network for 112 x 112 crops (reference: models/vggface.py:7-50; `--backbone vggface`).

Same module tree and state_dict keys (`conv{1-5}.convs.{i}.{weight,bias}`, `fc1.*`).  Every Conv2d(3x3, pad 1) + ReLU is
one implicit-GEMM launch with bias and ReLU in the epilogue (ops.Conv2dBiasAct; the 3 input channels of conv1 are
zero-padded to one 64-channel block), max_pool2d(2, 2, ceil_mode=True) is the 2x2 pooling kernel (the odd 7 x 7 map of
block 5 is zero-padded to 8 x 8: the pooled values are post-ReLU, so the padding never wins), fc1 + ReLU is a GEMM
epilogue, Dropout(0.5) the counter-based dropout kernel.
"""
import torch.nn as nn

from .. import ops


class _ConvBlock(nn.Module):
    def __init__(self, *units):
        super().__init__()
        self.convs = nn.ModuleList([nn.Conv2d(i, o, 3, 1, 1) for i, o in zip(units[:-1], units[1:])])

    def forward_cl(self, x):
        for c in self.convs:
            x = ops.Conv2dBiasAct.apply(x, c.weight, c.bias, 1, True)
        return ops.MaxPool2x2Ceil.apply(x)


class VGGFace(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = _ConvBlock(3, 64, 64)
        self.conv2 = _ConvBlock(64, 128, 128)
        self.conv3 = _ConvBlock(128, 256, 256, 256)
        self.conv4 = _ConvBlock(256, 512, 512, 512)
        self.conv5 = _ConvBlock(512, 512, 512, 512)
        self.dropout = nn.Dropout(0.5)
        self.fc1 = nn.Linear(4 * 4 * 512, 4096)

    def forward_bf16(self, x):
        """x: fp32 (N,3,H,W) -> bf16 (N,4096)."""
        from .. import raw
        h = ops.ToCLPad.apply(x, 64)                                  # (N,H,W,64), channels 3..63 zero
        for blk in (self.conv1, self.conv2, self.conv3, self.conv4, self.conv5):
            h = blk.forward_cl(h)
        # the reference flattens NCHW (`x.view(N, -1)`, models/vggface.py:26): channel-major feature order
        flat = h.permute(0, 3, 1, 2).reshape(h.shape[0], -1)
        f = ops.linear(flat, self.fc1.weight, self.fc1.bias, relu=True)
        if self.training and self.dropout.p > 0:
            f = ops.DropoutFn.apply(f, self.dropout.p)
        return f

    def forward(self, x):
        return ops.as_f32(self.forward_bf16(x))
