"""Launch the hot kernels once each at bench sizes (for `ncu --set full`):
   ncu --set full --clock-control none --import-source on -o gpurun_out/prof_r1 python tests/prof_kernels.py [frames]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from m3t_b200 import raw  # noqa: E402


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    B, T = F // 16, 16
    dev = "cuda"
    rb = lambda *s: torch.randn(s, device=dev).bfloat16()    # noqa: E731
    # stem fprop + wgrad (W-unrolled space-to-depth input)
    xs = rb(B, T, 56, 56, 64)
    xs[..., 48:] = 0
    w = (torch.randn(64, 1280, device=dev) * 0.03).bfloat16()
    g = raw.conv_geom(3, B, T, 56, 56, 64, 64, (5, 4, 1), (1, 1, 1), (2, 2, 0), (2, 1, 0), (1, 1, 1))
    st = torch.zeros(2, 64, device=dev)
    y = raw.stem_fprop_halo(xs, w, stats=st)                 # halo kernels (product path)
    raw.wgrad_stem_halo(xs, y)
    y = raw.conv_fprop(xs, w, g, stats=st, tag="stem")       # generic im2col kernels, for comparison
    raw.conv_wgrad(xs, y, g, splits=59)
    # stem tail fwd/bwd
    ss = torch.rand(4, 64, device=dev) + 0.5
    y4 = y.view(F, 56, 56, 64)
    out, idx = raw.bn_relu_maxpool(y4, ss[2], ss[3], True)
    raw.maxpool_bn_bwd(out, idx, y4, ss[0], ss[1], ss[2], ss[3], F * 3136)
    del xs, y, y4
    # layer1: halo fprop, generic fprop, wgrad
    x = rb(F, 28, 28, 64)
    w1 = (torch.randn(64, 576, device=dev) * 0.05).bfloat16()
    g1 = raw.conv_geom(2, F, 1, 28, 28, 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), (0, 1, 1), (1, 1, 1))
    y1 = raw.conv_fprop(x, w1, g1, stats=st)
    raw.USE_HALO = False
    raw.conv_fprop(x, w1, g1, stats=st)
    raw.USE_HALO = True
    raw.conv_wgrad(x, y1.view(F, 28, 28, 64), g1)             # halo wgrad
    raw.conv_wgrad(x, y1.view(F, 28, 28, 64), g1, splits=59)  # generic
    o = raw.bn_act(y1, ss[2], ss[3], res=x, relu=True)
    sums, dz = raw.bn_bwd_reduce(o, o, y1, ss[0], ss[1], True, True)
    raw.bn_bwd_apply(o, o, y1, ss[0], ss[1], ss[2], sums, F * 784, True, shift=ss[3])
    del x, y1, o, dz
    # layer2 / layer4 fprop
    x = rb(F, 14, 14, 128)
    w2 = (torch.randn(128, 1152, device=dev) * 0.03).bfloat16()
    g2 = raw.conv_geom(2, F, 1, 14, 14, 128, 128, (1, 3, 3), (1, 1, 1), (0, 1, 1), (0, 1, 1), (1, 1, 1))
    y2 = raw.conv_fprop(x, w2, g2, stats=torch.zeros(2, 128, device=dev))          # image-per-tile halo kernel
    raw.USE_HALO128 = False
    raw.conv_fprop(x, w2, g2, stats=torch.zeros(2, 128, device=dev))               # generic im2col, for comparison
    raw.USE_HALO128 = True
    raw.conv_wgrad(x, y2.view(F, 14, 14, 128), g2)
    # stride-2 dgrad of layer2.0.conv1 by parity (4 launches) into a 28x28x64 dX
    from m3t_b200 import ops
    wfull = torch.randn(128, 64, 3, 3, device=dev) * 0.05
    ops.conv2d_dgrad(y2.view(F, 14, 14, 128), wfull, (F, 28, 28, 64), 2, 1)
    # layer3 / layer4 3x3 on the persistent tile walker, and their weight gradients (256-row tiles)
    for HW, C in ((7, 256), (4, 512)):
        x = rb(F, HW, HW, C)
        w = (torch.randn(C, 9 * C, device=dev) * 0.02).bfloat16()
        g = raw.conv_geom(2, F, 1, HW, HW, C, C, (1, 3, 3), (1, 1, 1), (0, 1, 1), (0, 1, 1), (1, 1, 1))
        y = raw.conv_fprop(x, w, g, stats=torch.zeros(2, C, device=dev))
        raw.conv_wgrad(x, y.view(F, HW, HW, C), g)
    # GRU layer H=512 forward/backward, x-projection GEMM
    xg = rb(B * T, 512)
    wih = (torch.randn(3072, 512, device=dev) * 0.03).bfloat16()
    gi = raw.gemm(xg, wih, out_dtype=torch.float32)
    whh = (torch.randn(2, 1536, 512, device=dev) * 0.03).bfloat16()
    bhh = torch.zeros(2, 1536, device=dev)
    outg, _, saved = raw.gru_fwd(gi, whh, bhh, B, T, 512, True)
    raw.gru_bwd(outg, outg, saved, whh.transpose(1, 2).contiguous(), B, T, 512)
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
