"""Diagnostic (not a pytest file): where does the fused TemporalBlock's dropout mask land?"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from m3t_b200 import ops, raw
from oracle import dropout as OD

for (B, T, C, p, dil) in ((3, 20, 512, 0.25, 1), (6, 40, 512, 0.2, 2), (3, 17, 512, 0.5, 1)):
    torch.manual_seed(0)
    x = torch.randn(B, T, C).abs().bfloat16().cuda() + 0.5
    v = torch.zeros(C, C, 3)
    for i in range(C):
        v[i, i, 2] = 1.0          # identity on the current time step: conv(x) = x (> 0 everywhere)
    g = v.flatten(1).norm(dim=1).view(C, 1, 1).clone()
    b = torch.zeros(C)
    torch.manual_seed(77)
    seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
    torch.manual_seed(77)
    y = ops.TCNConvFn.apply(x, v.cuda(), g.cuda(), b.cuda(), None, dil, 2 * dil, p, True)
    yz = (y.float().cpu().numpy() == 0).reshape(B * T, C)
    keep = OD.keep_mask(seed, B * T * C, p).reshape(B * T, C)
    mm = (yz != ~keep)
    print(B, T, C, p, "zero fraction", yz.mean(), "mismatch", mm.mean(), "by 64-row band:",
          [round(float(mm[r:r + 64].mean()), 3) for r in range(0, B * T, 64)],
          "by 128-col band:", [round(float(mm[:, c:c + 128].mean()), 3) for c in range(0, C, 128)])
    # with a residual (second conv of the block): y = relu(t + res), t returned through ctx
    res = torch.zeros_like(x)
    torch.manual_seed(77)
    y2 = ops.TCNConvFn.apply(x, v.cuda(), g.cuda(), b.cuda(), res, dil, 2 * dil, p, True)
    print("   with residual: equal to the no-residual output:", bool((y2 == y).all()))
