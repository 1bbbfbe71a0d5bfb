"""Diagnostic (not a pytest file): where does the fused TemporalBlock's dropout mask land?"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from m3t_b200 import ops, raw
from oracle import dropout as OD

torch.manual_seed(0)
B, T, C = 3, 20, 512
x = torch.randn(B, T, C).abs().bfloat16().cuda() + 0.5
v = torch.randn(C, C, 3) * 0.02
v[:, :, :] = 0
for i in range(C):
    v[i, i, 2] = 1.0          # identity on the current time step: conv(x) = x (> 0 everywhere)
g = v.flatten(1).norm(dim=1).view(C, 1, 1).clone()
b = torch.zeros(C)
torch.manual_seed(77)
seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
torch.manual_seed(77)
y = ops.TCNConvFn.apply(x, v.cuda(), g.cuda(), b.cuda(), None, 1, 2, 0.25, True)
yz = (y.float().cpu().numpy() == 0)
print("zero fraction", yz.mean())
keep = OD.keep_mask(seed, B * T * C, 0.25)
for name, m in (("row-major (b,t,c)", ~keep.reshape(B, T, C)),
                ("transposed", ~keep.reshape(C, B * T).T.reshape(B, T, C))):
    print(name, "mismatch fraction", float((m != yz).mean()))
# stand-alone kernel on the same tensor and seed
y2 = raw.dropout_bf16(x.contiguous(), 0.25, seed)
print("stand-alone vs oracle", float(((y2.float().cpu().numpy() == 0) != ~keep.reshape(B, T, C)).mean()))
print("fused vs stand-alone values", float((y.float() - y2.float()).abs().max()))
for cand in range(0, 8):
    k2 = OD.keep_mask(seed >> cand, B * T * C, 0.25)
    print("seed >>", cand, float(((~k2.reshape(B, T, C)) != yz).mean()))
k3 = OD.keep_mask(seed & 0xFFFFFFFF, B * T * C, 0.25)
print("low 32 bits of seed", float(((~k3.reshape(B, T, C)) != yz).mean()))
