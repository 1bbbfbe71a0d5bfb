"""ncu target: the small-K convolutions on the one-tile-per-CTA and the persistent kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from m3t_b200 import raw  # noqa: E402

F = 4096
for name, HW, Cin, Cout, k, s, p, ph in (("ds", 28, 64, 128, 1, 2, 0, 0), ("par k1", 14, 128, 64, 1, 1, 0, 0),
                                          ("par k2", 14, 128, 64, 2, 1, 0, 1)):
    g = raw.conv_geom(2, F, 1, HW, HW, Cin, Cout, (1, k, k), (1, s, s), (0, p, p), (0, ph, ph), (1, 1, 1))
    x = torch.randn((F, HW, HW, Cin), device="cuda").bfloat16()
    w = (torch.randn((Cout, k * k * Cin), device="cuda") * 0.05).bfloat16()
    for hint in (32, 16):
        raw.conv_fprop(x, w, g, tile_hint=hint)
torch.cuda.synchronize()
print("done")
