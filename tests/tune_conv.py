"""Timing sweep of the implicit-GEMM conv kernels over the trunk geometries (run on the GPU box):
   python tests/tune_conv.py [frames]
Prints ms and algorithmic TFLOP/s per (geometry, variant); used to pick tile shapes / split counts."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from m3t_b200 import raw  # noqa: E402


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    geoms = [("l1 28x28 64->64 s1", 28, 64, 64, 3, 1, 1), ("l2 28->14 64->128 s2", 28, 64, 128, 3, 2, 1),
             ("l2 14x14 128->128 s1", 14, 128, 128, 3, 1, 1), ("l3 14->7 128->256 s2", 14, 128, 256, 3, 2, 1),
             ("l3 7x7 256->256 s1", 7, 256, 256, 3, 1, 1), ("l4 7->4 256->512 s2", 7, 256, 512, 3, 2, 1),
             ("l4 4x4 512->512 s1", 4, 512, 512, 3, 1, 1), ("ds 28->14 64->128 1x1", 28, 64, 128, 1, 2, 0),
             # parity sub-convolutions of the stride-2 dgrads (kernel k over dY, pad 0 below / k-1 above)
             ("par 14x14 128->64 k1", 14, 128, 64, 1, 1, 0), ("par 14x14 128->64 k2", 14, 128, 64, 2, 1, 0),
             ("par 7x7 256->128 k2", 7, 256, 128, 2, 1, 0), ("par 4x4 512->256 k2", 4, 512, 256, 2, 1, 0)]
    for name, HW, Cin, Cout, k, s, p in geoms:
        ph = k - 1 if name.startswith("par") else p
        g = raw.conv_geom(2, F, 1, HW, HW, Cin, Cout, (1, k, k), (1, s, s), (0, p, p), (0, ph, ph), (1, 1, 1))
        Z, P, Q = raw.conv_out_dims(g)
        x = torch.randn((F, HW, HW, Cin), device="cuda").bfloat16()
        w = (torch.randn((Cout, k * k * Cin), device="cuda") * 0.05).bfloat16()
        dy = torch.randn((F, P, Q, Cout), device="cuda").bfloat16()
        flops = 2.0 * F * P * Q * Cout * Cin * k * k
        res = []
        for hint, label in ((0, "auto"), (32 | 1, "mt1"), (32 | 2, "mt2"), (32 | 8, "bn128"), (32, "nohalo-auto"),
                            (16, "persist"), (16 | 1, "persist-mt1"), (16 | 2, "persist-mt2"), (16 | 8, "persist-bn128")):
            if (hint & 8) and Cout % 256:
                continue
            try:
                ms = timeit(lambda: raw.conv_fprop(x, w, g, tile_hint=hint))
                res.append("%s %.3fms %.0fTF" % (label, ms, flops / ms / 1e9))
            except Exception as e:  # noqa: BLE001
                res.append("%s ERR %s" % (label, str(e)[:40]))
            raw.USE_HALO = True
        print("fprop %-24s %s" % (name, " | ".join(res)), flush=True)
        res = []
        for splits, label in ((0, "auto"), (1 << 30, "mt1-auto"), (16, "s16"), (32, "s32"), (64, "s64"), (128, "s128")):
            try:
                ms = timeit(lambda: raw.conv_wgrad(x, dy, g, splits=splits))
                res.append("%s %.3fms %.0fTF" % (label, ms, flops / ms / 1e9))
            except Exception as e:  # noqa: BLE001
                res.append("%s ERR %s" % (label, str(e)[:40]))
        print("wgrad %-24s %s" % (name, " | ".join(res)), flush=True)
        del x, w, dy


if __name__ == "__main__":
    main()
