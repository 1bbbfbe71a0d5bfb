"""Achieved HBM bandwidth of the streaming temporal / fusion kernels at a size that fills the machine (the bench-size
launches move ~10 MB and are latency-bound): attention mix forward / backward, bf16 add, BN passes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from m3t_b200 import raw  # noqa: E402


def timed(fn, it=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(it):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / it


rows, C = 1 << 19, 512
xa = torch.randn(rows, C, device="cuda").bfloat16()
xv = torch.randn(rows, C, device="cuda").bfloat16()
sa = torch.randn(rows, device="cuda")
sv = torch.randn(rows, device="cuda")
df = torch.randn(rows, C, device="cuda").bfloat16()
ms = timed(lambda: raw.att_mix_fwd(xa, xv, sa, sv))
print("att_mix_fwd  %d x %d: %.3f ms  %.0f GB/s (3 x 2C + 8 B per row)" % (rows, C, ms, rows * (6 * C + 8) / ms / 1e6))
ms = timed(lambda: raw.att_mix_bwd(df, xa, xv, sa, sv))
print("att_mix_bwd  %d x %d: %.3f ms  %.0f GB/s (5 x 2C + 16 B per row)" % (rows, C, ms, rows * (10 * C + 16) / ms / 1e6))
ms = timed(lambda: raw.add_bf16(xa, xv))
print("add_bf16     %d x %d: %.3f ms  %.0f GB/s" % (rows, C, ms, rows * 6 * C / ms / 1e6))
