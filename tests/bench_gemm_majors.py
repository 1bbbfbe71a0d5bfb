"""Micro-benchmark: tcgen05 GEMM throughput by operand major-ness (K-major vs MN-major), 4096^3 and a wgrad-like shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from m3t_b200 import raw  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for M, N, K in ((4096, 4096, 4096), (1152, 128, 802816 // 8), (8192, 512, 4608)):
    a_k = torch.randn(M, K, device="cuda").bfloat16()
    b_k = torch.randn(N, K, device="cuda").bfloat16()
    a_mn = a_k.t().contiguous()      # [K][M]
    b_mn = b_k.t().contiguous()      # [K][N]
    fl = 2.0 * M * N * K
    for name, fn in (("K/K", lambda: raw.gemm(a_k, b_k)), ("K/MN", lambda: raw.gemm(a_k, b_mn, b_mn=True)),
                     ("MN/MN", lambda: raw.gemm(a_mn, b_mn, a_mn=True, b_mn=True, out_dtype=torch.float32))):
        try:
            ms = timeit(fn)
            print("%dx%dx%d %-6s %.3f ms %.0f TFLOP/s" % (M, N, K, name, ms, fl / ms / 1e9), flush=True)
        except Exception as e:  # noqa: BLE001
            print(M, N, K, name, "ERR", str(e)[:80])
