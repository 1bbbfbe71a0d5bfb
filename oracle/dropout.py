"""ORACLE (test infrastructure, never imported by the product): numpy restatement of `m3t_dropout_bf16`'s mask stream.

The reference's nn.Dropout (models/tcn.py:23,29) draws its mask from PyTorch's Philox generator, which no independent
implementation reproduces bit for bit; the arithmetic that must match is "inverted dropout": each element is kept with
probability 1 - p and kept values are multiplied by 1 / (1 - p).  The product's mask comes from a counter-based
generator (one splitmix64 evaluation per element); this file restates that generator with numpy uint64 arithmetic, so
the CUDA kernel's mask is checked exactly and its keep rate / independence statistically (tests/test_cpu_host.py)."""
import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def uniform_u32(seed, n):
    """u_i for i in [0, n): top 32 bits of splitmix64(seed + (i + 1) * golden), all arithmetic modulo 2^64."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (np.arange(1, n + 1, dtype=np.uint64)) * _GOLD
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(32)).astype(np.uint32)


def threshold(p):
    t = float(np.float32(p)) * 4294967296.0        # the ABI takes p as a C float
    return np.uint32(4294967295 if t >= 4294967295.0 else int(t))


def keep_mask(seed, n, p):
    return uniform_u32(seed, n) >= threshold(p)


def dropout(x, p, seed):
    """x: float32 ndarray holding bf16-representable values -> float32 (before the kernel's bf16 rounding)."""
    flat = np.asarray(x, dtype=np.float32).reshape(-1)
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return np.where(keep_mask(seed, flat.size, p), flat * scale, np.float32(0)).reshape(np.shape(x))
