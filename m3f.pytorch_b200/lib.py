"""ctypes binding of libm3t_b200.so (the C ABI declared in include/m3t_b200.h).

There is no fallback: if the shared library is missing or a call fails, the op raises.  The product path never
routes through oracle/ or through stock PyTorch kernels for the ops this library implements.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libm3t_b200.so")
HEADER_PATH = os.path.normpath(os.path.join(_HERE, "..", "include", "m3t_b200.h"))

_lib = None


class M3TError(RuntimeError):
    pass


def header_symbols():
    """Every function name declared in include/m3t_b200.h (used by the CPU-side ABI test)."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(m3t_[a-z0-9_]+)\s*\(", txt)))


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise M3TError(
                "libm3t_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`; "
                "there is no CPU / PyTorch fallback for the CUDA path." % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.m3t_abi_version.restype = ctypes.c_int
        _lib.m3t_launch_count.restype = ctypes.c_longlong
    return _lib


def launch_count():
    return int(load().m3t_launch_count())


def set_pdl(on):
    """Programmatic dependent launch of the library's kernels (include/m3t_b200.h m3t_set_pdl); returns the previous
    setting.  on=None only queries."""
    return bool(load().m3t_set_pdl(-1 if on is None else int(bool(on))))


def check(rc, what):
    if rc != 0:
        raise M3TError("%s failed with code %d" % (what, rc))


def ptr(t):
    """Device pointer of a torch tensor (or NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def i32(v):
    return ctypes.c_int(int(v))


def i64(v):
    return ctypes.c_longlong(int(v))


def f32(v):
    return ctypes.c_float(float(v))


def int_array(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])
