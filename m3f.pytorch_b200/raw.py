"""Raw tensor-level wrappers over the C ABI (no autograd).  Arguments are CUDA torch tensors already in the
layout the ABI wants (channels-last bf16 activations, packed bf16 filters); outputs are allocated here with
torch.empty so they live in PyTorch's caching allocator and on the current stream."""
import torch

from . import lib as L


def _chk_bf16(*ts):
    for t in ts:
        if t is not None:
            assert t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous(), (t.dtype, t.shape, t.is_contiguous())


def gemm(A, B, a_mn=False, b_mn=False, out_dtype=torch.bfloat16, scale=None, shift=None, residual=None,
         relu=False, stats=None, out=None):
    """D[M,N] = act((A . B^T) * scale + shift + residual).

    A: [M,K] (a_mn False) or [K,M] (a_mn True); B: [N,K] (b_mn False) or [K,N] (b_mn True); bf16, row-major,
    leading dimension = size(1) (must be a multiple of 8)."""
    _chk_bf16(A, B, residual)
    if a_mn:
        K, M = A.shape
    else:
        M, K = A.shape
    if b_mn:
        Kb, N = B.shape
    else:
        N, Kb = B.shape
    assert K == Kb, (A.shape, B.shape, a_mn, b_mn)
    if out is None:
        out = torch.empty((M, N), device=A.device, dtype=out_dtype)
    assert out.dtype in (torch.bfloat16, torch.float32) and out.stride(1) == 1
    rc = L.load().m3t_gemm_bf16(
        L.ptr(A), L.i64(A.stride(0)), L.i32(a_mn), L.ptr(B), L.i64(B.stride(0)), L.i32(b_mn), L.ptr(out),
        L.i64(out.stride(0)), L.i32(out.dtype == torch.float32), L.i32(M), L.i32(N), L.i32(K), L.ptr(scale),
        L.ptr(shift), L.ptr(residual), L.i64(residual.stride(0) if residual is not None else 0), L.i32(relu),
        L.ptr(stats), L.stream_ptr())
    L.check(rc, "m3t_gemm_bf16")
    return out


def conv_geom(nd, N, D, H, W, Cin, Cout, k, stride, pad_lo, pad_hi, dil):
    """Build the 22-int geometry vector of include/m3t_b200.h.  k/stride/pad/dil are (d,h,w) triples."""
    kd, kh, kw = k
    sd, sh, sw = stride
    pdl, phl, pwl = pad_lo
    pdu, phu, pwu = pad_hi
    dd, dh, dw = dil
    return [nd, N, D, H, W, Cin, Cout, kd, kh, kw, sd, sh, sw, pdl, pdu, phl, phu, pwl, pwu, dd, dh, dw]


def conv_out_dims(g):
    nd, N, D, H, W, Cin, Cout, kd, kh, kw, sd, sh, sw, pdl, pdu, phl, phu, pwl, pwu, dd, dh, dw = g
    Z = (D + pdl + pdu - dd * (kd - 1) - 1) // sd + 1
    P = (H + phl + phu - dh * (kh - 1) - 1) // sh + 1
    Q = (W + pwl + pwu - dw * (kw - 1) - 1) // sw + 1
    return Z, P, Q


def conv_fprop(x, w_packed, g, scale=None, shift=None, residual=None, relu=False, stats=None, tile_hint=0):
    """x: bf16 [N,D,H,W,Cin] contiguous (any view with that element order), w_packed: bf16 [Cout, taps*Cin].
    Returns bf16 [N,Z,P,Q,Cout]."""
    _chk_bf16(x, w_packed, residual)
    Z, P, Q = conv_out_dims(g)
    N, Cout = g[1], g[6]
    y = torch.empty((N, Z, P, Q, Cout), device=x.device, dtype=torch.bfloat16)
    rc = L.load().m3t_conv_fprop_bf16(L.ptr(x), L.ptr(w_packed), L.ptr(y), L.int_array(g), L.ptr(scale),
                                      L.ptr(shift), L.ptr(residual), L.i32(relu), L.ptr(stats), L.i32(tile_hint),
                                      L.stream_ptr())
    L.check(rc, "m3t_conv_fprop_bf16")
    return y


def conv_wgrad(x, dy, g, splits=0):
    """Returns fp32 [Cout, taps*Cin] (packed, tap-major / channel-minor)."""
    _chk_bf16(x, dy)
    Cin, Cout = g[5], g[6]
    taps = g[7] * g[8] * g[9]
    dw = torch.zeros((Cout, taps * Cin), device=x.device, dtype=torch.float32)
    rc = L.load().m3t_conv_wgrad_bf16(L.ptr(x), L.ptr(dy), L.ptr(dw), L.int_array(g), L.i32(splits), L.stream_ptr())
    L.check(rc, "m3t_conv_wgrad_bf16")
    return dw
