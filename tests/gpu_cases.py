"""GPU parity cases for the raw tensor-core ops (GEMM, implicit-GEMM conv fprop / wgrad).

Every case draws seeded inputs, rounds them to bf16 (so both sides see identical operands), runs the CUDA path
through the C ABI and compares with a plain fp32 PyTorch CPU evaluation of the same op.  Used two ways:
  * `python -m tests.gpu_cases <name>`  — one case in its own process (tests/gpu_probe.py isolates cases so a
    trapped kernel cannot take the rest of the run down),
  * imported by tests/test_gpu_*.py as pytest `-m gpu` tests.
Metric: max|y - ref| / max|ref| (range-normalised).
"""
import json
import sys

import torch
import torch.nn.functional as F


def _rnd(shape, g, scale=1.0):
    return (torch.randn(shape, generator=g) * scale).bfloat16()


def _err(y, ref):
    y = y.detach().float().cpu()
    ref = ref.detach().float().cpu()
    return float((y - ref).abs().max() / ref.abs().max().clamp_min(1e-12))


# ----------------------------------------------------------------------------------------------------------
# GEMM
# ----------------------------------------------------------------------------------------------------------
def case_gemm(M, N, K, a_mn=False, b_mn=False, out_f32=False, bias=False, relu=False, residual=False, stats=False,
              seed=0):
    from m3t_b200 import raw
    g = torch.Generator().manual_seed(seed)
    pad8 = lambda v: (v + 7) // 8 * 8
    A = _rnd((M, K), g)
    B = _rnd((N, K), g)
    ref = A.float() @ B.float().t()
    dev = "cuda"
    # storage with leading dimensions padded to 8 elements
    if a_mn:
        As = torch.zeros((K, pad8(M)), dtype=torch.bfloat16)
        As[:, :M] = A.t()
        Ad = As.to(dev)
    else:
        As = torch.zeros((M, pad8(K)), dtype=torch.bfloat16)
        As[:, :K] = A
        Ad = As.to(dev)
    if b_mn:
        Bs = torch.zeros((K, pad8(N)), dtype=torch.bfloat16)
        Bs[:, :N] = B.t()
        Bd = Bs.to(dev)
    else:
        Bs = torch.zeros((N, pad8(K)), dtype=torch.bfloat16)
        Bs[:, :K] = B
        Bd = Bs.to(dev)
    lib = __import__("m3t_b200.lib", fromlist=["x"])
    L = lib
    out = torch.empty((M, N), device=dev, dtype=torch.float32 if out_f32 else torch.bfloat16)
    sh = torch.randn(N, generator=g) if bias else None
    res = _rnd((M, N), g) if residual else None
    st = torch.zeros((2, N), device=dev, dtype=torch.float32) if stats else None
    shd = sh.to(dev) if bias else None
    resd = res.to(dev) if residual else None
    lda = Ad.stride(0)
    ldb = Bd.stride(0)
    rc = L.load().m3t_gemm_bf16(L.ptr(Ad), L.i64(lda), L.i32(a_mn), L.ptr(Bd), L.i64(ldb), L.i32(b_mn), L.ptr(out),
                                L.i64(out.stride(0)), L.i32(out_f32), L.i32(M), L.i32(N), L.i32(K), L.ptr(None),
                                L.ptr(shd), L.ptr(resd), L.i64(N if residual else 0), L.i32(relu), L.ptr(st),
                                L.stream_ptr())
    L.check(rc, "gemm")
    torch.cuda.synchronize()
    r = ref.clone()
    if bias:
        r = r + sh
    if residual:
        r = r + res.float()
    if relu:
        r = r.relu()
    errs = {"out": _err(out, r)}
    if stats:
        errs["sum"] = _err(st[0], ref.sum(0))
        errs["sumsq"] = _err(st[1], (ref * ref).sum(0))
    return errs


# ----------------------------------------------------------------------------------------------------------
# Convolution (channels-last bf16 through the ABI vs F.convNd in fp32 on the CPU)
# ----------------------------------------------------------------------------------------------------------
def _pack_w(w):
    """[Cout, Cin, *k] -> [Cout, taps*Cin] tap-major, channel-minor."""
    Cout, Cin = w.shape[:2]
    nd = w.dim() - 2
    perm = [0] + list(range(2, 2 + nd)) + [1]
    return w.permute(*perm).reshape(Cout, -1).contiguous()


def _conv_ref(x, w, nd, stride, pad_lo, pad_hi, dil):
    # x: [N, C, (D,) (H,) W] fp32 ; asymmetric zero padding applied explicitly
    pads = []
    for lo, hi in reversed(list(zip(pad_lo[3 - nd:], pad_hi[3 - nd:]))):
        pads += [lo, hi]
    xp = F.pad(x, pads)
    fn = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[nd]
    return fn(xp, w, stride=tuple(stride[3 - nd:]), dilation=tuple(dil[3 - nd:]))


def case_conv(nd, N, D, H, W, Cin, Cout, k, stride, pad_lo, pad_hi=None, dil=(1, 1, 1), fused=False, stats=False,
              tile_hint=0, wgrad=False, seed=0):
    from m3t_b200 import raw
    pad_hi = pad_hi if pad_hi is not None else pad_lo
    g = torch.Generator().manual_seed(seed)
    spatial = [D, H, W][3 - nd:]
    x = _rnd([N, Cin] + spatial, g)
    w = _rnd([Cout, Cin] + list(k[3 - nd:]), g, scale=(2.0 / (Cin * k[0] * k[1] * k[2])) ** 0.5)
    geom = raw.conv_geom(nd, N, D, H, W, Cin, Cout, k, stride, pad_lo, pad_hi, dil)
    y_ref = _conv_ref(x.float(), w.float(), nd, stride, pad_lo, pad_hi, dil)   # [N, Cout, ...]
    to_cl = lambda t: t.permute(0, *range(2, t.dim()), 1).contiguous()            # NC... -> N...C
    x_cl = to_cl(x).cuda()
    wp = _pack_w(w).cuda()
    errs = {}
    if not wgrad:
        scale = shift = res = st = None
        r = y_ref
        if fused:
            scale = torch.rand(Cout, generator=g) + 0.5
            shift = torch.randn(Cout, generator=g) * 0.1
            res = _rnd(list(y_ref.shape), g)
            bshape = [1, Cout] + [1] * nd
            r = (y_ref * scale.view(bshape) + shift.view(bshape) + res.float()).relu()
        if stats:
            st = torch.zeros((2, Cout), device="cuda", dtype=torch.float32)
        y = raw.conv_fprop(x_cl, wp, geom, scale=scale.cuda() if fused else None,
                           shift=shift.cuda() if fused else None,
                           residual=to_cl(res).cuda() if fused else None, relu=fused, stats=st, tile_hint=tile_hint)
        torch.cuda.synchronize()
        y_nc = y.reshape([N] + list(y_ref.shape[2:]) + [Cout]).permute(0, nd + 1, *range(1, nd + 1))
        errs["out"] = _err(y_nc, r)
        if stats:
            red = [0] + list(range(2, 2 + nd))
            errs["sum"] = _err(st[0], y_ref.sum(red))
            errs["sumsq"] = _err(st[1], (y_ref * y_ref).sum(red))
    else:
        dy = _rnd(list(y_ref.shape), g)
        xr = x.float().requires_grad_(False)
        wr = w.float().clone().requires_grad_(True)
        yr = _conv_ref(xr, wr, nd, stride, pad_lo, pad_hi, dil)
        yr.backward(dy.float())
        dw_ref = _pack_w(wr.grad)
        dw = raw.conv_wgrad(x_cl, to_cl(dy).cuda(), geom)
        torch.cuda.synchronize()
        errs["dw"] = _err(dw, dw_ref)
    return errs


def _c(**kw):
    return kw


CASES = {
    # ---- GEMM ----
    "gemm_nt_exact": (case_gemm, _c(M=256, N=128, K=128)),
    "gemm_nt_tails_f32": (case_gemm, _c(M=300, N=200, K=200, out_f32=True, bias=True, relu=True)),
    "gemm_nt_n9": (case_gemm, _c(M=130, N=9, K=512, out_f32=True, bias=True)),
    "gemm_nt_res_stats": (case_gemm, _c(M=384, N=64, K=576, residual=True, stats=True)),
    "gemm_kmn": (case_gemm, _c(M=256, N=192, K=128, b_mn=True)),
    "gemm_kmn_tails": (case_gemm, _c(M=200, N=200, K=72, b_mn=True, out_f32=True)),
    "gemm_mnmn": (case_gemm, _c(M=128, N=192, K=320, a_mn=True, b_mn=True, out_f32=True)),
    "gemm_mnmn_tails": (case_gemm, _c(M=9, N=512, K=130, a_mn=True, b_mn=True, out_f32=True)),
    "gemm_big": (case_gemm, _c(M=4096, N=1536, K=512)),
    # ---- conv forward ----
    "conv2d_3x3_s1_64": (case_conv, _c(nd=2, N=3, D=1, H=28, W=28, Cin=64, Cout=64, k=(1, 3, 3), stride=(1, 1, 1),
                                       pad_lo=(0, 1, 1))),
    "conv2d_3x3_s2_64_128": (case_conv, _c(nd=2, N=3, D=1, H=28, W=28, Cin=64, Cout=128, k=(1, 3, 3),
                                           stride=(1, 2, 2), pad_lo=(0, 1, 1))),
    "conv2d_1x1_s2": (case_conv, _c(nd=2, N=3, D=1, H=28, W=28, Cin=64, Cout=128, k=(1, 1, 1), stride=(1, 2, 2),
                                    pad_lo=(0, 0, 0))),
    "conv2d_3x3_7x7_256": (case_conv, _c(nd=2, N=5, D=1, H=7, W=7, Cin=256, Cout=256, k=(1, 3, 3),
                                         stride=(1, 1, 1), pad_lo=(0, 1, 1))),
    "conv2d_3x3_s2_7to4": (case_conv, _c(nd=2, N=5, D=1, H=7, W=7, Cin=256, Cout=512, k=(1, 3, 3),
                                         stride=(1, 2, 2), pad_lo=(0, 1, 1))),
    "conv2d_fused_stats": (case_conv, _c(nd=2, N=4, D=1, H=14, W=14, Cin=128, Cout=128, k=(1, 3, 3),
                                         stride=(1, 1, 1), pad_lo=(0, 1, 1), fused=True)),
    "conv2d_stats": (case_conv, _c(nd=2, N=4, D=1, H=14, W=14, Cin=128, Cout=128, k=(1, 3, 3),
                                   stride=(1, 1, 1), pad_lo=(0, 1, 1), stats=True)),
    "conv2d_mt2": (case_conv, _c(nd=2, N=3, D=1, H=28, W=28, Cin=64, Cout=64, k=(1, 3, 3), stride=(1, 1, 1),
                                 pad_lo=(0, 1, 1), tile_hint=2, stats=True)),
    "conv2d_bn256": (case_conv, _c(nd=2, N=9, D=1, H=4, W=4, Cin=512, Cout=512, k=(1, 3, 3), stride=(1, 1, 1),
                                   pad_lo=(0, 1, 1), tile_hint=4)),
    "conv1d_causal_dil2": (case_conv, _c(nd=1, N=4, D=1, H=1, W=32, Cin=512, Cout=512, k=(1, 1, 3),
                                         stride=(1, 1, 1), pad_lo=(0, 0, 4), pad_hi=(0, 0, 0), dil=(1, 1, 2))),
    "conv1d_k5": (case_conv, _c(nd=1, N=4, D=1, H=1, W=32, Cin=1024, Cout=512, k=(1, 1, 5), stride=(1, 1, 1),
                                pad_lo=(0, 0, 2))),
    "conv3d_3x3x3": (case_conv, _c(nd=3, N=2, D=8, H=12, W=12, Cin=64, Cout=128, k=(3, 3, 3), stride=(1, 1, 1),
                                   pad_lo=(1, 0, 0))),
    # ---- conv wgrad ----
    "wgrad2d_3x3_s1_64": (case_conv, _c(nd=2, N=3, D=1, H=28, W=28, Cin=64, Cout=64, k=(1, 3, 3), stride=(1, 1, 1),
                                        pad_lo=(0, 1, 1), wgrad=True)),
    "wgrad2d_3x3_s2": (case_conv, _c(nd=2, N=3, D=1, H=28, W=28, Cin=64, Cout=128, k=(1, 3, 3), stride=(1, 2, 2),
                                     pad_lo=(0, 1, 1), wgrad=True)),
    "wgrad2d_1x1_s2": (case_conv, _c(nd=2, N=3, D=1, H=14, W=14, Cin=128, Cout=256, k=(1, 1, 1), stride=(1, 2, 2),
                                     pad_lo=(0, 0, 0), wgrad=True)),
    "wgrad2d_7x7_256": (case_conv, _c(nd=2, N=5, D=1, H=7, W=7, Cin=256, Cout=256, k=(1, 3, 3), stride=(1, 1, 1),
                                      pad_lo=(0, 1, 1), wgrad=True)),
    "wgrad1d_causal": (case_conv, _c(nd=1, N=4, D=1, H=1, W=32, Cin=512, Cout=512, k=(1, 1, 3), stride=(1, 1, 1),
                                     pad_lo=(0, 0, 4), pad_hi=(0, 0, 0), dil=(1, 1, 2), wgrad=True)),
    "wgrad3d": (case_conv, _c(nd=3, N=2, D=8, H=12, W=12, Cin=64, Cout=128, k=(3, 3, 3), stride=(1, 1, 1),
                              pad_lo=(1, 0, 0), wgrad=True)),
}

TOL = 1.5e-2  # bf16 output rounding is 2^-8 relative; fp32-accumulated sums are far tighter


def run_case(name):
    fn, kw = CASES[name]
    return fn(**kw)


if __name__ == "__main__":
    name = sys.argv[1]
    errs = run_case(name)
    ok = all(v == v and v < TOL for v in errs.values())
    print("CASE_RESULT " + json.dumps({"case": name, "ok": ok, "errs": errs}))
    sys.exit(0 if ok else 1)
