"""GPU parity cases for the raw tensor-core ops (GEMM, implicit-GEMM conv fprop / wgrad).

Every case draws seeded inputs, rounds them to bf16 (so both sides see identical operands), runs the CUDA path
through the C ABI and compares with a plain fp32 PyTorch CPU evaluation of the same op.  Used two ways:
  * `python -m tests.gpu_cases <name>`  — one case in its own process (tests/gpu_probe.py isolates cases so a
    trapped kernel cannot take the rest of the run down),
  * imported by tests/test_gpu_*.py as pytest `-m gpu` tests.
Metric: max|y - ref| / max|ref| (range-normalised).
"""
import json
import os
import math
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
    "conv2d_bn128_forced": (case_conv, _c(nd=2, N=9, D=1, H=4, W=4, Cin=512, Cout=512, k=(1, 3, 3), stride=(1, 1, 1),
                                   pad_lo=(0, 1, 1), tile_hint=8)),
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




# ----------------------------------------------------------------------------------------------------------
# Elementwise kernels vs plain torch (CPU fp32)
# ----------------------------------------------------------------------------------------------------------
def case_elementwise(seed=0):
    from m3t_b200 import raw
    g = torch.Generator().manual_seed(seed)
    errs = {}
    # video prep (space-to-depth + normalise)
    v = torch.randint(0, 256, (2, 3, 3, 16, 20), generator=g).float()
    xs = raw.video_prep_s2d(v.cuda(), True).float().cpu()
    vn = (v - 127.5) / 127.5
    ref = torch.zeros(2, 3, 8, 10, 16)
    for ph in range(2):
        for pw in range(2):
            for c in range(3):
                ref[..., (ph * 2 + pw) * 3 + c] = vn[:, c, :, ph::2, pw::2]
    errs["prep"] = _err(xs, ref)
    xs8 = raw.video_prep_s2d(v.to(torch.uint8).cuda(), True).float().cpu()
    errs["prep_u8"] = _err(xs8, ref)
    x4 = raw.video_prep_s2d_w4(v.cuda(), True).float().cpu()          # (2,3,8,10,64)
    ref4 = torch.zeros(2, 3, 8, 10, 64)
    for jw in range(4):
        for w2 in range(10):
            ws = w2 + jw - 2
            if 0 <= ws < 10:
                ref4[:, :, :, w2, jw * 12:(jw + 1) * 12] = ref[:, :, :, ws, :12]
    errs["prep_w4"] = _err(x4, ref4)
    # layout conversion round trip
    x = torch.randn(3, 70, 5, 9, generator=g)
    cl = raw.ncs_to_nsc_bf16(x.cuda())
    errs["to_cl"] = _err(cl.float().cpu(), x.permute(0, 2, 3, 1))
    back = raw.nsc_to_ncs_f32(cl)
    errs["from_cl"] = _err(back.cpu(), x.bfloat16().float())
    # bn_act with residual + its own affine
    C = 128
    y = _rnd((6, 7, 7, C), g)
    r = _rnd((6, 7, 7, C), g)
    sc, sh = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    rs, rb = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    o = raw.bn_act(y.cuda(), sc.cuda(), sh.cuda(), r.cuda(), rs.cuda(), rb.cuda(), True)
    errs["bn_act"] = _err(o, (y.float() * sc + sh + r.float() * rs + rb).relu())
    # bn finalize
    st = torch.stack((y.float().sum((0, 1, 2)), (y.float() ** 2).sum((0, 1, 2))))
    gam, bet = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    rm, rv = torch.zeros(C), torch.ones(C)
    rmd, rvd = rm.cuda(), rv.cuda()
    n = y.numel() // C
    fin = raw.bn_finalize(st.cuda(), n, gam.cuda(), bet.cuda(), 1e-5, 0.1, rmd, rvd).cpu()
    mean, var = y.float().mean((0, 1, 2)), y.float().var((0, 1, 2), unbiased=False)
    errs["bn_mean"] = _err(fin[0], mean)
    errs["bn_invstd"] = _err(fin[1], 1 / torch.sqrt(var + 1e-5))
    errs["bn_scale"] = _err(fin[2], gam / torch.sqrt(var + 1e-5))
    errs["bn_shift"] = _err(fin[3], bet - mean * gam / torch.sqrt(var + 1e-5))
    errs["bn_rm"] = _err(rmd, 0.1 * mean)
    errs["bn_rv"] = _err(rvd, 0.9 + 0.1 * var * n / (n - 1))
    # stem tail: bn + relu + maxpool, forward and backward against autograd (even sizes take the 2x2-block
    # specialisation, odd sizes the generic kernels)
    for H, W, tag in ((12, 10, ""), (11, 9, "_odd")):
        F_, C = 3, 64
        yy = _rnd((F_, H, W, C), g)
        sc, sh = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
        out, idx = raw.bn_relu_maxpool(yy.cuda(), sc.cuda(), sh.cuda(), True)
        yr = yy.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
        act = (yr * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)).relu()
        pr = F.max_pool2d(act, 3, 2, 1)
        errs["maxpool" + tag] = _err(out.float().cpu().permute(0, 3, 1, 2), pr)
        # backward through pool+relu+BN(train) : use batch stats of yy so BN backward terms are exercised
        mean, var = yy.float().mean((0, 1, 2)), yy.float().var((0, 1, 2), unbiased=False)
        invstd = 1 / torch.sqrt(var + 1e-5)
        gam, bet = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
        scale, shift = gam * invstd, bet - mean * gam * invstd
        out2, idx2 = raw.bn_relu_maxpool(yy.cuda(), scale.cuda(), shift.cuda(), True)
        dout = _rnd(tuple(out2.shape), g)
        dy, sums = raw.maxpool_bn_bwd(dout.cuda(), idx2, yy.cuda(), mean.cuda(), invstd.cuda(), scale.cuda(), shift.cuda(),
                                      F_ * H * W)
        y3 = yy.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
        gp, bp = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
        z = F.batch_norm(y3, None, None, gp, bp, True, 0.1, 1e-5).relu()
        F.max_pool2d(z, 3, 2, 1).backward(dout.float().permute(0, 3, 1, 2))
        errs["pool_bwd_dy" + tag] = _err(dy.float().cpu().permute(0, 3, 1, 2), y3.grad)
        errs["pool_bwd_dgamma" + tag] = _err(sums[1], gp.grad)
        errs["pool_bwd_dbeta" + tag] = _err(sums[0], bp.grad)
        # the same backward with the BatchNorm sums taken over the POOLED tensors (raw y at the arg-max from the forward)
        out3, idx3, ymax = raw.bn_relu_maxpool(yy.cuda(), scale.cuda(), shift.cuda(), True, want_ymax=True)
        dy3, sums3 = raw.maxpool_bn_bwd(dout.cuda(), idx3, yy.cuda(), mean.cuda(), invstd.cuda(), scale.cuda(),
                                        shift.cuda(), F_ * H * W, ymax=ymax)
        errs["pool_ymax_fwd_exact" + tag] = float((out3 != out2).sum() + (idx3 != idx2).sum())
        errs["pool_ymax_dy" + tag] = _err(dy3.float().cpu().permute(0, 3, 1, 2), y3.grad)
        errs["pool_ymax_dgamma" + tag] = _err(sums3[1], gp.grad)
        errs["pool_ymax_dbeta" + tag] = _err(sums3[0], bp.grad)
    # bn_act backward (relu + residual)
    C = 64
    y = _rnd((4, 6, 6, C), g)
    res = _rnd((4, 6, 6, C), g)
    gam, bet = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
    mean, var = y.float().mean((0, 1, 2)), y.float().var((0, 1, 2), unbiased=False)
    invstd = 1 / torch.sqrt(var + 1e-5)
    scale, shift = gam * invstd, bet - mean * gam * invstd
    o = raw.bn_act(y.cuda(), scale.cuda(), shift.cuda(), res.cuda(), None, None, True)
    dout = _rnd(tuple(o.shape), g)
    sums, dz = raw.bn_bwd_reduce(dout.cuda(), o, y.cuda(), mean.cuda(), invstd.cuda(), True, True)
    dy = raw.bn_bwd_apply(dout.cuda(), o, y.cuda(), mean.cuda(), invstd.cuda(), scale.cuda(), sums, y.numel() // C,
                          True)
    # mask recomputed from y (unit without residual)
    o2 = raw.bn_act(y.cuda(), scale.cuda(), shift.cuda(), None, None, None, True)
    sums2, _ = raw.bn_bwd_reduce(dout.cuda(), None, y.cuda(), mean.cuda(), invstd.cuda(), True, False,
                                 scale=scale.cuda(), shift=shift.cuda())
    dy2 = raw.bn_bwd_apply(dout.cuda(), None, y.cuda(), mean.cuda(), invstd.cuda(), scale.cuda(), sums2,
                           y.numel() // C, True, shift=shift.cuda())
    yq = y.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    gq, bq = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
    F.batch_norm(yq, None, None, gq, bq, True, 0.1, 1e-5).relu().backward(dout.float().permute(0, 3, 1, 2))
    errs["bn_bwd2_dy"] = _err(dy2.float().cpu().permute(0, 3, 1, 2), yq.grad)
    errs["bn_bwd2_dgamma"] = _err(sums2[1], gq.grad)
    errs["bn_bwd2_dbeta"] = _err(sums2[0], bq.grad)
    yr = y.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    rr = res.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    gp, bp = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
    (F.batch_norm(yr, None, None, gp, bp, True, 0.1, 1e-5) + rr).relu().backward(dout.float().permute(0, 3, 1, 2))
    errs["bn_bwd_dy"] = _err(dy.float().cpu().permute(0, 3, 1, 2), yr.grad)
    errs["bn_bwd_dres"] = _err(dz.float().cpu().permute(0, 3, 1, 2), rr.grad)
    errs["bn_bwd_dgamma"] = _err(sums[1], gp.grad)
    errs["bn_bwd_dbeta"] = _err(sums[0], bp.grad)
    # avgpool fwd/bwd, filter pack, zero insert, colsum, add
    x = _rnd((5, 4, 4, 512), g)
    ob, of = raw.avgpool(x.cuda(), True)
    errs["avgpool"] = _err(of, x.float().mean((1, 2)))
    dx = raw.avgpool_bwd(of, (5, 4, 4, 512))
    errs["avgpool_bwd"] = _err(dx, (x.float().mean((1, 2)) / 16).view(5, 1, 1, 512).expand(5, 4, 4, 512))
    w = torch.randn(128, 64, 3, 3, generator=g)
    wf, wd = raw.pack_filter(w.cuda(), True)
    errs["pack_f"] = _err(wf, w.permute(0, 2, 3, 1).reshape(128, -1))
    errs["pack_d"] = _err(wd, w.flip(2, 3).permute(1, 2, 3, 0).reshape(64, -1))
    dwp = torch.randn(128, 9 * 64, generator=g)
    errs["unpack"] = _err(raw.unpack_filter_grad(dwp.cuda(), (128, 64, 3, 3)),
                          dwp.view(128, 3, 3, 64).permute(0, 3, 1, 2))
    dyy = _rnd((2, 4, 4, 64), g)
    up = raw.zero_insert2(dyy.cuda(), 7, 7).float().cpu()
    refu = torch.zeros(2, 7, 7, 64)
    refu[:, ::2, ::2] = dyy.float()
    errs["zero_insert"] = _err(up, refu)
    m = _rnd((1000, 24), g)
    errs["colsum"] = _err(raw.colsum(m.cuda(), 20), m.float()[:, :20].sum(0))
    return errs


# ----------------------------------------------------------------------------------------------------------
# Golden-fixture module cases: the CUDA path vs outputs of the unmodified reference modules
# ----------------------------------------------------------------------------------------------------------
def _build(fx):
    import argparse
    from oracle.ref_torch import synth_state_dict
    kind = fx["kind"]
    if kind == "GRU":
        from m3t_b200.models.rnn import GRU
        m = GRU(**fx["ctor"])
    elif kind == "AttFusion":
        from m3t_b200.models.att_fusion import AttFusion
        m = AttFusion(**fx["ctor"])
    elif kind == "TemporalConvNet":
        from m3t_b200.models.tcn import TemporalConvNet
        m = TemporalConvNet(**fx["ctor"])
    elif kind == "ResNet":
        from m3t_b200.models.resnet import BasicBlock, ResNet
        m = ResNet(BasicBlock, [2, 2, 2, 2], 512, zero_init_residual=True, agg_mode="ap", fmap_out_size=3)
    elif kind == "ResNetV2":
        from m3t_b200.models.resnet import BasicBlockV2, ResNetV2
        m = ResNetV2(BasicBlockV2, [2, 2, 2, 2], 512, zero_init_residual=False, agg_mode="ap", fmap_out_size=3)
    elif kind == "AttEncDec":
        from m3t_b200.models.rnn import AttEncDec
        m = AttEncDec()
    elif kind == "CBAM":
        from m3t_b200.models.cbam import CBAM
        m = CBAM(**fx["ctor"])
    elif kind == "VGGFace":
        from m3t_b200.models.vggface import VGGFace
        m = VGGFace()
    elif kind == "DenseNet52_3D":
        from m3t_b200.models.densenet import DenseNet52_3D
        m = DenseNet52_3D(392, agg_mode="ap", fmap_out_size=3)
    elif kind == "ResNetCBAM":
        from m3t_b200.models.resnet import BasicBlock, ResNet
        m = ResNet(BasicBlock, [1, 1, 1, 1], 512, zero_init_residual=True, agg_mode="ap", fmap_out_size=3, use_cbam=True)
    elif kind == "VA_3DResNet":
        from m3t_b200.models.backbone import VA_3DResNet
        m = VA_3DResNet(**fx["ctor"])
    elif kind == "VA_3DVGGM_Split":
        from m3t_b200.models.vggm import VA_3DVGGM_Split
        m = VA_3DVGGM_Split(**fx["ctor"])
    elif kind == "VA_3DVGGM":
        from m3t_b200.models.vggm import VA_3DVGGM
        m = VA_3DVGGM(**fx["ctor"])
    elif kind == "AffWild2VA":
        from m3t_b200.models.model import AffWild2VA
        m = AffWild2VA(argparse.Namespace(**fx["hparams"]))
    else:
        raise KeyError(kind)
    m.load_state_dict(synth_state_dict(fx["spec"], fx["seed"], **fx.get("synth_kw", {})), strict=True)
    if "dropout_p" in fx:
        for d in m.modules():
            if hasattr(d, "dropout_p"):
                d.dropout_p = fx["dropout_p"]
            if isinstance(d, torch.nn.Dropout):
                d.p = fx["dropout_p"]
    return m.cuda()


def _oracle_run(fx, emulate, want_grads):
    """Evaluate the oracle (optionally with bf16 storage emulation) on a fixture; returns (out, loss, grads dict)."""
    from oracle import ref_torch as R
    from tests.golden_util import hparams_ns, ref_batch
    import contextlib
    sd = R.synth_state_dict(fx["spec"], fx["seed"], **fx.get("synth_kw", {}))
    if want_grads:
        for k, v in sd.items():
            if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
                v.requires_grad_(True)
    kind, inp = fx["kind"], fx["inputs"]
    train = fx.get("mode", "train") == "train"
    leaves = {}
    ctx = R.bf16_emulation() if emulate else contextlib.nullcontext()
    loss = None
    with ctx, torch.set_grad_enabled(want_grads):
        if kind == "GRU":
            x = inp["x"].clone().requires_grad_(want_grads)
            leaves["x"] = x
            out = R.gru_module(x, {"m." + k: v for k, v in sd.items()}, "m")
        elif kind == "AttFusion":
            xa = inp["x_a"].clone().requires_grad_(want_grads)
            xv = inp["x_v"].clone().requires_grad_(want_grads)
            leaves.update(x_a=xa, x_v=xv)
            out = R.att_fusion(xa, xv, {"att_fuse." + k: v for k, v in sd.items()})
        elif kind == "TemporalConvNet":
            x = inp["x"].clone().requires_grad_(want_grads)
            leaves["x"] = x
            out = R.temporal_conv_net(x, sd, "", 2)
        elif kind == "ResNet":
            x = inp["x"].clone().requires_grad_(want_grads)
            leaves["x"] = x
            out = R.resnet_trunk(R.q(x), {"resnet." + k: v for k, v in sd.items()}, train=train)
        elif kind == "ResNetV2":
            x = inp["x"].clone().requires_grad_(want_grads)
            leaves["x"] = x
            out = R.resnet_v2_trunk(R.q(x), {"resnet." + k: v for k, v in sd.items()}, train=train)
        elif kind == "AttEncDec":
            x = inp["x"].clone().requires_grad_(want_grads)
            leaves["x"] = x
            out = R.att_enc_dec(x, {"fusion." + k: v for k, v in sd.items()}, "fusion")
        elif kind == "CBAM":
            x = inp["x"].clone().requires_grad_(want_grads)
            leaves["x"] = x
            out = R.cbam(R.q(x), sd, "", train=train)
        elif kind == "VGGFace":
            out = R.vggface((inp["image_u8"].float() - 127.5) / 127.5, sd)
        elif kind == "DenseNet52_3D":
            x = inp["x"].clone().requires_grad_(want_grads)
            leaves["x"] = x
            out = R.densenet52_3d(R.q(x), {"densenet." + k: v for k, v in sd.items()}, train=train)
        elif kind == "ResNetCBAM":
            x = inp["x"].clone().requires_grad_(want_grads)
            leaves["x"] = x
            out = R.resnet_trunk(R.q(x), {"resnet." + k: v for k, v in sd.items()}, layers=(1, 1, 1, 1), train=train)
        elif kind == "VA_3DResNet":
            out = R.va_3dresnet((inp["video_u8"].float() - 127.5) / 127.5, sd, fx["ctor"]["frameLen"], train=train)
        elif kind == "VA_3DVGGM_Split":
            out = R.va_3dvggm_split((inp["video_u8"].float() - 127.5) / 127.5, inp["se_features"], inp["se_features"],
                                    sd, "", fx["ctor"]["split_layer"], fx["ctor"]["backend"], train=train)
        elif kind == "VA_3DVGGM":
            out = R.va_3dvggm((inp["video_u8"].float() - 127.5) / 127.5, sd, fx["ctor"]["backend"], train=train)
        elif kind == "AffWild2VA":
            b = ref_batch(inp)
            hp = hparams_ns(fx["hparams"])
            out = R.affwild2va_forward(b, sd, hp, train=train)
            if train:
                loss = R.training_loss(out, b, hp.loss, hp.loss_lambda)
        if want_grads:
            (loss if loss is not None else (out * fx["cot"]).sum()).backward()
    grads = {}
    if want_grads:
        for k, v in sd.items():
            if v.is_floating_point() and v.grad is not None:
                grads["param." + k] = v.grad
        for k, v in leaves.items():
            grads["input." + k] = v.grad
    return out.detach(), (float(loss) if loss is not None else None), grads


def _l2(a, b):
    a, b = a.detach().float().cpu().reshape(-1), b.detach().float().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def case_golden(name, grads=True):
    """CUDA path on a golden fixture.  Reported errors:
      out_ref / va_ref : vs the fp32 output of the unmodified reference module (the fixture)     [bf16 target 2e-2]
      out_emu          : vs the oracle with bf16 storage emulation (same rounding points)         [tight]
      floor            : oracle(bf16-emulated) vs reference = intrinsic cost of bf16 storage on these weights
      grad_emu         : worst per-tensor relative L2 gradient error vs the bf16-emulated oracle
    """
    from tests.golden_util import load, ref_batch
    fx = load(name)
    m = _build(fx)
    mode = fx.get("mode", "train")
    m.train(mode == "train")
    kind = fx["kind"]
    inp = fx["inputs"]
    want_grads = grads and "grads" in fx
    leaves = {}
    loss = None
    with torch.set_grad_enabled(want_grads):
        if kind in ("GRU", "TemporalConvNet", "ResNet", "ResNetV2", "AttEncDec", "CBAM", "ResNetCBAM", "DenseNet52_3D"):
            x = inp["x"].cuda().requires_grad_(want_grads)
            leaves["x"] = x
            out = m(x)
        elif kind == "AttFusion":
            xa = inp["x_a"].cuda().requires_grad_(want_grads)
            xv = inp["x_v"].cuda().requires_grad_(want_grads)
            leaves.update(x_a=xa, x_v=xv)
            out = m(xa, xv)
        elif kind in ("VA_3DResNet", "VA_3DVGGM"):
            out = m((inp["video_u8"].float().cuda() - 127.5) / 127.5)
        elif kind == "VGGFace":
            out = m((inp["image_u8"].float().cuda() - 127.5) / 127.5)
        elif kind == "VA_3DVGGM_Split":
            se = inp["se_features"].cuda()
            out = m((inp["video_u8"].float().cuda() - 127.5) / 127.5, se, se)
        elif kind == "AffWild2VA":
            b = ref_batch(inp, "cuda")
            out = m(b)
            if mode == "train":
                loss, _ = m.compute_loss(out, b)
        if want_grads:
            (loss if loss is not None else (out * fx["cot"].cuda()).sum()).backward()
    o_emu, loss_emu, g_emu = _oracle_run(fx, True, want_grads)
    errs = {}
    is_av = kind == "AffWild2VA"
    sl = slice(None)
    if is_av:
        sl = slice(7, None) if "mtl" in fx["hparams"]["loss"] else slice(-2, None)
    if is_av and mode == "train":
        errs["loss_ref"] = abs(float(loss) - fx["loss"]) / max(abs(fx["loss"]), 1e-6)
        errs["loss_emu"] = abs(float(loss) - loss_emu) / max(abs(loss_emu), 1e-6)
    else:
        errs["out_ref"] = _err(out, fx["out"])
        errs["floor"] = _err(o_emu, fx["out"])
        if is_av:
            errs["va_ref"] = _err(out[..., sl], fx["out"][..., sl])
    errs["out_emu"] = _err(out, o_emu)
    if want_grads:
        params = dict(m.named_parameters())
        per = {}
        for k, ge in g_emu.items():
            kind_, nm = k.split(".", 1)
            t = params[nm] if kind_ == "param" else leaves[nm]
            per[k] = _l2(t.grad, ge) if t.grad is not None else float("nan")
        srt = sorted(per.items(), key=lambda kv: -(kv[1] if kv[1] == kv[1] else 1e9))
        errs["grad_emu"] = srt[0][1]
        errs["grad_top"] = {k: round(v, 4) for k, v in srt[:4]}
        allc = torch.cat([(params[k.split(".", 1)[1]].grad).detach().float().cpu().reshape(-1)
                          for k in g_emu if k.startswith("param.")])
        alle = torch.cat([g_emu[k].float().reshape(-1) for k in g_emu if k.startswith("param.")])
        errs["grad_all_l2"] = _l2(allc, alle)
    return errs


CASES.update({
    "stem_s2d_fprop": (case_conv, _c(nd=3, N=2, D=4, H=56, W=56, Cin=16, Cout=64, k=(5, 4, 4), stride=(1, 1, 1),
                                     pad_lo=(2, 2, 2), pad_hi=(2, 1, 1), stats=True)),
    "stem_s2d_wgrad": (case_conv, _c(nd=3, N=2, D=4, H=56, W=56, Cin=16, Cout=64, k=(5, 4, 4), stride=(1, 1, 1),
                                     pad_lo=(2, 2, 2), pad_hi=(2, 1, 1), wgrad=True)),
    "vggm1_s2d_fprop": (case_conv, _c(nd=3, N=2, D=4, H=56, W=56, Cin=16, Cout=64, k=(3, 2, 2), stride=(1, 1, 1),
                                      pad_lo=(1, 0, 0))),
    "elementwise": (case_elementwise, _c()),
    "golden_gru_audio": (case_golden, _c(name="gru_audio")),
    "golden_gru_scorer": (case_golden, _c(name="gru_scorer")),
    "golden_gru_nohead": (case_golden, _c(name="gru_nohead")),
    "golden_attfusion": (case_golden, _c(name="attfusion")),
    "golden_tcn": (case_golden, _c(name="tcn")),
    "golden_resnet_trunk_eval": (case_golden, _c(name="resnet_trunk_eval", grads=False)),
    "golden_resnet_trunk_train": (case_golden, _c(name="resnet_trunk_train")),
    "golden_va3dresnet_eval": (case_golden, _c(name="va3dresnet_eval")),
    "golden_va3dresnet_train": (case_golden, _c(name="va3dresnet_train")),
    "golden_av_resnet_attention_eval": (case_golden, _c(name="av_resnet_attention_eval")),
    "golden_av_resnet_attention_train": (case_golden, _c(name="av_resnet_attention_train")),
    "golden_vggm_split3_eval": (case_golden, _c(name="vggm_split3_eval")),
    "golden_av_v2psplit_attention_eval": (case_golden, _c(name="av_v2psplit_attention_eval")),
})

def case_golden_size(name, chunk=None):
    """Eval-mode parity AT BASELINE's configuration sizes (fixtures of oracle/make_golden_sizes.py: reference output and
    bf16-emulating oracle output stored, inputs / weights re-drawn from seeds).  Reported:
      out_ref / va_ref   CUDA vs the unmodified reference (range-normalised max; V/A channels: the 2e-2 north-star bar)
      out_emu            CUDA vs the bf16-storage-emulating oracle (implementation error only)
      floor / va_floor   emulating oracle vs reference = what bf16 storage itself costs on these weights (reported)
      ccc_ref_diff       |CCC(cuda, labels) - CCC(reference, labels)|, worst of valence / arousal, over ALL frames of
                         the fixture (>= 1024 for cfg3 / cfg4): the north star's "CCC equal to 3 decimals" (5e-4)
      va_rms_ref         RMS (not max) V/A error relative to the reference's V/A RMS (reported)
    """
    from tests.golden_util import load, ref_batch
    from m3t_b200.models.utils import concordance_cc2
    fx = load(name)
    m = _build(fx).eval()
    inp = fx["inputs"]
    with torch.no_grad():
        if fx["kind"] == "AffWild2VA":
            B = inp["video_u8"].shape[0]
            step = chunk or B
            outs = []
            for lo in range(0, B, step):
                outs.append(m(ref_batch({k: v[lo:lo + step] for k, v in inp.items()}, "cuda")).float().cpu())
            out = torch.cat(outs)
        else:
            out = m((inp["video_u8"].float().cuda() - 127.5) / 127.5).float().cpu()
    ref, emu = fx["out"], fx.get("out_emu")
    errs = {"out_ref": _err(out, ref)}
    nva = 2
    if out.shape[-1] >= 2:
        errs["va_ref"] = _err(out[..., -nva:], ref[..., -nva:])
        errs["va_rms_ref"] = {"value": round(float((out[..., -nva:] - ref[..., -nva:]).pow(2).mean().sqrt() /
                                                   ref[..., -nva:].pow(2).mean().sqrt()), 5)}
    if emu is not None:
        errs["out_emu"] = _err(out, emu)
        errs["floor"] = _err(emu, ref)
        errs["va_floor"] = {"value": round(_err(emu[..., -nva:], ref[..., -nva:]), 5)}
    if "labels" in fx or out.shape[-1] >= 2:
        if "labels" in fx:
            labs = (fx["labels"]["label_valence"], fx["labels"]["label_arousal"])
        else:
            g = torch.Generator().manual_seed(99)
            labs = tuple(torch.rand(out.shape[:-1], generator=g) * 2 - 1 for _ in range(2))
        worst, info = 0.0, {}
        for j, lab in enumerate(labs):
            ch = out.shape[-1] - 2 + j
            c_ref = float(concordance_cc2(ref[..., ch].reshape(-1), lab.reshape(-1), "none"))
            c_gpu = float(concordance_cc2(out[..., ch].reshape(-1), lab.reshape(-1), "none"))
            worst = max(worst, abs(c_ref - c_gpu))
            info["ccc_ref_%d" % j], info["ccc_gpu_%d" % j] = round(c_ref, 6), round(c_gpu, 6)
            # agreement of the two prediction tracks themselves (CCC between CUDA and reference predictions)
            info["ccc_between_%d" % j] = round(float(concordance_cc2(out[..., ch].reshape(-1),
                                                                     ref[..., ch].reshape(-1), "none")), 6)
        info["frames"] = int(out.shape[0] * out.shape[1])
        if info["frames"] >= 1024:      # "CCC equal to 3 decimals" is asserted where the CCC is a statistic, not 32 points
            errs["ccc_ref_diff"] = worst
        else:
            info["ccc_ref_diff_small_sample"] = round(worst, 6)
        errs["info"] = info
    return errs


for _n, _kw in (("cfg1_va3dresnet_eval", {}), ("cfg1_va3dresnet_eval_hard", {}), ("cfg3_av_resnet_eval", {}),
                ("cfg3_av_resnet_eval_hard", {}), ("cfg3_av_v2psplit_eval", {}),
                ("cfg4_av_resnet_eval_256x16", {"chunk": 64}), ("cfg4_av_resnet_eval_256x16_hard", {"chunk": 64}),
                ("vggm_tcn_eval", {})):
    CASES["size_" + _n] = (case_golden_size, _c(name=_n, **_kw))
CASES["golden_vggm_tcn_train"] = (case_golden, _c(name="vggm_tcn_train"))
# the rest of the zoo (fixtures of oracle/make_golden_zoo.py)
CASES["golden_resnetv2_trunk_eval"] = (case_golden, _c(name="resnetv2_trunk_eval", grads=False))
CASES["golden_resnetv2_trunk_train"] = (case_golden, _c(name="resnetv2_trunk_train"))
CASES["golden_va3dresnet_v2_eval"] = (case_golden, _c(name="va3dresnet_v2_eval"))
CASES["golden_attencdec_eval"] = (case_golden, _c(name="attencdec_eval", grads=False))
CASES["golden_attencdec_train"] = (case_golden, _c(name="attencdec_train"))
CASES["golden_vggface_eval"] = (case_golden, _c(name="vggface_eval"))
CASES["golden_vggface_train"] = (case_golden, _c(name="vggface_train"))
CASES["golden_densenet_eval"] = (case_golden, _c(name="densenet_eval", grads=False))
CASES["golden_densenet_train"] = (case_golden, _c(name="densenet_train"))
CASES["golden_cbam_eval"] = (case_golden, _c(name="cbam_eval"))
CASES["golden_cbam_train"] = (case_golden, _c(name="cbam_train"))
CASES["golden_resnet_cbam_train"] = (case_golden, _c(name="resnet_cbam_train"))
CASES["golden_av_v2psplit_attention_train"] = (case_golden, _c(name="av_v2psplit_attention_train"))

for _t in ("", "_odd"):
    pass
TOLS = {"out_ref": 3e-2, "va_ref": 2e-2, "floor": 1.0, "out_emu": 1.5e-2, "loss_ref": 2e-2, "loss_emu": 1e-2,
        "grad_emu": 0.12, "grad_all_l2": 0.05, "dx": 3e-2, "ccc_ref_diff": 5e-4}
for _k in ("conv1.weight", "conv2.weight", "bn1.weight", "bn1.bias", "bn2.weight", "bn2.bias", "downsample.0.weight",
           "downsample.1.weight", "downsample.1.bias"):
    TOLS["d_" + _k] = 3e-2


# Whole-network train-mode gradients of the 20-layer trunk are ill-conditioned under bf16 storage: the bf16-emulating
# oracle itself is 14-15 % (relative L2 over all parameters, up to 31 % per tensor) away from the fp32 reference on
# these 16-frame fixtures, and any summation-order difference is amplified the same way (DESIGN.md "Numerics").  Their
# backward chain is asserted tightly by the well-conditioned block_* / convnd_* / golden_{gru,attfusion,tcn} cases;
# here the whole-network gradient only has to stay within that floor (all-parameter L2 < 0.3) and forward / loss
# parity is asserted at the normal tolerances.
# The same holds for the VGG-M stacks (5 train-mode BN layers, 3 max-pools whose argmax routes flip under bf16 rounding):
# on vggm_tcn_train / av_v2psplit_attention_train (128 frames) the emulated oracle is 18 % / 14 % (all-parameter L2)
# from the fp32 oracle, the CUDA path 8.7 % / 7.4 % from the emulated oracle (measured).
CHAOTIC_GRADS = {"golden_resnet_trunk_train", "golden_va3dresnet_train", "golden_av_resnet_attention_train",
                 "golden_resnetv2_trunk_train", "golden_densenet_train", "golden_resnet_cbam_train",
                 "golden_vggface_train",
                 "golden_vggm_tcn_train", "golden_av_v2psplit_attention_train",
                 "va3dresnet_96px_train", "va3dresnet_15frames_train", "va3dresnet_1clip_2frames_train"}


# measured floors (bf16-emulating oracle vs fp32 oracle, all-parameter gradient L2, CPU): vggface_train 0.131,
# resnetv2_trunk_train 0.142, resnet_cbam_train 0.106, densenet_train 0.408 (49 train-mode BatchNorms on 8 frames whose
# last stage is 3x3 pixels) -- the CUDA path has to stay inside the same band
CHAOTIC_L2 = {"golden_densenet_train": 0.6}
# train-mode forward of the 49-BatchNorm DenseNet on synthetic high-gain weights: the bf16-emulating oracle itself is
# 2.4e-2 from the reference (floor, 32-frame fixture); eval mode sits at 5e-3
CASE_TOLS = {"golden_densenet_train": {"out_emu": 3e-2, "out_ref": 4e-2}}


def failures(name, errs):
    """The error keys of `errs` that break their tolerance (shared by pytest and tests/gpu_probe.py)."""
    tols = dict(TOLS)
    tols.update(CASE_TOLS.get(name, {}))
    if name in CHAOTIC_GRADS:
        errs = {k: v for k, v in errs.items() if k != "grad_emu"}
        tols["grad_all_l2"] = CHAOTIC_L2.get(name, 0.3)
    return {k: v for k, v in errs.items() if not isinstance(v, dict) and (v != v or v >= tols.get(k, TOL))}



# ----------------------------------------------------------------------------------------------------------
# Well-conditioned backward checks: one BasicBlock in train mode (conv fprop/dgrad/wgrad + BN fwd/bwd + residual)
# against autograd through the bf16-emulating oracle.  (Whole-network train-mode gradients on the tiny golden
# batches are chaotic: rounding ONLY the conv weights to bf16 in an exact fp64 evaluation already moves them by
# 30-45 %, see DESIGN.md "Numerics".)
# ----------------------------------------------------------------------------------------------------------
def case_block(inplanes, planes, stride, N=16, HW=28, seed=0):
    import torch.nn as nn
    from m3t_b200.models.resnet import BasicBlock
    from oracle import ref_torch as R
    torch.manual_seed(seed)
    down = None
    if stride != 1 or inplanes != planes:
        down = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
    blk = BasicBlock(inplanes, planes, stride, down)
    spec = {k: tuple(v.shape) for k, v in blk.state_dict().items()}
    sd = R.synth_state_dict(spec, 21 + seed)
    blk.load_state_dict(sd)
    blk = blk.cuda().train()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn((N, inplanes, HW, HW), generator=g).relu()
    xc = x.cuda().requires_grad_(True)
    out = blk(xc)
    cot = torch.randn(tuple(out.shape), generator=g)
    (out * cot.cuda()).sum().backward()
    # oracle
    sdr = {("b." + k): (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
           for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    with R.bf16_emulation():
        o = R.basic_block(R.q(xr), sdr, "b", stride, True)
        (o * cot).sum().backward()
    errs = {"out_emu": _err(out, o), "dx": _l2(xc.grad, xr.grad)}
    for n, p in blk.named_parameters():
        errs["d_" + n] = _l2(p.grad, sdr["b." + n].grad)
    return errs


CASES.update({
    "block_64_64_s1": (case_block, _c(inplanes=64, planes=64, stride=1)),
    "block_64_128_s2": (case_block, _c(inplanes=64, planes=128, stride=2)),
    "block_256_512_s2_7x7": (case_block, _c(inplanes=256, planes=512, stride=2, N=32, HW=7)),
    "block_512_512_4x4": (case_block, _c(inplanes=512, planes=512, stride=1, N=64, HW=4)),
})



def case_convnd(kind, seed=0):
    """ops.ConvNdBNAct (VGG-M conv groups, tcn_simple) forward + backward in train mode vs the emulating oracle."""
    import torch.nn as nn
    from m3t_b200 import ops, raw
    from oracle import ref_torch as R
    g = torch.Generator().manual_seed(seed)
    if kind == "first":
        conv, bn = nn.Conv3d(3, 64, 3, stride=(1, 2, 2), padding=(1, 0, 0)), nn.BatchNorm3d(64)
        x = torch.randint(0, 256, (2, 3, 4, 24, 24), generator=g).float()
        cfg = dict(nd=3, k=(3, 2, 2), pad_lo=(1, 0, 0), pad_hi=(1, 0, 0), relu=True, pool=(2, 2, 0), s2d_first=True)
        pool_fn = lambda t: F.max_pool3d(t, (1, 2, 2), (1, 2, 2))    # noqa: E731
        fn, kw = F.conv3d, dict(stride=(1, 2, 2), padding=(1, 0, 0))
    elif kind == "mid_pool":
        conv, bn = nn.Conv3d(64, 128, 3, 1, padding=(1, 0, 0)), nn.BatchNorm3d(128)
        x = torch.randn((2, 64, 4, 11, 11), generator=g).relu()
        cfg = dict(nd=3, k=(3, 3, 3), pad_lo=(1, 0, 0), pad_hi=(1, 0, 0), relu=True, pool=(2, 2, 0))
        pool_fn = lambda t: F.max_pool3d(t, (1, 2, 2), (1, 2, 2))    # noqa: E731
        fn, kw = F.conv3d, dict(stride=1, padding=(1, 0, 0))
    elif kind == "mid_nopool":
        conv, bn = nn.Conv3d(256, 512, 3, 1, padding=(1, 0, 0)), nn.BatchNorm3d(512)
        x = torch.randn((3, 256, 4, 5, 5), generator=g).relu()
        cfg = dict(nd=3, k=(3, 3, 3), pad_lo=(1, 0, 0), pad_hi=(1, 0, 0), relu=True, pool=None)
        pool_fn = None
        fn, kw = F.conv3d, dict(stride=1, padding=(1, 0, 0))
    else:
        conv, bn = nn.Conv1d(1024, 512, 5, 1, 2), nn.BatchNorm1d(512)
        x = torch.randn((6, 1024, 16), generator=g).relu()
        cfg = dict(nd=1, k=(1, 1, 5), pad_lo=(0, 0, 2), pad_hi=(0, 0, 2), relu=True, pool=None)
        pool_fn = None
        fn, kw = F.conv1d, dict(padding=2)
    mods = nn.ModuleDict({"c": conv, "b": bn})
    sd = R.synth_state_dict({k: tuple(v.shape) for k, v in mods.state_dict().items()}, 31 + seed)
    mods.load_state_dict(sd)
    mods = mods.cuda().train()
    conv, bn = mods["c"], mods["b"]
    if kind == "first":
        xin = raw.video_prep_s2d(x.cuda(), True)
        xr_in = (x - 127.5) / 127.5
    else:
        xleaf = x.cuda().requires_grad_(True)
        xin = ops.ToCL.apply(xleaf)
        xr_in = x
    out = ops.ConvNdBNAct.apply(xin, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, cfg,
                                True)
    out_nc = ops.FromCL.apply(out)
    cot = torch.randn(tuple(out_nc.shape), generator=g)
    (out_nc * cot.cuda()).sum().backward()
    sdr = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
           for k, v in sd.items()}
    xr = xr_in.clone().requires_grad_(kind != "first")
    with R.bf16_emulation():
        o = R._conv_bias_bn_relu(xr, sdr["c.weight"], sdr["c.bias"], sdr, "b", True, fn, pool_fn, **kw)
        (o * cot).sum().backward()
    errs = {"out_emu": _err(out_nc, o), "d_conv_w": _l2(conv.weight.grad, sdr["c.weight"].grad),
            "d_bn_w": _l2(bn.weight.grad, sdr["b.weight"].grad), "d_bn_b": _l2(bn.bias.grad, sdr["b.bias"].grad)}
    if kind != "first":
        errs["dx"] = _l2(xleaf.grad, xr.grad)
    return errs


CASES.update({
    "convnd_first": (case_convnd, _c(kind="first")),
    "convnd_mid_pool": (case_convnd, _c(kind="mid_pool")),
    "convnd_mid_nopool": (case_convnd, _c(kind="mid_nopool")),
    "convnd_1d_k5": (case_convnd, _c(kind="conv1d")),
})
for _k in ("d_conv_w", "d_bn_w", "d_bn_b"):
    TOLS[_k] = 3e-2



def case_logmel(seed=0):
    """Fused log-Mel kernel vs the numpy oracle (oracle/melspec.py; parity unpinned w.r.t. librosa itself)."""
    import numpy as np
    from m3t_b200.process import extract_melspec as P
    from oracle import melspec as OM
    g = torch.Generator().manual_seed(seed)
    errs = {}
    for name, fps, n, mode in (("fps30", 30.0, 16000 * 3 + 123, "constant"), ("fps25_reflect", 25.0, 16000 * 2, "reflect")):
        t = torch.arange(n) / 16000.0
        y = 0.3 * torch.sin(2 * math.pi * 440 * t) + 0.05 * torch.randn(n, generator=g) * (t > 0.5) + 1e-4
        ref = OM.logmel(y.numpy(), fps, pad_mode=mode)
        got = P.melspectrogram_db(y.cuda(), fps, pad_mode=mode).cpu().numpy()
        assert got.shape == ref.shape, (got.shape, ref.shape)
        errs[name + "_db_abs"] = float(np.abs(got - ref).max())
        st = P.stack_audio_windows(torch.from_numpy(got).cuda(), 5, 40).cpu().numpy()
        errs[name + "_stack"] = float(np.abs(st - OM.stack_windows(got, 5, 40)).max())
    return errs


def case_logmel_golden():
    """Fused log-Mel kernel vs features recorded from torchaudio (tests/golden/logmel_torchaudio.pt)."""
    from m3t_b200.process import extract_melspec as P
    fx = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "logmel_torchaudio.pt"))
    errs = {}
    for c in fx:
        got = P.melspectrogram_db(c["wave"].cuda(), c["fps"], pad_mode=c["pad_mode"]).cpu()
        assert got.shape == c["logmel_db"].shape
        errs["ta_%s_db_abs" % c["pad_mode"]] = float((got - c["logmel_db"]).abs().max())
    return errs


CASES["logmel"] = (case_logmel, _c())
CASES["logmel_torchaudio_golden"] = (case_logmel_golden, _c())
TOLS["ta_constant_db_abs"] = 2e-2      # dB
TOLS["ta_reflect_db_abs"] = 2e-2
for _k in ("fps30_db_abs", "fps25_reflect_db_abs"):
    TOLS[_k] = 2e-2      # dB



def case_optim(seed=0):
    """Flat-arena clip + Adam (optim.cu) vs torch.nn.utils.clip_grad_norm_ + torch.optim.Adam over 3 steps."""
    import torch.nn as nn
    from m3t_b200.engine import TrainEngine

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.l1, self.l2 = nn.Linear(37, 53), nn.Linear(53, 5)

        def forward(self, batch):
            return self.l2(torch.tanh(self.l1(batch["x"])))

        def compute_loss(self, y, batch, sync_free=False):
            return ((y - batch["t"]) ** 2).sum(), {}

    torch.manual_seed(seed)
    a, b = Net().cuda(), Net().cuda()
    b.load_state_dict(a.state_dict())
    eng = TrainEngine(a, lr=1e-2, weight_decay=1e-4, clip=1.0)
    opt = torch.optim.Adam(b.parameters(), lr=1e-2, weight_decay=1e-4)
    g = torch.Generator().manual_seed(seed)
    for _ in range(3):
        batch = {"x": torch.randn(64, 37, generator=g).cuda() * 3, "t": torch.randn(64, 5, generator=g).cuda()}
        eng.step(batch)
        opt.zero_grad()
        b.compute_loss(b(batch), batch)[0].backward()
        nn.utils.clip_grad_norm_(b.parameters(), 1.0)
        opt.step()
    torch.cuda.synchronize()
    return {"p_" + n: _err(p, dict(b.named_parameters())[n]) for n, p in a.named_parameters()}


CASES["optim_adam_clip"] = (case_optim, _c())
# the halo-tile kernel is selected automatically for 64->64 3x3/s1/p1 (conv2d_3x3_s1_64 above); more shapes:
CASES["halo_28_fused"] = (case_conv, _c(nd=2, N=5, D=1, H=28, W=28, Cin=64, Cout=64, k=(1, 3, 3), stride=(1, 1, 1),
                                        pad_lo=(0, 1, 1), fused=True))
CASES["halo_28_stats_big"] = (case_conv, _c(nd=2, N=67, D=1, H=28, W=28, Cin=64, Cout=64, k=(1, 3, 3),
                                            stride=(1, 1, 1), pad_lo=(0, 1, 1), stats=True))
CASES["halo_odd_13x9"] = (case_conv, _c(nd=2, N=3, D=1, H=13, W=9, Cin=64, Cout=64, k=(1, 3, 3), stride=(1, 1, 1),
                                        pad_lo=(0, 1, 1), stats=True))
CASES["wgrad_halo_28"] = (case_conv, _c(nd=2, N=5, D=1, H=28, W=28, Cin=64, Cout=64, k=(1, 3, 3), stride=(1, 1, 1),
                                        pad_lo=(0, 1, 1), wgrad=True))
CASES["wgrad_halo_13x9"] = (case_conv, _c(nd=2, N=3, D=1, H=13, W=9, Cin=64, Cout=64, k=(1, 3, 3), stride=(1, 1, 1),
                                          pad_lo=(0, 1, 1), wgrad=True))
CASES["wgrad_halo_5x40_big"] = (case_conv, _c(nd=2, N=301, D=1, H=5, W=40, Cin=64, Cout=64, k=(1, 3, 3),
                                              stride=(1, 1, 1), pad_lo=(0, 1, 1), wgrad=True))
CASES["halo_5x40"] = (case_conv, _c(nd=2, N=7, D=1, H=5, W=40, Cin=64, Cout=64, k=(1, 3, 3), stride=(1, 1, 1),
                                    pad_lo=(0, 1, 1)))
for _k in ("p_l1.weight", "p_l1.bias", "p_l2.weight", "p_l2.bias"):
    TOLS[_k] = 1e-4



def case_rowshift():
    """Probe: UMMA K-major SW128 operand starting at a non-1024-aligned row of a TMA-written tile (round-2 halo conv).
    Reports, per shift, the error with base_offset = 0 (m0_*) and base_offset = (addr>>7)&7 (m1_*)."""
    from m3t_b200 import lib as L
    g = torch.Generator().manual_seed(0)
    A = _rnd((256, 64), g)
    B = _rnd((64, 64), g)
    Ad, Bd = A.cuda(), B.cuda()
    errs = {}
    for shift in (0, 8, 1, 2, 3, 7, 30, 31, 61):
        ref = A[shift:shift + 128].float() @ B.float().t()
        for mode in (0, 1):
            out = torch.zeros((128, 64), device="cuda")
            L.check(L.load().m3t_debug_rowshift(L.ptr(Ad), L.ptr(Bd), L.ptr(out), L.i32(shift), L.i32(mode),
                                                L.stream_ptr()), "rowshift")
            torch.cuda.synchronize()
            errs["m%d_s%d" % (mode, shift)] = _err(out, ref)
    return errs


CASES["probe_rowshift"] = (case_rowshift, _c())



def case_audio_resnet(train, seed=0):
    """BASELINE config 2 composition (audio ResNet over log-Mel windows + TCN head) vs the same composition built
    from the oracle's functions (bf16-emulated) — there is no reference class for it."""
    from m3t_b200.models.audio_resnet import AudioResNetTCN
    from oracle import ref_torch as R
    torch.manual_seed(seed)
    m = AudioResNetTCN(dropout=0.0)
    spec = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = R.synth_state_dict(spec, 41 + seed)
    m.load_state_dict(sd)
    m = m.cuda().train(train)
    g = torch.Generator().manual_seed(seed + 1)
    B, T = 8, 16
    audio = torch.randn((B, T, 200), generator=g) * 20 - 40
    feats = []
    tcn_cl = m.tcn.forward_cl

    def _capture(f):
        y = tcn_cl(f)
        feats.append(y.detach().float().cpu())
        return y

    m.tcn.forward_cl = _capture      # the fc input (the model calls ops.linear directly, so no module hook fires)
    out = m(audio.cuda())

    def oracle(emulate):
        import contextlib
        with (R.bf16_emulation() if emulate else contextlib.nullcontext()), torch.no_grad():
            x = audio.reshape(B * T, 1, 5, 40)
            h = R.q(R._conv_bn(R.q(x), sd["stem.0.weight"], sd, "stem.1", train, 1, 1).relu())
            f = R.resnet_trunk(h, sd, "resnet", train=train).view(B, T, 512)
            f = R.temporal_conv_net(f.transpose(1, 2), sd, "tcn", 2).transpose(1, 2)
            return R._linear(f, sd["fc.weight"], sd["fc.bias"], keep_f32=True), f

    (o32, _), (oemu, femu) = oracle(False), oracle(True)
    # The synthetic fc head cancels (|W||f| is ~3x |Wf|), so one flipped bf16 ulp in a feature is amplified in the
    # output: `feat_emu` (the fc INPUT, tight tolerance) is the implementation check, `head_emu` is the output after
    # that amplification and shares the tolerance of out_ref; `floor` is what bf16 storage alone costs the oracle.
    gain = float(((femu.reshape(-1, femu.shape[-1]).abs() @ sd["fc.weight"].abs().t()).max()) / oemu.abs().max())
    return {"out_ref": _err(out, o32), "floor": _err(oemu, o32), "feat_emu": _err(feats[0].reshape(femu.shape), femu),
            "head_emu": _err(out, oemu), "info": {"head_gain": gain}}


TOLS["feat_emu"] = 1.5e-2
TOLS["head_emu"] = 3e-2
CASES["audio_resnet_tcn_eval"] = (case_audio_resnet, _c(train=False))
CASES["audio_resnet_tcn_train_fwd"] = (case_audio_resnet, _c(train=True))



def case_stem_wgrad_halo(B=2, T=5, seed=0):
    """Halo-tile stem wgrad (activation box resident, temporal taps in two passes) vs the generic im2col wgrad and
    vs autograd.  B*T*14 K-blocks: the small case gives every CTA one block (taps with no valid frame stay unstarted),
    the large one makes CTAs accumulate over several frames and clips."""
    from m3t_b200 import raw
    g = torch.Generator().manual_seed(seed)
    H2, W2 = 56, 56
    xs = _rnd((B, T, H2, W2, 64), g).cuda()
    dy = _rnd((B * T, H2, W2, 64), g).cuda()
    geom = raw.conv_geom(3, B, T, H2, W2, 64, 64, (5, 4, 1), (1, 1, 1), (2, 2, 0), (2, 1, 0), (1, 1, 1))
    ref = raw.conv_wgrad(xs, dy.view(B, T, H2, W2, 64), geom)
    got = raw.wgrad_stem_halo(xs, dy)
    torch.cuda.synchronize()
    # independent CPU check on one output channel block via conv3d autograd
    x_nc = xs.float().cpu().permute(0, 4, 1, 2, 3)
    w = torch.zeros((64, 64, 5, 4, 1), requires_grad=True)
    y = F.conv3d(F.pad(x_nc, (0, 0, 2, 1, 2, 2)), w)
    y.backward(dy.float().cpu().view(B, T, H2, W2, 64).permute(0, 4, 1, 2, 3))
    dw_cpu = w.grad.permute(0, 2, 3, 4, 1).reshape(64, -1)
    return {"vs_generic": _err(got, ref), "vs_cpu": _err(got, dw_cpu)}


CASES["stem_wgrad_halo"] = (case_stem_wgrad_halo, _c())
CASES["stem_wgrad_halo_3x16"] = (case_stem_wgrad_halo, _c(B=3, T=16))
TOLS["vs_generic"] = 5e-3
TOLS["vs_cpu"] = 1e-4



def case_stem_fprop_halo(seed=0):
    """Halo-tile stem forward vs the generic im2col kernel on the same inputs (and their BN statistics)."""
    from m3t_b200 import raw
    g = torch.Generator().manual_seed(seed)
    B, T, H2, W2 = 3, 5, 56, 56
    xs = _rnd((B, T, H2, W2, 64), g)
    xs[..., 48:] = 0            # structural zeros of the W-unrolled layout: the halo kernel never multiplies them
    xs = xs.cuda()
    w = _rnd((64, 1280), g, scale=0.03).cuda()
    geom = raw.conv_geom(3, B, T, H2, W2, 64, 64, (5, 4, 1), (1, 1, 1), (2, 2, 0), (2, 1, 0), (1, 1, 1))
    st0 = torch.zeros((2, 64), device="cuda")
    st1 = torch.zeros((2, 64), device="cuda")
    ref = raw.conv_fprop(xs, w, geom, stats=st0).view(B * T, H2, W2, 64)
    got = raw.stem_fprop_halo(xs, w, stats=st1)
    torch.cuda.synchronize()
    # independent CPU check
    x_nc = xs.float().cpu().permute(0, 4, 1, 2, 3)
    wk = w.float().cpu().view(64, 5, 4, 64).permute(0, 3, 1, 2).unsqueeze(-1)       # [co, ci, kt, jh, 1]
    y_cpu = F.conv3d(F.pad(x_nc, (0, 0, 2, 1, 2, 2)), wk).permute(0, 2, 3, 4, 1).reshape(B * T, H2, W2, 64)
    return {"vs_generic": _err(got, ref), "out": _err(got, y_cpu), "sum": _err(st1[0], st0[0]),
            "sumsq": _err(st1[1], st0[1])}


CASES["stem_fprop_halo"] = (case_stem_fprop_halo, _c())


# 128 -> 128 at 14x14: the image-per-tile halo kernel (conv_halo128.cu); 5 images = one tile per CTA, 333 images =
# persistent CTAs over three tiles each (A double buffer, filter ring and TMEM phases wrap)
CASES["conv2d_c128_halo_stats"] = (case_conv, _c(nd=2, N=5, D=1, H=14, W=14, Cin=128, Cout=128, k=(1, 3, 3),
                                                 stride=(1, 1, 1), pad_lo=(0, 1, 1), stats=True))
CASES["conv2d_c128_halo_fused"] = (case_conv, _c(nd=2, N=5, D=1, H=14, W=14, Cin=128, Cout=128, k=(1, 3, 3),
                                                 stride=(1, 1, 1), pad_lo=(0, 1, 1), fused=True))
CASES["conv2d_c128_halo_many"] = (case_conv, _c(nd=2, N=333, D=1, H=14, W=14, Cin=128, Cout=128, k=(1, 3, 3),
                                                stride=(1, 1, 1), pad_lo=(0, 1, 1), stats=True))


# persistent tile walker (umma_persist.cuh), forced with tile_hint bit 4: many tiles per CTA (ring and TMEM phases wrap),
# two column blocks, statistics in registers across tiles, fused epilogue, one k-iteration per tile
CASES["conv2d_persist_stats"] = (case_conv, _c(nd=2, N=300, D=1, H=14, W=14, Cin=128, Cout=128, k=(1, 3, 3),
                                               stride=(1, 1, 1), pad_lo=(0, 1, 1), stats=True, tile_hint=16))
CASES["conv2d_persist_fused_512"] = (case_conv, _c(nd=2, N=700, D=1, H=7, W=7, Cin=256, Cout=512, k=(1, 3, 3),
                                                   stride=(1, 1, 1), pad_lo=(0, 1, 1), fused=True, tile_hint=16))
CASES["conv2d_persist_stats_512"] = (case_conv, _c(nd=2, N=700, D=1, H=7, W=7, Cin=256, Cout=512, k=(1, 3, 3),
                                                   stride=(1, 1, 1), pad_lo=(0, 1, 1), stats=True, tile_hint=16))
CASES["conv2d_persist_1x1_s2"] = (case_conv, _c(nd=2, N=40, D=1, H=28, W=28, Cin=64, Cout=128, k=(1, 1, 1),
                                                stride=(1, 2, 2), pad_lo=(0, 0, 0), stats=True, tile_hint=16))
CASES["conv2d_persist_mt1_64"] = (case_conv, _c(nd=2, N=30, D=1, H=28, W=28, Cin=64, Cout=64, k=(1, 3, 3),
                                                stride=(1, 1, 1), pad_lo=(0, 1, 1), stats=True, tile_hint=17))


def case_dgrad_s2(H, K, pad, Cin, Cout, N=5, seed=0):
    """Stride-2 data gradient by output parity (four small stride-1 convolutions of dY scattered into dX) vs the
    zero-inserted formulation and vs conv_transpose on the CPU."""
    from m3t_b200 import ops
    g = torch.Generator().manual_seed(seed)
    P = (H + 2 * pad - K) // 2 + 1
    w = (torch.randn((Cout, Cin, K, K), generator=g) * 0.05).cuda()
    dy = _rnd((N, P, P, Cout), g).cuda()
    ops.DGRAD_S2_PARITY = True
    dx1 = ops.conv2d_dgrad(dy, w, (N, H, H, Cin), 2, pad)
    ops.DGRAD_S2_PARITY = False
    dx0 = ops.conv2d_dgrad(dy, w, (N, H, H, Cin), 2, pad)
    ops.DGRAD_S2_PARITY = True
    torch.cuda.synchronize()
    wq = w.bfloat16().float().cpu()
    ref = F.conv_transpose2d(dy.float().cpu().permute(0, 3, 1, 2), wq, stride=2, padding=pad,
                             output_padding=H - ((P - 1) * 2 - 2 * pad + K))
    return {"vs_zero_insert": _err(dx1, dx0), "out": _err(dx1.float().cpu().permute(0, 3, 1, 2), ref)}


TOLS["vs_zero_insert"] = 1e-2
CASES["dgrad_s2_3x3_28"] = (case_dgrad_s2, _c(H=28, K=3, pad=1, Cin=64, Cout=128))
CASES["dgrad_s2_3x3_7"] = (case_dgrad_s2, _c(H=7, K=3, pad=1, Cin=256, Cout=512))
CASES["dgrad_s2_1x1_14"] = (case_dgrad_s2, _c(H=14, K=1, pad=0, Cin=128, Cout=256))
CASES["dgrad_s2_1x1_7"] = (case_dgrad_s2, _c(H=7, K=1, pad=0, Cin=256, Cout=512))


def case_postproc(seed=0):
    """Device evaluation post-processing (overlap-add, Wiener window 35, masked CCC) vs the golden outputs of the
    reference's validation_end / smooth_predictions / concordance_cc2_np, and vs the oracle on a larger ragged set."""
    import numpy as np
    from m3t_b200.process import postproc as PP
    from oracle import postproc as O
    fx = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "postproc.pt"))
    segs = fx["segs"]
    vids, starts, lens = [s[0] for s in segs], [s[1] for s in segs], [s[2] for s in segs]
    V = len(fx["lengths"])
    tp = PP.overlap_add(fx["preds"].cuda(), starts, vids, lens, fx["window"], V)
    tg = PP.overlap_add(fx["gts"].cuda(), starts, vids, lens, fx["window"], V)
    errs = {}
    errs["oadd_exact"] = float(max((a.cpu() != b).sum() for a, b in zip(tp.split(), fx["track_pred"])) +
                               max((a.cpu() != b).sum() for a, b in zip(tg.split(), fx["track_gt"])))
    sm = PP.smooth_predictions(tp, 35, mode="wiener")
    errs["wiener"] = max(float((a.cpu() - b).abs().max()) for a, b in zip(sm.split(), fx["smooth"]))
    pv, ov = PP.concordance_cc2_np(sm, tg)
    errs["ccc_video"] = float((pv.cpu() - fx["ccc_per_video"]).abs().max())
    errs["ccc_all"] = float((ov.cpu() - fx["ccc_overall"]).abs().max())
    # larger ragged set vs the oracle: 37 videos of 40..3000 frames, window 16, some invalid annotations
    rng = np.random.default_rng(seed)
    window, C = 16, 2
    lengths = [int(x) for x in rng.integers(40, 3000, size=37)]
    segs = []
    for v, n in enumerate(lengths):
        st = 0
        while True:
            ln = min(window, n - st)
            segs.append((v, st, ln))
            if st + ln >= n:
                break
            st += window // 2
    S = len(segs)
    preds = np.tanh(rng.standard_normal((S, window, C)).cumsum(1) * 0.2).astype(np.float32)
    gts = np.clip(preds * 0.8 + rng.standard_normal((S, window, C)) * 0.2, -1, 1).astype(np.float32)
    gts[rng.integers(0, S, size=50), rng.integers(0, window, size=50), rng.integers(0, C, size=50)] = -5.0
    order = rng.permutation(S)
    vids, starts, lens = [segs[i][0] for i in order], [segs[i][1] for i in order], [segs[i][2] for i in order]
    tp = PP.overlap_add(torch.from_numpy(preds[order]).cuda(), starts, vids, lens, window, len(lengths))
    tg = PP.overlap_add(torch.from_numpy(gts[order]).cuda(), starts, vids, lens, window, len(lengths))
    otp = O.overlap_add(preds[order], starts, vids, lens, window, len(lengths))
    otg = O.overlap_add(gts[order], starts, vids, lens, window, len(lengths))
    errs["oadd_exact_big"] = float(sum(int((a.cpu().numpy() != b).sum()) for a, b in zip(tp.split(), otp)))
    sm = PP.smooth_predictions(tp, 35, mode="wiener")
    pv, ov = PP.concordance_cc2_np(sm, tg)
    opv, oov = O.smoothed_ccc(otp, otg, 35)
    errs["wiener_big"] = max(float(np.abs(a.cpu().numpy()[:, c] - O.wiener(b[:, c], 35)).max())
                             for a, b in list(zip(sm.split(), otp))[:6] for c in range(C))
    errs["ccc_video_big"] = float(np.abs(pv.cpu().numpy() - opv).max())
    errs["ccc_all_big"] = float(np.abs(ov.cpu().numpy() - oov).max())
    return errs


CASES["postproc_eval"] = (case_postproc, _c())
# overlap-add: bit-exact (count of differing elements must be 0 -> below any positive tolerance); Wiener in float64:
# summation order only; CCC: the reference reduces the float32 ground truth in float32, the device in float64
for _k in ("oadd_exact", "oadd_exact_big"):
    TOLS[_k] = 0.5
for _k in ("wiener", "wiener_big"):
    TOLS[_k] = 1e-10
for _k in ("ccc_video", "ccc_all", "ccc_video_big", "ccc_all_big"):
    TOLS[_k] = 1e-6


def case_va3dresnet_shapes(B, T, HW, train, seed=0):
    """VA_3DResNet on shapes off the tuned path - frame counts that are not tile multiples, a 96x96 input (generic
    im2col stem and stem weight gradient instead of the 56-wide halo kernels; 24/12/6/3-pixel trunk maps) - vs the
    bf16-emulating oracle, forward and (train) all-parameter gradient L2."""
    from m3t_b200.models.backbone import VA_3DResNet
    from oracle import ref_torch as R
    torch.manual_seed(seed)
    m = VA_3DResNet(frameLen=T, nClasses=9, nFCs=2, resnet_ver="v1")
    spec = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = R.synth_state_dict(spec, 77 + seed)
    m.load_state_dict(sd)
    m = m.cuda().train(train)
    g = torch.Generator().manual_seed(seed + 5)
    x = torch.rand((B, 3, T, HW, HW), generator=g) * 2 - 1
    cot = torch.randn((B, T, 9), generator=g)
    with torch.set_grad_enabled(train):
        out = m(x.cuda())
        if train:
            (out * cot.cuda()).sum().backward()
    sdr = {k: v.clone().requires_grad_(train and v.is_floating_point() and not k.endswith(("running_mean", "running_var")))
           for k, v in sd.items()}
    with R.bf16_emulation(), torch.set_grad_enabled(train):
        ref = R.va_3dresnet(x, sdr, T, train=train)
        if train:
            (ref * cot).sum().backward()
    errs = {"out_emu": _err(out, ref)}
    if train:
        num = den = 0.0
        for k, p in m.named_parameters():
            if p.grad is None or sdr[k].grad is None:
                continue
            num += float((p.grad.float().cpu() - sdr[k].grad).pow(2).sum())
            den += float(sdr[k].grad.pow(2).sum())
        errs["grad_all_l2"] = (num / max(den, 1e-30)) ** 0.5
    return errs


CASES["va3dresnet_96px_eval"] = (case_va3dresnet_shapes, _c(B=2, T=3, HW=96, train=False))
CASES["va3dresnet_96px_train"] = (case_va3dresnet_shapes, _c(B=2, T=3, HW=96, train=True))
CASES["va3dresnet_15frames_train"] = (case_va3dresnet_shapes, _c(B=3, T=5, HW=112, train=True))
CASES["va3dresnet_1clip_1frame_eval"] = (case_va3dresnet_shapes, _c(B=1, T=1, HW=112, train=False))
CASES["va3dresnet_1clip_2frames_train"] = (case_va3dresnet_shapes, _c(B=1, T=2, HW=112, train=True))


def case_golden_fp32(name, terms=3):
    """fp32-parity inference mode (m3t_b200.fp32.parity_mode: float32 activations, split-operand bf16 tensor-core
    launches) on a golden fixture, against the fp32 output of the UNMODIFIED reference module.  North-star bar:
    1e-4 range-normalised on the outputs / the V-A predictions."""
    from m3t_b200 import fp32
    from tests.golden_util import load, ref_batch
    fx = load(name)
    m = _build(fx).eval()
    kind, inp = fx["kind"], fx["inputs"]
    with torch.no_grad(), fp32.parity_mode(terms=terms):
        if kind in ("GRU", "ResNet", "TemporalConvNet"):
            out = m(inp["x"].cuda())
        elif kind == "AttFusion":
            out = m(inp["x_a"].cuda(), inp["x_v"].cuda())
        elif kind in ("VA_3DResNet", "VA_3DVGGM"):
            out = m((inp["video_u8"].float().cuda() - 127.5) / 127.5)
        elif kind == "VA_3DVGGM_Split":
            se = inp["se_features"].cuda()
            out = m((inp["video_u8"].float().cuda() - 127.5) / 127.5, se, se)
        else:
            out = m(ref_batch(inp, "cuda"))
    sfx = "" if terms == 3 else "x6"
    errs = {"out_ref32" + sfx: _err(out, fx["out"])}
    if kind == "AffWild2VA":
        sl = slice(7, None) if "mtl" in fx["hparams"]["loss"] else slice(-2, None)
        errs["va_ref32" + sfx] = _err(out[..., sl], fx["out"][..., sl])
    return errs


for _n in ("gru_audio", "gru_scorer", "gru_nohead", "attfusion", "tcn", "resnet_trunk_eval", "va3dresnet_eval",
           "av_resnet_attention_eval", "vggm_split3_eval", "av_v2psplit_attention_eval"):
    CASES["fp32_" + _n] = (case_golden_fp32, _c(name=_n))
# six-term variant (three bf16 pieces per value): removes the 2^-16 representation / dropped-term error; what remains is
# the tensor core's per-MMA truncation of the fp32 accumulator on the main term (measured: V/A 5.0e-5 -> 1.1e-5 on the
# AV-ResNet fixture, no change on the K = 13824 VGG-M convolutions)
for _n in ("vggm_split3_eval", "av_resnet_attention_eval", "av_v2psplit_attention_eval"):
    CASES["fp32x6_" + _n] = (case_golden_fp32, _c(name=_n, terms=6))
TOLS["out_ref32"] = 1e-4
TOLS["va_ref32"] = 1e-4
TOLS["out_ref32x6"] = 5e-5
TOLS["va_ref32x6"] = 3e-5


def case_train_trajectory(steps=4, clips=4, lr=2e-4, seed=0):
    """Several optimisation steps end to end (TrainEngine: forward, ccc_mtl loss, backward, clip 1.0, Adam with weight
    decay 1e-4 through the flat arenas) vs the same loop on the oracle with bf16 storage emulation and
    torch.optim.Adam: the loss trajectory must agree step by step."""
    import bench as BN
    from m3t_b200.engine import TrainEngine
    from m3t_b200.models.model import AffWild2VA
    from oracle import ref_torch as R
    hp = BN.hparams()
    torch.manual_seed(12345 + seed)
    m = AffWild2VA(hp)
    BN.randomise_bn(m, 7)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    batch = BN.synth_batch(clips, 99 + seed, pin=False)
    # oracle loop
    params = []
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
            params.append(v)
    opt = torch.optim.Adam(params, lr=lr, weight_decay=1e-4)
    ref_losses = []
    for _ in range(steps):
        with R.bf16_emulation():
            y = R.affwild2va_forward(batch, sd, hp, train=True)
            loss = R.training_loss(y, batch, hp.loss, hp.loss_lambda)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        ref_losses.append(float(loss))
    # device loop
    m = m.cuda().train()
    eng = TrainEngine(m, lr=lr, weight_decay=1e-4, clip=1.0)
    dbatch = {k: v.cuda() for k, v in batch.items()}
    losses = [float(eng.step(dbatch)) for _ in range(steps)]
    errs = {"loss_step%d" % i: abs(a - b) / max(abs(b), 1e-6) for i, (a, b) in enumerate(zip(losses, ref_losses))}
    errs["info"] = {"gpu": [round(x, 5) for x in losses], "oracle": [round(x, 5) for x in ref_losses]}
    return errs


CASES["train_trajectory_4steps"] = (case_train_trajectory, _c())
for _i in range(4):
    TOLS["loss_step%d" % _i] = 1e-2


def case_cuda_graph(seed=0):
    """graphs.GraphedInference: the eval forward captured into a CUDA graph (tcgen05 / TMA / cooperative GRU launches
    included) replays bit-identically to the eager path, also on inputs it was not captured with; AV model (dict
    batch) and VA_3DResNet (tensor)."""
    import bench as BN
    from m3t_b200.graphs import GraphedInference
    from m3t_b200.models.backbone import VA_3DResNet
    from m3t_b200.models.model import AffWild2VA
    errs = {}
    torch.manual_seed(seed)
    m = VA_3DResNet(hiddenDim=512, frameLen=16, backend="gru", resnet_ver="v1", nClasses=9, nFCs=2).cuda().eval()
    BN.randomise_bn(m, 3)
    x = (torch.randint(0, 256, (2, 3, 16, 112, 112)).float().cuda() - 127.5) / 127.5
    x2 = torch.flip(x, dims=[0, 2]).contiguous()
    with torch.no_grad():
        y, y2 = m(x).clone(), m(x2).clone()
    g = GraphedInference(m, x)
    errs["graph_exact"] = float((g(x) != y).sum() + (g(x2) != y2).sum())
    hp = BN.hparams()
    m = AffWild2VA(hp).cuda().eval()
    BN.randomise_bn(m, 5)
    b1 = {k: v.cuda() for k, v in BN.synth_batch(2, 11, pin=False).items()}
    b2 = {k: v.cuda() for k, v in BN.synth_batch(2, 12, pin=False).items()}
    with torch.no_grad():
        y, y2 = m(b1).clone(), m(b2).clone()
    g = GraphedInference(m, b1)
    errs["graph_exact_av"] = float((g(b1) != y).sum() + (g(b2) != y2).sum())
    # ADVICE r1 (graphs.py): the graph reads the derived weight copies by address.  (a) dropping the cache and
    # allocating over the freed blocks must not disturb a replay (the instance holds the copies); (b) an in-place
    # weight update / an optimizer-arena rewrite must re-capture, not replay stale weights
    from m3t_b200 import ops
    ops._pack_cache.clear()
    junk = [torch.full((1 << 20,), 7.0, device="cuda") for _ in range(64)]    # reuse whatever was freed
    errs["graph_exact_after_cache_drop"] = float((g(b2) != y2).sum())
    del junk
    with torch.no_grad():
        for p_ in m.parameters():
            p_.mul_(1.01)
        y3 = m(b1).clone()
    n_cap = g.captures
    errs["graph_exact_after_update"] = float((g(b1) != y3).sum()) + (0.0 if g.captures == n_cap + 1 else 1.0)
    with torch.no_grad():
        for p_ in m.parameters():
            p_.data.mul_(0.99)               # no version bump ...
        ops.clear_caches()                   # ... but the engine announces arena rewrites this way
        y4 = m(b2).clone()
    errs["graph_exact_after_arena_step"] = float((g(b2) != y4).sum())
    return errs


CASES["cuda_graph_inference"] = (case_cuda_graph, _c())
TOLS["graph_exact"] = 0.5
TOLS["graph_exact_av"] = 0.5
for _k in ("graph_exact_after_cache_drop", "graph_exact_after_update", "graph_exact_after_arena_step"):
    TOLS[_k] = 0.5


def case_fullsize_properties(clips=256, seed=0):
    """Size-independent properties at the BASELINE config-4 size (256 clips x 16 frames, too large for the CPU oracle):
      * batch independence (eval): the predictions of the first 4 clips inside the 256-clip batch equal, bit for bit,
        those of the same 4 clips run alone (every output element has a fixed summation order; no cross-clip op);
      * determinism: two runs of the 256-clip forward are bit-identical;
      * exact power-of-two linearity of the tensor-core stem at full size: conv(2x) == 2 conv(x) in bf16;
      * training step: the loss of the 256-clip batch is finite and the gradient arena has no NaN / Inf."""
    import bench as BN
    from m3t_b200 import raw
    from m3t_b200.engine import TrainEngine
    from m3t_b200.models.model import AffWild2VA
    hp = BN.hparams()
    torch.manual_seed(12345 + seed)
    m = AffWild2VA(hp)
    BN.randomise_bn(m, 7)
    m = m.cuda().eval()
    batch = {k: v.cuda() for k, v in BN.synth_batch(clips, 77 + seed, pin=False).items()}
    small = {k: v[:4].contiguous() for k, v in batch.items()}
    errs = {}
    with torch.no_grad():
        big = m(batch)
        big2 = m(batch)
        sub = m(small)
    errs["batch_indep_exact"] = float((big[:4] != sub).sum())
    errs["determinism_exact"] = float((big != big2).sum())
    xs = raw.video_prep_s2d_w4(batch["video"], True)
    w = (torch.randn(64, 1280, device="cuda") * 0.03).bfloat16()
    y1 = raw.stem_fprop_halo(xs, w)
    y2 = raw.stem_fprop_halo((xs.float() * 2).bfloat16(), w)
    errs["pow2_linearity_exact"] = float(((y1.float() * 2).bfloat16().view(torch.int16) != y2.view(torch.int16)).sum())
    del xs, y1, y2, big, big2
    m.train()
    eng = TrainEngine(m, lr=5e-5, weight_decay=1e-4, clip=1.0)
    loss = eng.step(batch)
    errs["train_nonfinite"] = float((~torch.isfinite(eng.flat_g)).sum() + (~torch.isfinite(loss)).sum())
    return errs


CASES["fullsize_256clips_properties"] = (case_fullsize_properties, _c())
for _k in ("batch_indep_exact", "determinism_exact", "pow2_linearity_exact", "train_nonfinite"):
    TOLS[_k] = 0.5


def case_video_input(seed=0):
    """On-device input pipeline: m3t_video_augment_prep_s2d_w4 on decoded uint8 frames + parameter rows vs the layout
    pass applied to the clips the reference's load_video produced (golden), bit for bit; and the visual stream fed
    either way gives identical features."""
    from m3t_b200 import ops, raw
    from m3t_b200.models.backbone import VA_3DResNet
    clips = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "video_input.pt"))
    frames = torch.stack([c["frames"] for c in clips]).cuda()                       # [B,T,128,128,3]
    params = torch.tensor([c["params"] for c in clips], dtype=torch.int32).cuda()
    seq = torch.stack([c["seq"].float() for c in clips]).cuda()                     # [B,3,T,112,112]
    errs = {}
    for norm in (True, False):
        a = raw.video_augment_prep_s2d_w4(frames, params, 112, 112, norm)
        b = raw.video_prep_s2d_w4(seq, norm)
        errs["prep_exact_norm%d" % norm] = float((a.view(torch.int16) != b.view(torch.int16)).sum())
    torch.manual_seed(seed)
    m = VA_3DResNet(frameLen=2, nClasses=9, nFCs=2, resnet_ver="v1").cuda().eval()
    with torch.no_grad():
        f0 = m.features_cl(seq, True)
        f1 = m.features_cl(ops.RawClips(frames, params), True)
    errs["features_exact"] = float((f0.view(torch.int16) != f1.view(torch.int16)).sum())
    return errs


CASES["video_input_pipeline"] = (case_video_input, _c())
for _k in ("prep_exact_norm1", "prep_exact_norm0", "features_exact"):
    TOLS[_k] = 0.5


# ----------------------------------------------------------------------------------------------------------
# the scripts' call surface end to end — Trainer.fit / Trainer.test on a synthetic Aff-Wild2 tree
# ----------------------------------------------------------------------------------------------------------
def case_trainer_scripts(seed=0):
    """`python -m m3t_b200.run <script> ...` (tests/scripts/fit_script.py has the imports and calls of the reference's
    train.py / eval.py; on the build box the reference's own files are used instead when present): 2 epochs of
    audio-visual training on the synthetic tree of tests/synth_affwild.py (uint8 frames + on-device augmentation), the
    checkpoint it leaves, `--test_on_val` evaluation from that checkpoint, and the hooks run in-process on the same
    weights: validation_end's overlap-added tracks vs the oracle's overlap-add (bit-exact) and vs the file the
    evaluation script wrote (bit-exact: the eval forward is deterministic)."""
    import argparse
    import glob
    import subprocess
    import tempfile

    import numpy as np
    from oracle import postproc as O
    from tests import synth_affwild
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tmp = tempfile.mkdtemp(prefix="m3t_trainer_")
    data = os.path.join(tmp, "data")
    videos = synth_affwild.build(data, tmp, input_size=128, seed=seed)
    ref_root = os.environ.get("M3T_REFERENCE", "/root/reference")
    have_ref = os.path.isfile(os.path.join(ref_root, "train.py"))
    own = os.path.join(root, "tests", "scripts", "fit_script.py")
    common = ["--gpus", "0", "--modality", "audiovisual", "--backbone", "resnet", "--fusion_type", "attention",
              "--split_layer", "5", "--window", "8", "--windows_per_epoch", "4", "--batch_size", "4",
              "--dataset_path", data, "--release", "vipl", "--input_size", "128", "--workers", "0",
              "--checkpoint_path", tmp, "--device_augment"]
    env = dict(os.environ, PYTHONPATH=root)

    def run(script, extra):
        r = subprocess.run([sys.executable, "-m", "m3t_b200.run", script] + common + extra, cwd=tmp, env=env,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (script, r.stdout[-1500:], r.stderr[-3000:])
        return r.stdout + r.stderr

    errs = {}
    fit_log = run(os.path.join(ref_root, "train.py") if have_ref else own, ["--max_nb_epochs", "2", "--cutout"])
    ck = glob.glob(os.path.join(tmp, "lightning_logs", "version_0", "checkpoints", "_ckpt_epoch_*.ckpt"))
    errs["one_checkpoint"] = float(len(ck) != 1)
    errs["val_logged"] = float("val_ccc_v" not in fit_log or "fused flat-arena Adam" not in fit_log)
    run(os.path.join(ref_root, "eval.py") if have_ref else own,
        ["--checkpoint", ck[0], "--test_on_val"] + ([] if have_ref else ["--evaluate"]))
    by_script = torch.load(os.path.join(tmp, "predictions_val.pt"))
    val_videos = [v for v, d in videos.items() if d["split"] == "val"]
    errs["track_lengths"] = float(any(len(by_script["valence_pred"][v]) != videos[v]["frames"] for v in val_videos))
    errs["nonfinite"] = float(sum(int((~torch.isfinite(t)).sum()) for d in by_script.values() for t in d.values()))

    # the same weights in-process: hooks vs the oracle's overlap-add and vs the script's file
    from m3t_b200.models.model import AffWild2VA
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        model = AffWild2VA.load_from_checkpoint(ck[0])
        model.hparams.test_on_val = True
        model = model.cuda().eval()
        outputs = []
        with torch.no_grad():
            for bi, batch in enumerate(model.val_dataloader()[0]):
                batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
                outputs.append(model.validation_step(batch, bi))
            res = model.validation_end(outputs)
            in_proc = torch.load("predictions_val.pt")
            model.hparams.test_on_val = False
            test_out = [model.test_step({k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}, i)
                        for i, b in enumerate(model.test_dataloader()[0])]
            model.test_end(test_out)
            test_file = torch.load("predictions_test.pt")
    finally:
        os.chdir(cwd)
    errs["val_loss_nonfinite"] = float(not bool(torch.isfinite(res["val_loss"])))
    diff = 0
    for key in by_script:
        for v in val_videos:
            diff += int((by_script[key][v] != in_proc[key][v]).sum())
    errs["script_vs_inproc_exact"] = float(diff)
    names = [n for o in outputs for n in o["vid_names"]]
    starts = [int(s) for o in outputs for s in o["start_frames"]]
    vid_ix = {v: i for i, v in enumerate(dict.fromkeys(names))}
    segs = [torch.stack([o[k][j] for k in ("v_gt", "a_gt", "v_pred", "a_pred")], -1) for o in outputs
            for j in range(len(o["vid_names"]))]
    lens = [len(s) for s in segs]
    padded = np.zeros((len(segs), 8, 4), dtype=np.float32)
    for i, s in enumerate(segs):
        padded[i, :len(s)] = s.numpy()
    want = O.overlap_add(padded, starts, [vid_ix[n] for n in names], lens, 8, len(vid_ix))
    diff = 0
    for v, i in vid_ix.items():
        for c, key in enumerate(("valence_gt", "arousal_gt", "valence_pred", "arousal_pred")):
            diff += int((in_proc[key][v].numpy() != want[i][:, c]).sum())
    errs["val_tracks_exact"] = float(diff)
    tv = [v for v, d in videos.items() if d["split"] == "test"][0]
    errs["test_track_length"] = float(len(test_file["valence_pred"][tv]) != videos[tv]["frames"])
    return errs


CASES["trainer_scripts_fit_eval"] = (case_trainer_scripts, _c())
for _k in ("one_checkpoint", "val_logged", "track_lengths", "nonfinite", "val_loss_nonfinite",
           "script_vs_inproc_exact", "val_tracks_exact", "test_track_length"):
    TOLS[_k] = 0.5


# ----------------------------------------------------------------------------------------------------------
# Inference forward with its independent recurrent branches forked onto side streams (m3t_b200.streams)
# ----------------------------------------------------------------------------------------------------------
def case_streams(seed=0):
    """The multi-stream inference forward is bit-identical to the single-stream one (attention / concat fusion,
    ResNet and VGG-M split backbones, repeated calls, CUDA-graph capture of the forked forward); `info` reports the
    latency of a BASELINE config-5 shaped forward (16 clips x T=64) both ways."""
    import argparse

    import bench as BN
    from m3t_b200 import streams
    from m3t_b200.graphs import GraphedInference
    from m3t_b200.models.model import AffWild2VA
    errs, info = {}, {}

    def hp(**kw):
        d = dict(backbone="resnet", backend="gru", modality="audiovisual", fusion_type="attention", window=8,
                 loss="ccc_mtl", loss_lambda=0.5, num_hidden=512, split_layer=5, num_fc_layers=2, learning_rate=5e-5,
                 optimizer="adam")
        d.update(kw)
        return argparse.Namespace(**d)

    def batch(B, T, s):
        g = torch.Generator().manual_seed(s)
        return {"video": torch.randint(0, 256, (B, 3, T, 112, 112), generator=g, dtype=torch.uint8).float().cuda(),
                "audio": (torch.randn((B, T, 200), generator=g) * 20 - 40).cuda(),
                "se_features": torch.randn((B, 512, T), generator=g).cuda()}

    prev = streams.set_enabled(True)
    try:
        for tag, h in (("resnet_att", hp()), ("resnet_concat", hp(fusion_type="concat")),
                       ("v2psplit_att", hp(backbone="v2p_split", split_layer=3))):
            torch.manual_seed(seed)
            m = AffWild2VA(h)
            BN.randomise_bn(m, 5)
            m = m.cuda().eval()
            diff = 0
            with torch.no_grad():
                for s in (1, 2, 3):
                    b = batch(4, 8, s)
                    streams.set_enabled(False)
                    y0 = m(b).clone()
                    streams.set_enabled(True)
                    y1 = m(b).clone()
                    diff += int((y0 != y1).sum()) + int((~torch.isfinite(y1)).sum())
            errs["streams_exact_" + tag] = float(diff)
            if tag == "resnet_att":
                b = batch(4, 8, 9)
                streams.set_enabled(False)
                with torch.no_grad():
                    y0 = m(b).clone()
                streams.set_enabled(True)
                g = GraphedInference(m, batch(4, 8, 8))
                errs["streams_graph_exact"] = float((g(b) != y0).sum())
        # latency at BASELINE config-5 shapes (16 clips per GPU)
        for T in (64, 256):
            torch.manual_seed(seed)
            m = AffWild2VA(hp(window=T))
            BN.randomise_bn(m, 5)
            m = m.cuda().eval()
            b = batch(16, T, 4)
            for flag in (False, True):
                streams.set_enabled(flag)
                with torch.no_grad():
                    for _ in range(3):
                        m(b)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(10):
                        m(b)
                    e1.record()
                    torch.cuda.synchronize()
                info["ms_16x%d_streams_%s" % (T, "on" if flag else "off")] = round(e0.elapsed_time(e1) / 10, 3)
            del m, b
    finally:
        streams.set_enabled(prev)
    errs["info"] = info
    return errs


CASES["streams_inference_exact"] = (case_streams, _c())
for _k in ("streams_exact_resnet_att", "streams_exact_resnet_concat", "streams_exact_v2psplit_att",
           "streams_graph_exact"):
    TOLS[_k] = 0.5


def case_smooth_predictions_api(seed=0):
    """models.utils.smooth_predictions with the reference's call conventions (get_smoothed_ccc.py:15-16: a float32
    torch track, window 35; create_submission.py:35-36: a float32 ndarray, default window 13) against
    scipy.signal.wiener applied the way the reference applies it, and the script's CCC lines on top."""
    import numpy as np
    from scipy.signal import wiener
    from m3t_b200.models.utils import concordance_cc2_np, smooth_predictions
    rng = np.random.default_rng(seed)
    errs = {"wiener_api": 0.0, "ccc_api": 0.0}
    for n in (37, 400, 5000):
        pred = torch.from_numpy(np.tanh(rng.standard_normal(n).cumsum() * 0.1).astype(np.float32))
        gt = np.clip(pred.numpy() * 0.7 + rng.standard_normal(n).astype(np.float32) * 0.2, -1, 1)
        gt[rng.integers(0, n, 3)] = -5.0
        for window in (35, 13):
            got = smooth_predictions(pred, window, mode="wiener") if window == 35 else smooth_predictions(pred.numpy())
            want = np.apply_along_axis(lambda x: wiener(x, window), 0, pred.numpy())
            assert got.dtype == np.float64 and got.shape == want.shape
            errs["wiener_api"] = max(errs["wiener_api"], float(np.abs(got - want).max()))
            valid = gt >= -1
            errs["ccc_api"] = max(errs["ccc_api"], abs(float(concordance_cc2_np(got[valid], gt[valid])) -
                                                       float(concordance_cc2_np(want[valid], gt[valid]))))
    two = np.tanh(rng.standard_normal((300, 2)).cumsum(0) * 0.1).astype(np.float32)
    want = np.apply_along_axis(lambda x: wiener(x, 35), 0, two)
    errs["wiener_api"] = max(errs["wiener_api"], float(np.abs(smooth_predictions(two, 35) - want).max()))
    return errs


CASES["smooth_predictions_api"] = (case_smooth_predictions_api, _c())
TOLS["wiener_api"] = 1e-10
TOLS["ccc_api"] = 1e-10


def case_dropout(seed=0):
    """m3t_dropout_bf16 vs oracle/dropout.py: the mask bit for bit, the scaled values (bf16 rounding only), the
    backward pass (same mask on the gradient), and a TemporalBlock training step routed through it."""
    import numpy as np
    from m3t_b200 import ops, raw
    from oracle import dropout as D
    g = torch.Generator().manual_seed(seed)
    errs = {}
    x = (torch.randn((37, 24, 512), generator=g) + 3.0).bfloat16()        # no zeros: the mask is readable from y
    for p, sd in ((0.2, 12345), (0.5, 2 ** 61 + 7), (0.0, 1)):
        y = raw.dropout_bf16(x.cuda(), p, sd).cpu()
        keep = D.keep_mask(sd, x.numel(), p).reshape(tuple(x.shape))
        want = torch.from_numpy(D.dropout(x.float().numpy(), p, sd)).bfloat16()
        errs["mask_exact_p%g" % p] = float(((y != 0).numpy() != keep).sum())
        errs["values_p%g" % p] = _err(y.float(), want.float())
    torch.manual_seed(seed)
    xr = x.cuda().requires_grad_(True)
    y = ops.DropoutFn.apply(xr, 0.2)
    dy = torch.ones_like(y)
    y.backward(dy)
    errs["bwd_mask_exact"] = float(((xr.grad != 0) != (y != 0)).sum())
    errs["bwd_scale"] = abs(float(xr.grad.float().max()) - 1.25) / 1.25
    errs["keep_rate"] = abs(float((y != 0).float().mean()) - 0.8)
    return errs


CASES["dropout_bf16"] = (case_dropout, _c())
for _k in ("mask_exact_p0.2", "mask_exact_p0.5", "mask_exact_p0", "bwd_mask_exact"):
    TOLS[_k] = 0.5
for _k in ("values_p0.2", "values_p0.5", "values_p0", "bwd_scale"):
    TOLS[_k] = 4e-3                # one bf16 rounding of x * scale
TOLS["keep_rate"] = 5e-3           # 454 656 draws: sigma = 6e-4


def case_gru_cluster(seed=0):
    """m3t_gru_fwd_cluster (thread-block cluster + DSMEM exchange, B <= 64) against m3t_gru_fwd on the same operands:
    bf16 and fp32 outputs bit for bit (same operand values, k order and gate arithmetic), all three hidden sizes of
    the model, ragged batch / sequence sizes; `info` = microseconds per time step of both kernels."""
    from m3t_b200 import raw
    g = torch.Generator().manual_seed(seed)
    errs, info = {"gru_cluster_exact": 0.0, "gru_cluster_f32_exact": 0.0, "gru_cluster_nonfinite": 0.0}, {}
    for B, T, H in ((16, 64, 512), (2, 16, 512), (7, 33, 256), (16, 40, 128), (1, 5, 512), (16, 256, 512),
                    (32, 32, 512), (40, 9, 256), (64, 16, 128), (100, 7, 512), (128, 16, 512), (128, 16, 256),
                    (112, 5, 128)):
        gi = (torch.randn((B * T, 6 * H), generator=g) * 0.8).cuda()
        w = (torch.randn((2, 3 * H, H), generator=g) / H ** 0.5).bfloat16().cuda()
        bh = (torch.randn((2, 3 * H), generator=g) * 0.1).cuda()
        a, a32, sa = raw.gru_fwd(gi, w, bh, B, T, H, True, want_f32=True, cluster=False)
        c, c32, sc = raw.gru_fwd(gi, w, bh, B, T, H, True, want_f32=True, cluster=True)
        torch.cuda.synchronize()
        errs["gru_cluster_saved_exact"] = errs.get("gru_cluster_saved_exact", 0.0) + float((sa != sc).sum())
        errs["gru_cluster_exact"] += float((a.view(torch.int16) != c.view(torch.int16)).sum())
        errs["gru_cluster_f32_exact"] += float((a32 != c32).sum())
        errs["gru_cluster_nonfinite"] += float((~torch.isfinite(c32)).sum())
        for tag, flag in (("l2", False), ("cluster", True)):
            for _ in range(2):
                raw.gru_fwd(gi, w, bh, B, T, H, False, cluster=flag)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                raw.gru_fwd(gi, w, bh, B, T, H, False, cluster=flag)
            e1.record()
            torch.cuda.synchronize()
            info["us_per_step_%s_B%d_T%d_H%d" % (tag, B, T, H)] = round(e0.elapsed_time(e1) / 5 / T * 1e3, 2)
    errs["info"] = info
    return errs


CASES["gru_cluster_exact"] = (case_gru_cluster, _c())
for _k in ("gru_cluster_exact", "gru_cluster_f32_exact", "gru_cluster_nonfinite", "gru_cluster_saved_exact"):
    TOLS[_k] = 0.5


def case_gru_bwd_cluster(seed=0):
    """m3t_gru_bwd_cluster (K-split recurrent product, DSMEM reduce-scatter) against m3t_gru_bwd on the same operands:
    hprev bit for bit, gate gradients / bias gradients to the fp32 summation-order noise of the recurrent product
    (the bf16 outputs then differ by single roundings), run-to-run bit-identical (rank-order reduction), ragged batch /
    sequence sizes incl. several 16-row slices per cluster; `info` = microseconds per time step of both kernels."""
    from m3t_b200 import raw
    g = torch.Generator().manual_seed(seed)
    errs, info = {"gru_bwd_cluster_hprev_exact": 0.0, "gru_bwd_cluster_rerun_bits": 0.0, "gru_bwd_cluster_dgi": 0.0,
                  "gru_bwd_cluster_dgh": 0.0, "gru_bwd_cluster_dbias": 0.0, "gru_bwd_cluster_nonfinite": 0.0}, {}
    for B, T, H in ((16, 16, 512), (32, 16, 512), (2, 16, 512), (7, 33, 256), (16, 40, 128), (1, 5, 512), (3, 1, 256),
                    (40, 9, 256), (64, 16, 128), (256, 16, 512), (256, 16, 128), (100, 7, 512)):
        gi = (torch.randn((B * T, 6 * H), generator=g) * 0.8).cuda()
        w = (torch.randn((2, 3 * H, H), generator=g) / H ** 0.5).bfloat16().cuda()
        wt = w.transpose(1, 2).contiguous()
        bh = (torch.randn((2, 3 * H), generator=g) * 0.1).cuda()
        out, _, saved = raw.gru_fwd(gi, w, bh, B, T, H, True, want_f32=False, cluster=False)
        dout = torch.randn((B, T, 2 * H), generator=g).bfloat16().cuda()
        ref = raw.gru_bwd(dout, out, saved, wt, B, T, H, cluster=False)
        got = raw.gru_bwd(dout, out, saved, wt, B, T, H, cluster=True)
        got2 = raw.gru_bwd(dout, out, saved, wt, B, T, H, cluster=True)
        torch.cuda.synchronize()
        errs["gru_bwd_cluster_hprev_exact"] += float((ref[2].view(torch.int16) != got[2].view(torch.int16)).sum())
        errs["gru_bwd_cluster_rerun_bits"] += float(sum((a.view(torch.int16) != b.view(torch.int16)).sum()
                                                         for a, b in zip(got[:3], got2[:3])))
        errs["gru_bwd_cluster_dgi"] = max(errs["gru_bwd_cluster_dgi"], _l2(got[0].float(), ref[0].float()))
        errs["gru_bwd_cluster_dgh"] = max(errs["gru_bwd_cluster_dgh"], _l2(got[1].float(), ref[1].float()))
        errs["gru_bwd_cluster_dbias"] = max(errs["gru_bwd_cluster_dbias"], _l2(got[3], ref[3]))
        errs["gru_bwd_cluster_nonfinite"] += float(sum((~torch.isfinite(x.float())).sum() for x in got))
        for tag, flag in (("l2", False), ("cluster", True)):
            for _ in range(2):
                raw.gru_bwd(dout, out, saved, wt, B, T, H, cluster=flag)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                raw.gru_bwd(dout, out, saved, wt, B, T, H, cluster=flag)
            e1.record()
            torch.cuda.synchronize()
            info["us_per_step_%s_B%d_T%d_H%d" % (tag, B, T, H)] = round(e0.elapsed_time(e1) / 5 / T * 1e3, 2)
    from m3t_b200 import lib
    info["max_resident_clusters"] = {H: int(lib.load().m3t_gru_bwd_cluster_max(H)) for H in (128, 256, 512)}
    errs["info"] = info
    return errs


CASES["gru_bwd_cluster"] = (case_gru_bwd_cluster, _c())
TOLS["gru_bwd_cluster_hprev_exact"] = 0.5
TOLS["gru_bwd_cluster_rerun_bits"] = 0.5
TOLS["gru_bwd_cluster_nonfinite"] = 0.5
TOLS["gru_bwd_cluster_dgi"] = 2e-3        # bf16 outputs of fp32 values that differ in the last bits: single roundings
TOLS["gru_bwd_cluster_dgh"] = 2e-3
TOLS["gru_bwd_cluster_dbias"] = 5e-4


# ----------------------------------------------------------------------------------------------------------
# The whole training step as one CUDA-graph launch (engine.TrainEngine.capture): same trajectory as the eager step,
# scheduler changes of lr reach the replays, launch count per replay = 0
# ----------------------------------------------------------------------------------------------------------
def case_train_graph(steps=5, clips=4, seed=0):
    import bench as BN
    from m3t_b200 import lib
    from m3t_b200.engine import TrainEngine
    from m3t_b200.models.model import AffWild2VA
    hp = BN.hparams()
    batches = [{k: v.cuda() for k, v in BN.synth_batch(clips, 100 + i, pin=False).items()} for i in range(steps)]
    lrs = [2e-4, 2e-4, 1e-4, 1e-4, 5e-5][:steps]

    def run(graph):
        torch.manual_seed(seed)
        m = AffWild2VA(hp)
        BN.randomise_bn(m, 7)
        m = m.cuda().train()
        eng = TrainEngine(m, lr=lrs[0], weight_decay=1e-4, clip=1.0)
        losses, launches = [], []
        if graph:
            # capture() runs warm-up steps on its example batch: rewind the model / optimiser state afterwards so both
            # arms start from the same point
            sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
            eng.capture(batches[0], warmup=2)
            with torch.no_grad():
                for k, v in m.state_dict().items():
                    v.copy_(sd0[k])
            eng.m.zero_()
            eng.v.zero_()
            eng.set_step_count(0)
            from m3t_b200 import ops
            ops.clear_caches()
        for b, lr in zip(batches, lrs):
            eng.lr = lr
            n0 = lib.launch_count()
            losses.append(float(eng.step(b)))
            launches.append(lib.launch_count() - n0)
        return losses, launches, {k: v.detach().float().cpu().clone() for k, v in m.state_dict().items()}

    def dist(a, b):
        num = sum(float((a[k] - b[k]).pow(2).sum()) for k in b if b[k].is_floating_point())
        den = sum(float((b[k]).pow(2).sum()) for k in b if b[k].is_floating_point())
        return (num / den) ** 0.5

    le, ne, sde = run(False)
    le2, _, sde2 = run(False)     # a training step is not bit-reproducible (fp32 atomics): the
    lg, ng, sdg = run(True)       # eager-vs-eager distance on this high-gain synthetic model is the yardstick
    errs = {"graph_loss_step%d" % i: abs(a - b) / max(abs(b), 1e-6) for i, (a, b) in enumerate(zip(lg, le))}
    noise = dist(sde2, sde)
    errs["graph_params_l2"] = dist(sdg, sde)
    errs["graph_vs_noise"] = errs["graph_params_l2"] / max(noise, 1e-5)
    errs["graph_replay_launches"] = float(max(ng))
    errs["info"] = {"eager": [round(x, 5) for x in le], "eager_again": [round(x, 5) for x in le2],
                    "graph": [round(x, 5) for x in lg], "eager_vs_eager_params_l2": noise,
                    "eager_launches_per_step": ne[-1], "graph_launches_per_step": ng[-1]}
    return errs


CASES["train_graph_step"] = (case_train_graph, _c())
for _i in range(5):
    TOLS["graph_loss_step%d" % _i] = 1e-2       # run-to-run noise of this model is ~2e-3 (info: eager vs eager_again)
TOLS["graph_params_l2"] = 1e-2
TOLS["graph_vs_noise"] = 5.0
TOLS["graph_replay_launches"] = 0.5


# ----------------------------------------------------------------------------------------------------------
# Fused TemporalBlock (ops.TCNConvFn / m3t_tcn_conv_bf16) IN TRAINING MODE WITH DROPOUT against the oracle evaluated with
# the same mask stream (oracle/dropout.py): forward, input gradient and every parameter gradient
# ----------------------------------------------------------------------------------------------------------
def case_tcn_block_dropout(cin=512, cout=512, dilation=2, p=0.2, B=6, T=40, seed=0):
    from m3t_b200.models.tcn import TemporalBlock
    from oracle import ref_torch as R
    torch.manual_seed(seed)
    blk = TemporalBlock(cin, cout, 3, stride=1, dilation=dilation, padding=2 * dilation, dropout=p)
    spec = {k: tuple(v.shape) for k, v in blk.state_dict().items()}
    sd = R.synth_state_dict(spec, 41)
    blk.load_state_dict(sd)
    blk = blk.cuda().train()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn((B, cin, T), generator=g)
    cot = torch.randn((B, cout, T), generator=g)
    from m3t_b200 import raw
    used, orig = [], raw.tcn_conv

    def spy(*a, **kw):                 # the seeds the module really drew (torch's CPU generator)
        used.append(int(kw.get("seed", 0)))
        return orig(*a, **kw)

    raw.tcn_conv = spy
    try:
        xg = x.cuda().requires_grad_(True)
        out = blk(xg)
    finally:
        raw.tcn_conv = orig
    (out * cot.cuda()).sum().backward()
    seeds = used[-2:]
    # the oracle reads the values the module really holds (`net.{0,4}.*` alias `conv{1,2}.*` inside one parameter)
    sdo = {"b." + k: v.detach().float().cpu().clone().requires_grad_(v.is_floating_point())
           for k, v in blk.state_dict().items()}
    xo = x.clone().requires_grad_(True)
    with R.bf16_emulation():
        ref = R.temporal_block(xo, sdo, "b", dilation, dropout=(p, seeds[0], seeds[1]))
        (ref * cot).sum().backward()
    errs = {"out_emu": _err(out, ref), "dx": _l2(xg.grad, xo.grad)}
    zeros_gpu = float((out.detach().float().cpu() == 0).float().mean())
    zeros_ref = float((ref.detach() == 0).float().mean())
    errs["zero_fraction_diff"] = abs(zeros_gpu - zeros_ref)
    params = dict(blk.named_parameters())
    for k in ("conv1.weight_v", "conv1.weight_g", "conv1.bias", "conv2.weight_v", "conv2.weight_g", "conv2.bias") + \
            (("downsample.weight", "downsample.bias") if cin != cout else ()):
        errs["d_" + k] = _l2(params[k].grad, sdo["b." + k].grad)
    return errs


CASES["tcn_block_dropout"] = (case_tcn_block_dropout, _c())
CASES["tcn_block_dropout_downsample"] = (case_tcn_block_dropout, _c(cin=1024, cout=512, dilation=1, p=0.5, B=3, T=17))
for _k in ("conv1.weight_v", "conv1.weight_g", "conv1.bias", "conv2.weight_v", "conv2.weight_g", "conv2.bias",
           "downsample.weight", "downsample.bias"):
    TOLS["d_" + _k] = 3e-2
TOLS["zero_fraction_diff"] = 2e-3


# ----------------------------------------------------------------------------------------------------------
# torch.library custom ops (m3t_b200/custom_ops.py): the dispatcher path gives the autograd.Function path's results
# ----------------------------------------------------------------------------------------------------------
def case_torch_library(seed=0):
    import m3t_b200.custom_ops  # noqa: F401
    from m3t_b200 import ops, raw
    g = torch.Generator().manual_seed(seed)
    errs = {}
    x = _rnd((6, 10, 512), g).float().cuda()
    w, b = _rnd((264, 512), g, 0.05).float().cuda(), _rnd((264,), g, 0.1).float().cuda()
    cot = _rnd((6, 10, 264), g).float().cuda()

    def run(fn):
        xs, ws, bs = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        y = fn(xs, ws, bs)
        (y.float() * cot).sum().backward()
        return y.detach().float(), xs.grad, ws.grad, bs.grad

    ref = run(lambda a, ww, bb: ops.LinearFn.apply(ops.as_bf16(a), ww, bb, True, False))
    got = run(lambda a, ww, bb: torch.ops.m3t.linear(ops.as_bf16(a), ww, bb, True, False))
    errs["lib_linear_exact"] = float(sum((r.float() != t.float()).sum() for r, t in zip(ref, got)))
    xa, xv = _rnd((4, 9, 512), g).cuda().bfloat16(), _rnd((4, 9, 512), g).cuda().bfloat16()
    sa, sv = _rnd((4, 9, 1), g).float().cuda(), _rnd((4, 9, 1), g).float().cuda()

    def run_mix(fn):
        leaves = [t.clone().requires_grad_(True) for t in (xa, xv, sa, sv)]
        f = fn(*leaves)
        f.float().square().sum().backward()
        return [f.detach()] + [t.grad for t in leaves]

    ref = run_mix(ops.AttMixFn.apply)
    got = run_mix(torch.ops.m3t.att_mix)
    errs["lib_att_mix_exact"] = float(sum((r.float() != t.float()).sum() for r, t in zip(ref, got)))
    H = 128
    prm = [(_rnd(s, g, 0.05)).float().cuda() for s in ((3 * H, 512), (3 * H, H), (3 * H,), (3 * H,)) * 2]
    xg = _rnd((3, 7, 512), g).cuda().bfloat16()

    def run_gru(fn):
        leaves = [t.clone().requires_grad_(True) for t in [xg] + prm]
        out = fn(*leaves)
        out.float().square().sum().backward()
        return [out.detach()] + [t.grad for t in leaves]

    ref = run_gru(lambda *a: ops.GRULayerFn.apply(*a, True))
    got = run_gru(torch.ops.m3t.gru_layer)
    errs["lib_gru_out_exact"] = float((ref[0] != got[0]).sum())
    errs["lib_gru_grads"] = max(_l2(t, r) for r, t in zip(ref[1:], got[1:]))
    A, Bm = _rnd((200, 96), g).cuda().bfloat16(), _rnd((72, 96), g).cuda().bfloat16()
    errs["lib_gemm_exact"] = float((torch.ops.m3t.gemm(A, Bm, False, False, True, None, None, None, False) !=
                                    raw.gemm(A, Bm, out_dtype=torch.float32)).sum())
    # the whole AV model through the dispatcher path (M3T_TORCH_OPS=1) equals the default path bit for bit
    import bench as BN
    from m3t_b200.models.model import AffWild2VA
    torch.manual_seed(3)
    m = AffWild2VA(BN.hparams()).cuda().eval()
    BN.randomise_bn(m, 5)
    bt = {k: v.cuda() for k, v in BN.synth_batch(2, 11, pin=False).items()}
    with torch.no_grad():
        y0 = m(bt).clone()
        os.environ["M3T_TORCH_OPS"] = "1"
        try:
            y1 = m(bt).clone()
        finally:
            os.environ.pop("M3T_TORCH_OPS", None)
    errs["lib_model_exact"] = float((y0 != y1).sum())
    try:
        torch.library.opcheck(torch.ops.m3t.att_mix, (xa, xv, sa, sv), test_utils=("test_schema", "test_faketensor"))
        errs["lib_opcheck"] = 0.0
    except Exception as e:  # noqa: BLE001
        errs["lib_opcheck"] = 1.0
        errs["info"] = {"opcheck": str(e)[:300]}
    return errs


CASES["torch_library_ops"] = (case_torch_library, _c())
for _k in ("lib_linear_exact", "lib_att_mix_exact", "lib_gru_out_exact", "lib_gemm_exact", "lib_model_exact", "lib_opcheck"):
    TOLS[_k] = 0.5
TOLS["lib_gru_grads"] = 1e-5


# ----------------------------------------------------------------------------------------------------------
# One-launch re-pack of every convolution filter (ops.prepack / m3t_pack_filters_batched) == the per-filter packs
# ----------------------------------------------------------------------------------------------------------
def case_prepack(seed=0):
    from m3t_b200 import ops, raw
    from m3t_b200.models.resnet import BasicBlock, ResNet
    from m3t_b200.models.rnn import GRU
    torch.manual_seed(seed)
    trunk = ResNet(BasicBlock, [2, 2, 2, 2], 512, zero_init_residual=False, agg_mode="ap", fmap_out_size=3)
    head = GRU(512, 256, 2, 9, 2)                 # 2 BiGRU layers + a 2-layer FC head (Linear weights)
    m = torch.nn.ModuleDict({"trunk": trunk, "head": head}).cuda().train()
    x = torch.randn(32, 64, 28, 28, device="cuda")

    def loss():
        f = m["trunk"](x)
        return m["head"](f.view(4, 8, 512)).square().mean()

    tags = ("filter", "dgrad_s2", "gru_w", "bf16")
    ops.clear_caches()
    ops._prepack_wish.clear()
    ops._prepack_gru.clear()
    ops._prepack_lin.clear()
    loss().backward()                             # records which packs the model asks for
    ref = {}
    for k, v in ops._pack_cache.items():
        if k[1] in tags:
            val = v[2] if isinstance(v[2], tuple) else (v[2],)
            ref[k] = [None if t is None else t.clone() for t in val]
    ops.clear_caches()
    n = ops.prepack({id(p) for p in m.parameters()})
    bad = 0.0
    seen = 0
    for k, want in ref.items():
        got = ops._pack_cache.get(k)
        if got is None:
            bad += 1
            continue
        gval = got[2] if isinstance(got[2], tuple) else (got[2],)
        for a, b in zip(gval, want):
            if (a is None) != (b is None):
                bad += 1
            elif a is not None:
                seen += 1
                if a.shape != b.shape:
                    bad += 1
                elif a.dtype == torch.bfloat16:
                    bad += float((a.reshape(-1).view(torch.int16) != b.reshape(-1).view(torch.int16)).sum())
                else:
                    bad += float((a != b).sum())
    # 19 conv filters + 2 GRU layers x 8 entries + 2 Linear weights
    errs = {"prepack_exact": bad, "prepack_missing": 0.0 if (n == 19 + 16 + 2 and seen >= 53 + 8 + 2) else 1.0}
    errs["info"] = {"entries": n, "tensors_compared": seen, "by_tag": {t: sum(1 for k in ref if k[1] == t) for t in tags}}
    # and the model gives the same gradients with the pre-packed cache (cache hits) as with per-filter packs; the
    # yardstick is the run-to-run distance of two per-filter runs (train-mode BN statistics use fp32 atomics)
    def grads(prepacked):
        ops.clear_caches()
        if prepacked:
            ops.prepack({id(p) for p in m.parameters()})
        m.zero_grad(set_to_none=True)
        loss().backward()
        return [p.grad.clone() for p in m.parameters() if p.grad is not None]

    g0, g0b, g1 = grads(False), grads(False), grads(True)
    noise = max(_l2(a, b) for a, b in zip(g0b, g0))
    errs["prepack_grad_l2"] = max(0.0, max(_l2(a, b) for a, b in zip(g1, g0)) - 3.0 * noise)
    errs["info"]["run_to_run_grad_l2"] = noise
    return errs


CASES["prepack_filters"] = (case_prepack, _c())
TOLS["prepack_exact"] = 0.5
TOLS["prepack_missing"] = 0.5
TOLS["prepack_grad_l2"] = 1e-4


# ----------------------------------------------------------------------------------------------------------
# Fused training loss (m3t_av_loss: both CCC terms + masked CE + dL/dy_hat in one launch) vs the reference-shaped
# PyTorch composition (models/model.py compute_loss) and vs the oracle's training_loss on the CPU
# ----------------------------------------------------------------------------------------------------------
def case_av_loss(seed=0):
    import argparse
    from m3t_b200.models.model import AffWild2VA
    from oracle import ref_torch as R
    g = torch.Generator().manual_seed(seed)
    errs = {}
    for tag, loss_name, C, B, T in (("mtl", "ccc_mtl", 9, 32, 16), ("va", "ccc", 2, 5, 7), ("big", "ccc_mtl", 9, 256, 16)):
        y = torch.randn((B, T, C), generator=g)
        batch = {"label_valence": torch.rand((B, T), generator=g) * 2 - 1,
                 "label_arousal": torch.rand((B, T), generator=g) * 2 - 1,
                 "class_expr": torch.randint(0, 7, (B, T), generator=g),
                 "expr_valid": torch.rand((B, T), generator=g) > 0.3}
        hp = argparse.Namespace(loss=loss_name, loss_lambda=0.3)
        stub = argparse.Namespace(hparams=hp)
        for k in ("ccc_loss", "ce_loss", "mse_loss"):
            setattr(stub, k, getattr(AffWild2VA, k).__get__(stub))
        cb = {k: v.cuda() for k, v in batch.items()}

        def run(fused):
            os.environ["M3T_FUSED_LOSS"] = "1" if fused else "0"
            yc = y.cuda().requires_grad_(True)
            loss, logs = AffWild2VA.compute_loss(stub, yc, cb, sync_free=True)
            loss.backward()
            return float(loss), yc.grad.cpu(), {k: float(v) for k, v in logs.items()}

        try:
            lf, gf, logs_f = run(True)
            lt, gt, logs_t = run(False)
        finally:
            os.environ.pop("M3T_FUSED_LOSS", None)
        yo = y.clone().requires_grad_(True)
        lo = R.training_loss(yo, batch, loss_name, 0.3)
        lo.backward()
        errs["loss_%s_vs_torch" % tag] = abs(lf - lt) / max(abs(lt), 1e-6)
        errs["loss_%s_vs_oracle" % tag] = abs(lf - float(lo)) / max(abs(float(lo)), 1e-6)
        errs["dloss_%s_vs_torch" % tag] = _l2(gf, gt)
        errs["dloss_%s_vs_oracle" % tag] = _l2(gf, yo.grad)
        errs["logs_%s" % tag] = max(abs(logs_f[k] - logs_t[k]) for k in logs_f)
    return errs


CASES["av_loss_fused"] = (case_av_loss, _c())
for _t in ("mtl", "va", "big"):
    for _k in ("loss_%s_vs_torch", "loss_%s_vs_oracle", "logs_%s"):
        TOLS[_k % _t] = 2e-5
    TOLS["dloss_%s_vs_torch" % _t] = 2e-5
    TOLS["dloss_%s_vs_oracle" % _t] = 2e-5


# ----------------------------------------------------------------------------------------------------------
# Deterministic training (raw.set_deterministic / M3T_DETERMINISTIC=1; reference train.py:17): two runs from the same
# state are BIT-identical (losses, every gradient, every parameter after 3 optimizer steps), for the ResNet-backbone
# AV model and for the VGG-M split model; and the slotted accumulation computes the same gradients as the default path.
# ----------------------------------------------------------------------------------------------------------
def case_deterministic(seed=0, steps=3, clips=6):
    import bench as BN
    from m3t_b200 import ops, raw
    from m3t_b200.engine import TrainEngine
    from m3t_b200.models.model import AffWild2VA
    errs, info = {}, {}
    for tag, kw in (("resnet", {}), ("v2psplit", dict(backbone="v2p_split", split_layer=3))):
        hp = BN.hparams()
        for k, v in kw.items():
            setattr(hp, k, v)
        batches = [{k: v.cuda() for k, v in BN.synth_batch(clips, 200 + i, pin=False).items()} for i in range(steps)]

        def run(det):
            prev = raw.set_deterministic(det)
            try:
                ops.clear_caches()
                torch.manual_seed(seed)
                m = AffWild2VA(hp)
                BN.randomise_bn(m, 7)
                m = m.cuda().train()
                eng = TrainEngine(m, lr=1e-4, weight_decay=1e-4, clip=1.0)
                losses, g_first = [], None
                for b in batches:
                    losses.append(eng.step(b).clone())
                    if g_first is None:
                        g_first = eng.flat_g.clone()
                torch.cuda.synchronize()
                return torch.stack(losses), g_first, eng.flat_p.clone()
            finally:
                raw.set_deterministic(prev)

        l1, g1, p1 = run(True)
        l2, g2, p2 = run(True)
        l0, g0, p0 = run(False)
        errs["det_loss_bits_" + tag] = float((l1 != l2).sum())
        errs["det_grad_bits_" + tag] = float((g1 != g2).sum())
        errs["det_param_bits_" + tag] = float((p1 != p2).sum())
        errs["det_vs_default_grad_" + tag] = _l2(g1, g0)
        l0b, g0b, _ = run(False)
        info[tag] = {"default_run_to_run_grad_bits_differ": int((g0 != g0b).sum()),
                     "default_run_to_run_grad_l2": _l2(g0b, g0), "losses": [round(float(x), 5) for x in l1]}
    errs["info"] = info
    return errs


CASES["deterministic_training"] = (case_deterministic, _c())
for _t in ("resnet", "v2psplit"):
    for _k in ("det_loss_bits_", "det_grad_bits_", "det_param_bits_"):
        TOLS[_k + _t] = 0.5
    TOLS["det_vs_default_grad_" + _t] = 0.3       # same math, other summation order, on a chaotic model (see info)


# ----------------------------------------------------------------------------------------------------------
# Programmatic dependent launch (csrc/common.cuh, m3t_set_pdl): every kernel orders itself behind its predecessor with
# griddepcontrol.wait instead of the stream's full serialisation.  A kernel that read its input BEFORE the wait would
# see stale data; in deterministic mode the step is bit-reproducible, so PDL on vs off must agree to the last bit
# (eager, 3 optimizer steps, both model families), and so must eval inference replayed from a graph captured with it.
# ----------------------------------------------------------------------------------------------------------
def case_pdl_exact(seed=0, steps=3, clips=6):
    import bench as BN
    from m3t_b200 import lib, ops, raw
    from m3t_b200.engine import TrainEngine
    from m3t_b200.graphs import GraphedInference
    from m3t_b200.models.model import AffWild2VA
    errs = {}
    for tag, kw in (("resnet", {}), ("v2psplit", dict(backbone="v2p_split", split_layer=3))):
        hp = BN.hparams()
        for k, v in kw.items():
            setattr(hp, k, v)
        batches = [{k: v.cuda() for k, v in BN.synth_batch(clips, 300 + i, pin=False).items()} for i in range(steps)]

        def run(pdl):
            prev_d, prev_p = raw.set_deterministic(True), lib.set_pdl(pdl)
            try:
                ops.clear_caches()
                torch.manual_seed(seed)
                m = AffWild2VA(hp)
                BN.randomise_bn(m, 7)
                m = m.cuda().train()
                eng = TrainEngine(m, lr=1e-4, weight_decay=1e-4, clip=1.0)
                losses = [eng.step(b).clone() for b in batches]
                torch.cuda.synchronize()
                return torch.stack(losses), eng.flat_g.clone(), eng.flat_p.clone()
            finally:
                raw.set_deterministic(prev_d)
                lib.set_pdl(prev_p)

        l0, g0, p0 = run(False)
        l1, g1, p1 = run(True)
        errs["pdl_loss_bits_" + tag] = float((l0 != l1).sum())
        errs["pdl_grad_bits_" + tag] = float((g0 != g1).sum())
        errs["pdl_param_bits_" + tag] = float((p0 != p1).sum())
    # eval inference: graph captured with PDL edges vs plain eager
    from m3t_b200.models.backbone import VA_3DResNet
    torch.manual_seed(seed)
    net = VA_3DResNet(resnet_ver='v1').cuda().eval()
    x = torch.randn(2, 3, 16, 112, 112, device="cuda")
    with torch.no_grad():
        ref = net(x).clone()
    prev = lib.set_pdl(True)
    try:
        gi = GraphedInference(net, x)
        out = gi(x).clone()
        out2 = gi(x).clone()
    finally:
        lib.set_pdl(prev)
    errs["pdl_graph_bits"] = float((out != ref).sum() + (out2 != ref).sum())
    return errs


CASES["pdl_exact"] = (case_pdl_exact, _c())
for _t in ("resnet", "v2psplit"):
    for _k in ("pdl_loss_bits_", "pdl_grad_bits_", "pdl_param_bits_"):
        TOLS[_k + _t] = 0.5
TOLS["pdl_graph_bits"] = 0.5


# ----------------------------------------------------------------------------------------------------------
# A captured training step of a TCN-backend model with dropout (reference models/tcn.py:23,29; `--backend tcn`): the
# launch arguments of a CUDA graph are frozen, so the fused-dropout seed is host seed + a device-resident counter the
# captured step advances (raw.dropout_counter, m3t_tcn_conv_bf16_dseed).  With lr = 0 and one fixed batch the loss
# changes from replay to replay only through the masks: every replay must draw a new one; with p = 0 the same replays
# give one loss (BatchNorm running statistics do not enter a training-mode forward).
# ----------------------------------------------------------------------------------------------------------
def case_tcn_graph_dropout(seed=0, clips=4, replays=4):
    import bench as BN
    from m3t_b200 import ops, raw
    from m3t_b200.engine import TrainEngine
    from m3t_b200.models.model import AffWild2VA
    hp = BN.hparams()
    hp.modality, hp.backbone, hp.backend, hp.loss = "visual", "v2p", "tcn", "ccc"
    batch = {k: v.cuda() for k, v in BN.synth_batch(clips, 77, pin=False).items()}
    errs = {}

    def losses(p_drop):
        ops.clear_caches()
        torch.manual_seed(seed)
        m = AffWild2VA(hp).cuda()
        m.train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = p_drop
        eng = TrainEngine(m, lr=0.0, weight_decay=0.0, clip=1.0)
        eng.capture(batch, warmup=2)
        c0 = int(raw.dropout_counter("cuda").item())
        out = [float(eng.step(batch)) for _ in range(replays)]
        c1 = int(raw.dropout_counter("cuda").item())
        return out, (c1 - c0) % (1 << 64)

    prev = raw.set_deterministic(True)      # fixed-order accumulation: without dropout the replays are bit-identical,
    try:                                    # so distinct losses can only come from distinct masks
        with_drop, adv = losses(0.2)
        without, _ = losses(0.0)
    finally:
        raw.set_deterministic(prev)
    errs["tcn_graph_nodrop_differs"] = float(len(set(without)) != 1)       # control: lr = 0, no dropout -> one loss
    errs["tcn_graph_counter_advances"] = 0.0 if adv == (replays * raw._DROPOUT_STRIDE) % (1 << 64) else 1.0
    errs["tcn_graph_masks_repeat"] = float(len(set(with_drop)) != len(with_drop))
    errs["tcn_graph_nonfinite"] = float(sum(1 for v in with_drop if not (v == v and abs(v) < 1e6)))
    errs["info"] = {"losses_with_dropout": [round(v, 5) for v in with_drop], "losses_without": [round(v, 5) for v in without]}
    return errs


CASES["tcn_graph_dropout"] = (case_tcn_graph_dropout, _c())
for _k in ("tcn_graph_counter_advances", "tcn_graph_masks_repeat", "tcn_graph_nonfinite", "tcn_graph_nodrop_differs"):
    TOLS[_k] = 0.5


if __name__ == "__main__":
    name = sys.argv[1]
    errs = run_case(name)
    ok = not failures(name, errs)
    print("CASE_RESULT " + json.dumps({"case": name, "ok": ok, "errs": errs}))
    sys.exit(0 if ok else 1)
