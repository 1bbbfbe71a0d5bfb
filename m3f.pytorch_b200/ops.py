"""Autograd-level ops of the B200 path.  Each `torch.autograd.Function` below is a fused unit of the reference's
module graph (conv + BatchNorm + residual + ReLU, stem, GRU layer, Linear, attention mix) whose forward and backward
are sequences of C-ABI calls (raw.py -> libm3t_b200.so).  Activations between units are channels-last bf16 tensors
("CL"); parameters stay fp32 in the reference's layouts (state_dict contract) and are re-packed to bf16 on the fly.

Nothing here falls back to stock PyTorch kernels for the ops the library implements; torch is used for memory
(torch.empty / zeros on the current stream), autograd bookkeeping and tiny parameter-side glue.
"""
import os
import weakref

import torch

from . import raw

BN_EPS = 1e-5
BN_MOMENTUM = 0.1

# When a TrainEngine drives the step it bumps every BatchNorm's num_batches_tracked with ONE multi-tensor op per step
# instead of one tiny kernel per layer.
DEFER_NUM_BATCHES_TRACKED = False


def bump_num_batches_tracked(bn):
    if not DEFER_NUM_BATCHES_TRACKED and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)

_pack_cache = {}


def _cached(key_tensor, tag, fn):
    """Cache derived (packed / bf16) copies of a parameter until it is modified in place (optimizer step).
    Entries are keyed by object identity and validated through a weak reference (an address can be recycled by a
    different tensor), the in-place version counter and the storage pointer (`p.data = ...` re-seats storage)."""
    return _cached_multi((key_tensor,), tag, fn)


CACHE_GENERATION = 0      # bumped by clear_caches(): consumers holding raw addresses (graphs.GraphedInference) watch it


def clear_caches():
    global CACHE_GENERATION
    CACHE_GENERATION += 1
    _pack_cache.clear()


# ------------------------------------------------------------------------------------------------------------
# All convolution filters of the model re-packed by ONE launch per step (m3t_pack_filters_batched).  The first step
# records which parameters ask for a pack (and which stride-2 convolutions ask for parity sub-filters); from the
# second step on TrainEngine calls prepack() before the forward, which fills the derived-weight cache in one launch,
# so packed_filter() / the parity lookup are cache hits (was: 19 pack launches + 15 index_select per step).
# ------------------------------------------------------------------------------------------------------------
_prepack_wish = {}      # id(w) -> [weakref(w), parity index lists (4 x list | None) or None]
_prepack_plans = {}


def _wish(w, parity=None):
    e = _prepack_wish.get(id(w))
    if e is None or e[0]() is not w:
        e = _prepack_wish[id(w)] = [weakref.ref(w), None]
    if parity is not None:
        e[1] = parity


_prepack_gru = {}       # tuple(id of the 8 tensors) -> [weakrefs..., (I, H)]
_prepack_lin = {}       # id(w) -> weakref(w)


def _wish_gru(tensors, I, H):
    _prepack_gru[tuple(id(t) for t in tensors)] = ([weakref.ref(t) for t in tensors], (I, H))


def _wish_linear(w):
    _prepack_lin[id(w)] = weakref.ref(w)


def prepack(allowed=None):
    """Fill the derived-weight cache with ONE launch (m3t_pack_filters_batched): every recorded convolution filter
    (fprop / flipped-dgrad packs + stride-2 parity sub-filters), every BiGRU layer (W_ih of both directions stacked,
    W_hh, W_hh^T, the stacked bias vectors) and every Linear weight (bf16 copy).  Returns the number of table entries.
    allowed: optional set of id(parameter) restricting the pack to one model's parameters."""
    import ctypes
    ok = lambda t: t is not None and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and \
        (allowed is None or id(t) in allowed)                                                      # noqa: E731
    convs = [(e[0](), e[1]) for e in _prepack_wish.values()]
    convs = [(w, par) for w, par in convs if ok(w) and w.dim() >= 3 and w.numel() // (w.shape[0] * w.shape[1]) <= 27]
    grus = []
    for refs, (I, H) in _prepack_gru.values():
        ts = [r() for r in refs]
        if all(ok(t) for t in ts) and I % 8 == 0:
            grus.append((ts, I, H))
    lins = [r() for r in _prepack_lin.values()]
    lins = [w for w in lins if ok(w) and w.dim() == 2 and w.shape[1] % 8 == 0]
    n_entries = len(convs) + 8 * len(grus) + len(lins)
    if n_entries == 0 or n_entries > 128:
        return 0
    key = (tuple((id(w), w.data_ptr(), tuple(w.shape), None if par is None else tuple(
        None if p is None else tuple(p) for p in par)) for w, par in convs),
        tuple(tuple((id(t), t.data_ptr()) for t in ts) for ts, _, _ in grus),
        tuple((id(w), w.data_ptr(), tuple(w.shape)) for w in lins))
    plan = _prepack_plans.get("plan")
    if plan is None or plan["key"] != key:
        class Entry(ctypes.Structure):
            _fields_ = [("src", ctypes.c_void_p), ("wf", ctypes.c_void_p), ("wd", ctypes.c_void_p),
                        ("par", ctypes.c_void_p * 4), ("start", ctypes.c_longlong), ("Cout", ctypes.c_int),
                        ("Cin", ctypes.c_int), ("taps", ctypes.c_int), ("has_parity", ctypes.c_int),
                        ("ntaps_par", ctypes.c_int * 4), ("par_of_tap", ctypes.c_int * 27),
                        ("pos_of_tap", ctypes.c_int * 27)]

        dev = (convs[0][0] if convs else (grus[0][0][0] if grus else lins[0])).device
        need = 64
        for w, par in convs:
            need += 3 * w.numel() + 64
        for ts, I, H in grus:
            need += 6 * H * I + 12 * H * H + 64
        for w in lins:
            need += w.numel() + 16
        buf = torch.empty(need, device=dev, dtype=torch.bfloat16)
        fbuf = torch.empty(sum(12 * H for _, _, H in grus) + 8, device=dev, dtype=torch.float32)
        table = (Entry * n_entries)()
        state = {"off": 0, "foff": 0, "start": 0, "i": 0}

        tiles_of = raw._lib().m3t_pack_entry_tiles
        tiles_of.restype = ctypes.c_longlong

        def take(numel):
            t = buf[state["off"]:state["off"] + numel]
            state["off"] += (numel + 7) // 8 * 8
            return t

        def entry(src, numel, Cout, Cin, taps, wf=None, wd=None, kind=0):
            e = table[state["i"]]
            state["i"] += 1
            e.src = src.data_ptr()
            e.wf = wf.data_ptr() if wf is not None else None
            e.wd = wd.data_ptr() if wd is not None else None
            e.start, e.Cout, e.Cin, e.taps, e.has_parity = state["start"], Cout, Cin, taps, kind
            for t in range(27):
                e.par_of_tap[t] = -1
                e.pos_of_tap[t] = 0
            state["start"] += int(tiles_of(Cout, Cin, taps, kind))      # entry.start counts tiles (= blocks)
            return e

        conv_views, gru_views, lin_views = [], [], []
        for w, par in convs:
            Cout, Cin = w.shape[0], w.shape[1]
            taps = w.numel() // (Cout * Cin)
            wf = take(w.numel()).view(Cout, taps * Cin)
            wd = take(w.numel()).view(Cin, taps * Cout)
            e = entry(w, w.numel(), Cout, Cin, taps, wf, wd)
            subs = None
            if par is not None:
                e.has_parity = 1
                subs = []
                for p, idx in enumerate(par):
                    if idx is None:
                        subs.append(None)
                        continue
                    sub = take(Cin * len(idx) * Cout).view(Cin, len(idx) * Cout)
                    subs.append(sub)
                    e.par[p] = sub.data_ptr()
                    e.ntaps_par[p] = len(idx)
                    for j, flipped in enumerate(idx):
                        t = taps - 1 - int(flipped)
                        e.par_of_tap[t], e.pos_of_tap[t] = p, j
                subs = tuple(subs)
            conv_views.append((w, wf, wd, subs))
        for ts, I, H in grus:
            w_ih, w_ih_r, w_hh, w_hh_r, b_ih, b_ih_r, b_hh, b_hh_r = ts
            wih = take(6 * H * I).view(6 * H, I)
            whh = take(6 * H * H).view(2, 3 * H, H)
            whht = take(6 * H * H).view(2, H, 3 * H)
            bias = fbuf[state["foff"]:state["foff"] + 12 * H].view(2, 6 * H)
            state["foff"] += 12 * H
            entry(w_ih, 3 * H * I, 3 * H, I, 1, wf=wih[:3 * H])
            entry(w_ih_r, 3 * H * I, 3 * H, I, 1, wf=wih[3 * H:])
            entry(w_hh, 3 * H * H, 3 * H, H, 1, wf=whh[0], wd=whht[0])
            entry(w_hh_r, 3 * H * H, 3 * H, H, 1, wf=whh[1], wd=whht[1])
            for src, dst in ((b_ih, bias[0, :3 * H]), (b_ih_r, bias[0, 3 * H:]), (b_hh, bias[1, :3 * H]),
                             (b_hh_r, bias[1, 3 * H:])):
                entry(src, 3 * H, 3 * H, 1, 1, wf=dst, kind=-1)           # fp32 copy
            gru_views.append((ts, (wih, whh, whht, bias)))
        for w in lins:
            wb = take(w.numel()).view(w.shape[0], w.shape[1])
            entry(w, w.numel(), w.shape[0], w.shape[1], 1, wf=wb)
            lin_views.append((w, wb))
        tab = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).to(dev)
        plan = _prepack_plans["plan"] = dict(key=key, buf=buf, fbuf=fbuf, tab=tab, convs=conv_views, grus=gru_views,
                                             lins=lin_views, total=state["start"], n=n_entries)
    raw.L.check(raw._lib().m3t_pack_filters_batched(raw.L.ptr(plan["tab"]), raw.L.i32(plan["n"]),
                                                    raw.L.i64(plan["total"]), raw.L.stream_ptr()),
                "pack_filters_batched")
    for w, wf, wd, subs in plan["convs"]:
        ver = ((w._version, w.data_ptr()),)
        _pack_cache[((id(w),), "filter")] = (ver, (weakref.ref(w),), (wf, wd))
        if subs is not None:
            _pack_cache[((id(w),), "dgrad_s2")] = (ver, (weakref.ref(w),), subs)
    for ts, val in plan["grus"]:
        # cache key order of _gru_packed: (w_ih, w_ih_r, w_hh, w_hh_r) + (b_ih, b_ih_r, b_hh, b_hh_r)
        _pack_cache[(tuple(id(t) for t in ts), "gru_w")] = (tuple((t._version, t.data_ptr()) for t in ts),
                                                            tuple(weakref.ref(t) for t in ts), val)
    for w, wb in plan["lins"]:
        _pack_cache[((id(w),), "bf16")] = (((w._version, w.data_ptr()),), (weakref.ref(w),), wb)
    return plan["n"]


def packed_filter(w, want_dgrad):
    def make():
        _wish(w)
        return raw.pack_filter(w.detach(), True)
    wf, wd = _cached(w, "filter", make)
    return wf, wd


# ------------------------------------------------------------------------------------------------------------
# boundary conversions (module API <-> internal CL/bf16)
# ------------------------------------------------------------------------------------------------------------
class ToCL(torch.autograd.Function):
    """fp32 [N,C,*S] -> bf16 [N,*S,C]."""

    @staticmethod
    def forward(ctx, x):
        ctx.C = x.shape[1]
        return raw.ncs_to_nsc_bf16(x.contiguous())

    @staticmethod
    def backward(ctx, g):
        return raw.nsc_to_ncs_f32(g.contiguous(), ctx.C)


class ToCLPad(torch.autograd.Function):
    """fp32 [N,C,*S] -> bf16 [N,*S,cpad] with zero channels C..cpad-1 (an image entering a 64-channel conv block)."""

    @staticmethod
    def forward(ctx, x, cpad):
        ctx.C = x.shape[1]
        return raw.ncs_to_nsc_bf16(x.contiguous(), cpad)

    @staticmethod
    def backward(ctx, g):
        return raw.nsc_to_ncs_f32(g.contiguous(), ctx.C), None


class FromCL(torch.autograd.Function):
    """bf16 [N,*S,C] -> fp32 [N,C,*S]."""

    @staticmethod
    def forward(ctx, x):
        return raw.nsc_to_ncs_f32(x.contiguous())

    @staticmethod
    def backward(ctx, g):
        return raw.ncs_to_nsc_bf16(g.contiguous())


class CastBF16(torch.autograd.Function):
    """fp32 [..., C] -> bf16 [..., C] (C % 8 == 0)."""

    @staticmethod
    def forward(ctx, x):
        shp = x.shape
        x2 = x.contiguous().view(-1, shp[-1])
        return raw.cast_bf16(x2, shp[-1]).view(shp)

    @staticmethod
    def backward(ctx, g):
        shp = g.shape
        return raw.cast_f32(g.contiguous().view(-1, shp[-1])).view(shp)


class CastF32(torch.autograd.Function):
    """bf16 [..., C] -> fp32."""

    @staticmethod
    def forward(ctx, x):
        shp = x.shape
        return raw.cast_f32(x.contiguous().view(-1, shp[-1])).view(shp)

    @staticmethod
    def backward(ctx, g):
        shp = g.shape
        return raw.cast_bf16(g.contiguous().view(-1, shp[-1]), shp[-1]).view(shp)


def as_bf16(x):
    return x if x.dtype == torch.bfloat16 else CastBF16.apply(x)


def as_f32(x):
    return x if x.dtype == torch.float32 else CastF32.apply(x)


# ------------------------------------------------------------------------------------------------------------
# conv + BatchNorm (+ residual) (+ ReLU) on CL tensors
# ------------------------------------------------------------------------------------------------------------
def _geom2d(x_shape, Cout, k, stride, pad_lo, pad_hi, dil=1, nd=2):
    if nd == 2:
        N, H, W, Cin = x_shape
        return raw.conv_geom(2, N, 1, H, W, Cin, Cout, (1, k[0], k[1]), (1, stride, stride), (0, pad_lo[0], pad_lo[1]),
                             (0, pad_hi[0], pad_hi[1]), (1, dil, dil))
    N, W, Cin = x_shape
    return raw.conv_geom(1, N, 1, 1, W, Cin, Cout, (1, 1, k[0]), (1, 1, stride), (0, 0, pad_lo[0]), (0, 0, pad_hi[0]),
                         (1, 1, dil))


def _cba_train_fwd(x, w, gamma, beta, running_mean, running_var, residual, stride, pad, relu):
    """Train-mode conv + BN (+ residual) (+ ReLU).  Returns (out, saved tensors, meta)."""
    Cout = w.shape[0]
    k = (w.shape[2], w.shape[3])
    geom = _geom2d(x.shape, Cout, k, stride, (pad, pad), (pad, pad))
    wf, _ = packed_filter(w, True)
    stats = raw.new_stats(Cout, x.device)
    y = raw.conv_fprop(x, wf, geom, stats=stats)
    y = y.view(y.shape[0], y.shape[2], y.shape[3], y.shape[4])
    count = y.numel() // Cout
    fin = raw.bn_finalize(stats, count, gamma.detach(), beta.detach(), BN_EPS, BN_MOMENTUM, running_mean, running_var)
    out = raw.bn_act(y, fin[2], fin[3], res=residual, relu=relu)
    # the ReLU mask of a unit without residual is recomputed from y in backward: `out` is not kept for it
    saved = (x, w, y, out if (relu and residual is not None) else None, fin)
    meta = (geom, stride, pad, relu, count, residual is not None)
    return out, saved, meta


def _cba_bwd(saved, meta, dout, need_dx, dgrad_residual=None, dx_accum=None):
    """Backward of _cba_train_fwd.  dgrad_residual: bf16 tensor added to dX in the dgrad epilogue (stride 1);
    dx_accum: dX tensor the (stride-2) data gradient is accumulated into in place.
    Returns (dx, dw, dgamma, dbeta, dres)."""
    x, w, y, out, fin = saved
    geom, stride, pad, relu, count, has_res = meta
    dout = dout.contiguous()
    mean, invstd, scale = fin[0], fin[1], fin[2]
    need_dz = has_res and relu
    sums, dz = raw.bn_bwd_reduce(dout, out, y, mean, invstd, relu, need_dz, scale=scale, shift=fin[3])
    if need_dz:   # dz = relu'(out) * dout is already materialised (exactly): two tensor reads instead of three
        dy = raw.bn_bwd_apply(dz, None, y, mean, invstd, scale, sums, count, False)
    else:
        dy = raw.bn_bwd_apply(dout, out, y, mean, invstd, scale, sums, count, relu, shift=fin[3])
    dres = (dz if relu else dout) if has_res else None
    dwp = raw.conv_wgrad(x, dy, geom)
    dw = raw.unpack_filter_grad(dwp, tuple(w.shape))
    dx = None
    if need_dx:
        dx = conv2d_dgrad(dy, w, x.shape, stride, pad, residual=dgrad_residual, accum_into=dx_accum)
    return dx, dw, sums[1], sums[0], dres


class ConvBNAct(torch.autograd.Function):
    """out = act( BN(conv2d(x, w)) + residual )   (models/resnet.py:40-54, one conv of a BasicBlock).

    train: batch statistics come out of the conv epilogue (fp32 column sums), a finalize kernel produces
    scale/shift and updates the running stats, one HBM pass applies BN + residual + ReLU.
    eval: BN is folded to scale/shift and applied in the conv epilogue together with residual + ReLU."""

    @staticmethod
    def forward(ctx, x, w, gamma, beta, running_mean, running_var, residual, stride, pad, relu, training):
        if not training:
            Cout = w.shape[0]
            geom = _geom2d(x.shape, Cout, (w.shape[2], w.shape[3]), stride, (pad, pad), (pad, pad))
            wf, _ = packed_filter(w, True)
            ss = raw.bn_fold(gamma.detach(), beta.detach(), running_mean, running_var, None, BN_EPS)
            out = raw.conv_fprop(x, wf, geom, scale=ss[0], shift=ss[1], residual=residual, relu=relu)
            return out.view(out.shape[0], out.shape[2], out.shape[3], out.shape[4])
        out, saved, meta = _cba_train_fwd(x, w, gamma, beta, running_mean, running_var, residual, stride, pad, relu)
        ctx.save_for_backward(*saved)
        ctx.meta = meta
        return out

    @staticmethod
    def backward(ctx, dout):
        dx, dw, dgamma, dbeta, dres = _cba_bwd(ctx.saved_tensors, ctx.meta, dout, ctx.needs_input_grad[0])
        return dx, dw, dgamma, dbeta, None, None, dres, None, None, None, None


class BasicBlockFn(torch.autograd.Function):
    """One train-mode BasicBlock (models/resnet.py:18-63: conv1-bn1-relu, conv2-bn2, (+downsample), add, relu) as a
    single autograd node, so that the two gradients that meet at the block input are summed inside the dgrad
    epilogue instead of by a separate pass: identity shortcut -> dX = dgrad(conv1) + dz (residual operand of the
    dgrad kernel); 1x1/s2 downsample -> its data gradient accumulates in place into conv1's dX."""

    @staticmethod
    def forward(ctx, x, stride, w1, g1, b1, rm1, rv1, w2, g2, b2, rm2, rv2, wd, gd, bd, rmd, rvd):
        h, s1, m1 = _cba_train_fwd(x, w1, g1, b1, rm1, rv1, None, stride, 1, True)
        if wd is not None:
            idt, sd, md = _cba_train_fwd(x, wd, gd, bd, rmd, rvd, None, stride, 0, False)
        else:
            idt, sd, md = x, (), None
        out, s2, m2 = _cba_train_fwd(h, w2, g2, b2, rm2, rv2, idt, 1, 1, True)
        ctx.save_for_backward(*s1, *s2, *sd)
        ctx.metas = (m1, m2, md)
        return out

    @staticmethod
    def backward(ctx, dout):
        t = ctx.saved_tensors
        m1, m2, md = ctx.metas
        s1, s2, sd = t[0:5], t[5:10], t[10:15]
        need_dx = ctx.needs_input_grad[0]
        dh, dw2, dg2, db2, dres = _cba_bwd(s2, m2, dout, True)
        if md is None:
            dx, dw1, dg1, db1, _ = _cba_bwd(s1, m1, dh, need_dx, dgrad_residual=dres)
            if not need_dx:
                dx = None
            return (dx, None, dw1, dg1, db1, None, None, dw2, dg2, db2, None, None, None, None, None, None, None)
        dx, dw1, dg1, db1, _ = _cba_bwd(s1, m1, dh, need_dx)
        dx, dwd, dgd, dbd, _ = _cba_bwd(sd, md, dres, need_dx, dx_accum=dx)
        return (dx, None, dw1, dg1, db1, None, None, dw2, dg2, db2, None, None, dwd, dgd, dbd, None, None)


class BNActFn(torch.autograd.Function):
    """out = [relu](BatchNorm(x)) on a channels-last bf16 tensor, as a stand-alone unit: the pre-activation of
    BasicBlockV2 / ResNetV2's bn5 (models/resnet.py:163-165,214-215,248-249), DenseNet's norm layers.
    train: batch statistics by one streaming pass (sum x, sum x^2 through m3t_bn_bwd_reduce with mean 0 / invstd 1),
    m3t_bn_finalize (running statistics), one m3t_bn_act pass.  backward: m3t_bn_bwd_reduce + m3t_bn_bwd_apply, the
    ReLU mask recomputed from x."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, relu, training, momentum=BN_MOMENTUM):
        C = x.shape[-1]
        x = x.contiguous()
        if not training:
            ss = raw.bn_fold(gamma.detach(), beta.detach(), running_mean, running_var, None, BN_EPS)
            return raw.bn_act(x, ss[0], ss[1], relu=relu)
        zero, one = _unit(C, x.device)[1], _unit(C, x.device)[0]
        stats, _ = raw.bn_bwd_reduce(x, None, x, zero, one, False, False)       # (sum x, sum x*x)
        count = x.numel() // C
        fin = raw.bn_finalize(stats, count, gamma.detach(), beta.detach(), BN_EPS, momentum, running_mean, running_var)
        out = raw.bn_act(x, fin[2], fin[3], relu=relu)
        ctx.save_for_backward(x, fin)
        ctx.cfg = (relu, count)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, fin = ctx.saved_tensors
        relu, count = ctx.cfg
        dout = dout.contiguous()
        sums, _ = raw.bn_bwd_reduce(dout, None, x, fin[0], fin[1], relu, False, scale=fin[2], shift=fin[3])
        dx = raw.bn_bwd_apply(dout, None, x, fin[0], fin[1], fin[2], sums, count, relu, shift=fin[3])
        return dx, sums[1], sums[0], None, None, None, None, None


class ConvPlainFn(torch.autograd.Function):
    """out = conv2d(x, w) (+ residual) without normalisation: conv2 and the bare 1x1 `downsample` of BasicBlockV2
    (models/resnet.py:146-150,166-176), DenseNet's convolutions.  The residual add rides in the conv epilogue."""

    @staticmethod
    def forward(ctx, x, w, residual, stride, pad):
        Cout = w.shape[0]
        geom = _geom2d(x.shape, Cout, (w.shape[2], w.shape[3]), stride, (pad, pad), (pad, pad))
        wf, _ = packed_filter(w, True)
        out = raw.conv_fprop(x, wf, geom, residual=residual)
        ctx.save_for_backward(x, w)
        ctx.cfg = (geom, stride, pad, residual is not None)
        return out.view(out.shape[0], out.shape[2], out.shape[3], out.shape[4])

    @staticmethod
    def backward(ctx, dout):
        x, w = ctx.saved_tensors
        geom, stride, pad, has_res = ctx.cfg
        dout = dout.contiguous()
        dw = raw.unpack_filter_grad(raw.conv_wgrad(x, dout, geom), tuple(w.shape))
        dx = conv2d_dgrad(dout, w, x.shape, stride, pad) if ctx.needs_input_grad[0] else None
        return dx, dw, (dout if has_res else None), None, None


class Conv3dPlainFn(torch.autograd.Function):
    """out = conv3d(x, w, padding=pad) without normalisation on CL bf16 (N,D,H,W,Cin), stride 1: the 3x3x3 growth
    convolution of a DenseNet layer (models/densenet.py:13-15).  Cin and Cout must be multiples of 64 (the caller
    zero-pads a narrower filter bank)."""

    @staticmethod
    def forward(ctx, x, w, pad):
        N, D, H, W, Cin = x.shape
        Cout = w.shape[0]
        k = tuple(w.shape[2:])
        geom = raw.conv_geom(3, N, D, H, W, Cin, Cout, k, (1, 1, 1), (pad,) * 3, (pad,) * 3, (1, 1, 1))
        wf, wd = raw.pack_filter(w.detach().contiguous(), True)
        out = raw.conv_fprop(x, wf, geom)
        ctx.save_for_backward(x, w, wd)
        ctx.cfg = (geom, pad, k)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w, wd = ctx.saved_tensors
        geom, pad, k = ctx.cfg
        dout = dout.contiguous()
        dw = raw.unpack_filter_grad(raw.conv_wgrad(x, dout, geom), tuple(w.shape))
        dx = None
        if ctx.needs_input_grad[0]:
            N, D, H, W, Cin = x.shape
            lo = tuple(k[i] - 1 - pad for i in range(3))
            g2 = raw.conv_geom(3, N, dout.shape[1], dout.shape[2], dout.shape[3], w.shape[0], Cin, k, (1, 1, 1), lo, lo,
                               (1, 1, 1))
            dx = raw.conv_fprop(dout, wd, g2, tag="dgrad").view(x.shape)
        return dx, dw, None


class AvgPool2x2Fn(torch.autograd.Function):
    """AvgPool3d((1,2,2), stride (1,2,2)) on CL bf16 (F,H,W,C) (models/densenet.py:38)."""

    @staticmethod
    def forward(ctx, x):
        ctx.shape = tuple(x.shape)
        return raw.avgpool2x2(x.contiguous())

    @staticmethod
    def backward(ctx, dy):
        return raw.avgpool2x2_bwd(dy.contiguous(), ctx.shape)


_PARITY_TAPS = {}
_PARITY_LISTS = {}


def _parity_taps(K, pad, ph, pw, device):
    """Stride-2 dgrad by output parity: dX[2a+ph, 2b+pw] only sees the filter taps kh = ph + pad - 2d (d = p - a),
    i.e. a (nh x nw) stride-1 filter over dY.  Returns (index of those taps in the flipped dgrad pack, nh, nw,
    pad_lo_h, pad_lo_w) or None if the parity sees no tap (its quarter of dX is zero)."""
    key = (K, pad, ph, pw, str(device))
    if key not in _PARITY_TAPS:
        def rng(par):
            lo = -((-(par + pad - K + 1)) // 2)
            hi = (par + pad) // 2
            return lo, hi
        (hlo, hhi), (wlo, whi) = rng(ph), rng(pw)
        if hhi < hlo or whi < wlo:
            _PARITY_TAPS[key] = None
        else:
            taps = K * K
            idx = [taps - 1 - ((ph + pad - 2 * dh) * K + (pw + pad - 2 * dw))
                   for dh in range(hlo, hhi + 1) for dw in range(wlo, whi + 1)]
            _PARITY_LISTS[(K, pad, ph, pw)] = idx
            _PARITY_TAPS[key] = (torch.tensor(idx, device=device, dtype=torch.long), hhi - hlo + 1, whi - wlo + 1,
                                 -hlo, -wlo)
    return _PARITY_TAPS[key]


DGRAD_S2_PARITY = True   # stride-2 data gradients as four parity sub-convolutions (False: zero-inserted dY)


def _dgrad_s2_parity(dy, w, x_shape, pad, accum_into=None):
    N, H, W, Cin = x_shape
    Cout, _, K, _ = w.shape
    _, P, Q, _ = dy.shape
    plans = []
    for ph in range(2):
        for pw in range(2):
            tp = _parity_taps(K, pad, ph, pw, dy.device)
            Ha, Wa = (H - ph + 1) // 2, (W - pw + 1) // 2
            if tp is None or Ha == 0 or Wa == 0:
                plans.append(None)
                continue
            idx, nh, nw, plh, plw = tp
            phh, pwh = Ha - P - plh + nh - 1, Wa - Q - plw + nw - 1
            if phh < 0 or pwh < 0 or plh < 0 or plw < 0:
                return None
            plans.append((ph, pw, idx, nh, nw, plh, plw, phh, pwh))
    _, wd = packed_filter(w, True)

    def make():
        _wish(w, [None if pl is None else _PARITY_LISTS[(K, pad, pl[0], pl[1])] for pl in plans])
        wd3 = wd.view(Cin, K * K, Cout)
        return tuple(None if pl is None else wd3.index_select(1, pl[2]).reshape(Cin, -1).contiguous() for pl in plans)

    subs = _cached(w, "dgrad_s2", make)
    full = all(pl is not None for pl in plans)
    if accum_into is not None:
        dx = accum_into
    else:
        dx = (torch.empty if full else torch.zeros)((N, H, W, Cin), device=dy.device, dtype=torch.bfloat16)
    flat = dx.view(-1)
    for pl, wsub in zip(plans, subs):
        if pl is None:
            continue
        ph, pw, _, nh, nw, plh, plw, phh, pwh = pl
        geom = _geom2d(dy.shape, Cin, (nh, nw), 1, (plh, plw), (phh, pwh))
        raw.conv_fprop_scatter(dy, wsub, geom, flat[(ph * W + pw) * Cin:], H * W, 2 * W, 2,
                               accumulate=accum_into is not None)
    return dx


def conv2d_dgrad(dy, w, x_shape, stride, pad, residual=None, accum_into=None):
    """dX of conv2d as stride-1 implicit-GEMM convolutions of dY with the flipped, transposed filter: one for
    stride 1; one per output parity for stride 2 (fallback: a single one over zero-inserted dY).
    residual (stride 1): added in the epilogue.  accum_into: dX buffer to accumulate into (returned)."""
    N, H, W, Cin = x_shape
    Cout, _, kh, kw = w.shape
    if stride == 2 and DGRAD_S2_PARITY and kh == kw and Cout % 64 == 0 and residual is None:
        dx = _dgrad_s2_parity(dy, w, x_shape, pad, accum_into)
        if dx is not None:
            return dx
    if accum_into is not None:
        assert residual is None
        residual = accum_into
    _, wd = packed_filter(w, True)
    if stride == 1:
        src = dy
    else:
        assert stride == 2
        src = raw.zero_insert2(dy, H, W)
    lo = (kh - 1 - pad, kw - 1 - pad)
    hi = (pad, pad) if stride == 2 else lo
    geom = _geom2d(src.shape, Cin, (kh, kw), 1, lo, hi)
    flops = 2.0 * dy.shape[0] * dy.shape[1] * dy.shape[2] * Cout * Cin * kh * kw   # true dgrad work
    if residual is not None and raw.halo_route(geom):
        # the persistent halo kernel has one CTA per SM and its epilogue is on the critical path: a residual read
        # there costs more (+0.2 ms at 4096x28x28x64) than a separate, bandwidth-bound add (0.17 ms)
        dx = raw.conv_fprop(src, wd, geom, algo_flops=flops, tag="dgrad").view(N, H, W, Cin)
        return raw.add_bf16(dx, residual.view(N, H, W, Cin))
    dx = raw.conv_fprop(src, wd, geom, residual=residual, algo_flops=flops, tag="dgrad")
    return dx.view(N, H, W, Cin)


class RawClips:
    """Decoded uint8 frames + per-clip augmentation parameters as the stem's input: frames uint8
    [B,T,Hs,Ws,3] (HWC as decoded) and params int32 [B,8] = {crop_x, crop_y, flip, cut_y1, cut_y2, cut_x1, cut_x2, 0}
    on the device; `shape` is the (B,3,T,H,W) of the clip tensor the reference's load_video would have produced."""

    def __init__(self, frames_u8, params, H=112, W=112):
        self.frames, self.params, self.H, self.W = frames_u8, params, H, W
        self.shape = (frames_u8.shape[0], 3, frames_u8.shape[1], H, W)
        self.device = frames_u8.device


class Stem3D(torch.autograd.Function):
    """Conv3d(3,64,(5,7,7),s(1,2,2),p(2,3,3)) -> BN3d -> ReLU -> MaxPool3d((1,3,3),s(1,2,2),p(0,1,1))
    (models/backbone.py:327-332) on raw video.  The stride-2 7x7 spatial filter is evaluated as a stride-1 4x4 filter
    over the 2x2 space-to-depth image (12 -> 16 channels), which makes it an ordinary im2col-TMA implicit GEMM."""

    @staticmethod
    def forward(ctx, video, w, gamma, beta, running_mean, running_var, normalise, training):
        B, _, T, H, W = video.shape
        if isinstance(video, RawClips):   # crop / mirror / cutout / normalise fused into the layout pass
            xs = raw.video_augment_prep_s2d_w4(video.frames, video.params, H, W, normalise)
        else:
            xs = raw.video_prep_s2d_w4(video.contiguous(), normalise)
        idx = stem_s2d_index(w.device)
        wp = _cached(w, "stem", lambda: raw.gather_pack(w.detach().view(64, -1), idx, 80 * 16))
        # (5,4,1) filter over the 64 = 4 taps x 16 channels of the W-unrolled space-to-depth image; the packed
        # filter [64][(kt,jh,jw,ch)] is the same buffer either way
        geom = raw.conv_geom(3, B, T, H // 2, W // 2, 64, 64, (5, 4, 1), (1, 1, 1), (2, 2, 0), (2, 1, 0), (1, 1, 1))
        flops = 2.0 * B * T * (H // 2) * (W // 2) * 64 * 3 * 5 * 7 * 7     # unpadded 5x7x7x3 filter
        ctx.flops = flops
        halo = raw.USE_HALO_STEM and (H // 2) % 8 == 0 and (W // 2) == 56
        if training:
            stats = raw.new_stats(64, video.device)
            if halo:
                y = raw.stem_fprop_halo(xs, wp, stats=stats, algo_flops=flops)
            else:
                y = raw.conv_fprop(xs, wp, geom, stats=stats, algo_flops=flops, tag="stem")
            y = y.view(B * T, H // 2, W // 2, 64)
            count = y.numel() // 64
            fin = raw.bn_finalize(stats, count, gamma.detach(), beta.detach(), BN_EPS, BN_MOMENTUM, running_mean,
                                  running_var)
            scale, shift = fin[2], fin[3]
        else:
            if halo:
                y = raw.stem_fprop_halo(xs, wp, algo_flops=flops)
            else:
                y = raw.conv_fprop(xs, wp, geom, algo_flops=flops, tag="stem").view(B * T, H // 2, W // 2, 64)
            ss = raw.bn_fold(gamma.detach(), beta.detach(), running_mean, running_var, None, BN_EPS)
            scale, shift = ss[0], ss[1]
        if training:
            out, pidx, ymax = raw.bn_relu_maxpool(y, scale, shift, True, want_ymax=True)
            ctx.save_for_backward(xs, w, y, pidx, fin, ymax)
            ctx.geom, ctx.count = geom, count
        else:
            out, pidx = raw.bn_relu_maxpool(y, scale, shift, False)
        return out

    @staticmethod
    def backward(ctx, dout):
        xs, w, y, pidx, fin, ymax = ctx.saved_tensors
        dy, sums = raw.maxpool_bn_bwd(dout.contiguous(), pidx, y, fin[0], fin[1], fin[2], fin[3], ctx.count, ymax=ymax)
        if raw.USE_HALO_WGRAD:
            dwp = raw.wgrad_stem_halo(xs, dy, algo_flops=ctx.flops)
        else:
            dwp = raw.conv_wgrad(xs, dy, ctx.geom, algo_flops=ctx.flops)
        dw = raw.scatter_unpack(dwp, stem_s2d_index(w.device), (64, 3 * 5 * 7 * 7)).view(w.shape)
        return None, dw, sums[1], sums[0], None, None, None, None


_stem_idx = {}


def stem_s2d_index(device):
    """idx[(kt*4 + jh)*64 + jw*12 + (ph*2+pw)*3 + c] = flat offset of W[c, kt, kh, kw] inside one filter, with
    kh = 2*jh + ph - 1, kw = 2*jw + pw - 1 (entries outside the 7x7 window and channels 48..63 of every tap are -1):
    the channel order m3t_video_prep_s2d_w4 writes."""
    key = str(device)
    if key not in _stem_idx:
        idx = torch.full((5, 4, 64), -1, dtype=torch.int32)
        for kt in range(5):
            for jh in range(4):
                for jw in range(4):
                    for ph in range(2):
                        for pw in range(2):
                            kh, kw = 2 * jh + ph - 1, 2 * jw + pw - 1
                            if 0 <= kh < 7 and 0 <= kw < 7:
                                for c in range(3):
                                    idx[kt, jh, jw * 12 + (ph * 2 + pw) * 3 + c] = ((c * 5 + kt) * 7 + kh) * 7 + kw
        _stem_idx[key] = idx.view(-1).to(device)
    return _stem_idx[key]


class AvgPoolCL(torch.autograd.Function):
    """AdaptiveAvgPool2d(1) + flatten on CL (models/resnet.py:117-119): [F,H,W,C] -> [F,C] bf16."""

    @staticmethod
    def forward(ctx, x):
        ctx.shape = tuple(x.shape)
        return raw.avgpool(x)[0]

    @staticmethod
    def backward(ctx, g):
        return raw.avgpool_bwd(g.contiguous(), ctx.shape)


# ------------------------------------------------------------------------------------------------------------
# Linear, GRU layer, attention mix
# ------------------------------------------------------------------------------------------------------------
def _w_bf16(w):
    def make():
        _wish_linear(w)
        return raw.cast_bf16(w.detach().view(w.shape[0], -1))
    return _cached(w, "bf16", make)


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) on bf16 rows; W, b are fp32 parameters (nn.Linear layout)."""

    @staticmethod
    def forward(ctx, x, w, b, relu, out_f32):
        shp = x.shape
        x2 = x.contiguous().view(-1, shp[-1])
        wb = _w_bf16(w)
        if wb.shape[1] != x2.shape[1]:       # K padded to a multiple of 8 on the weight side only
            raise RuntimeError("LinearFn: in_features must be a multiple of 8")
        N = w.shape[0]
        out = raw.gemm(x2, wb, shift=b.detach() if b is not None else None, relu=relu,
                       out_dtype=torch.float32 if out_f32 else torch.bfloat16,
                       out=torch.empty((x2.shape[0], N), device=x.device,
                                       dtype=torch.float32 if out_f32 else torch.bfloat16))
        ctx.save_for_backward(x2, w, out if relu else None)
        ctx.relu, ctx.has_bias, ctx.shp = relu, b is not None, shp
        return out.view(*shp[:-1], N)

    @staticmethod
    def backward(ctx, dout):
        x2, w, out = ctx.saved_tensors
        N, K = w.shape
        M = x2.shape[0]
        d2 = dout.contiguous().view(M, N)
        npad = (N + 7) // 8 * 8
        if d2.dtype == torch.float32:
            d2 = raw.cast_bf16(d2, npad)
        elif npad != N:
            d2 = raw.cast_bf16(raw.cast_f32(d2), npad)
        if ctx.relu:
            o2 = out if out.dtype == torch.bfloat16 else raw.cast_bf16(out)
            d2 = raw.relu_bwd(d2, o2)
        wb = _w_bf16(w)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = raw.gemm(d2[:, :N] if npad != N else d2, wb, b_mn=True, out_dtype=torch.bfloat16)
            dx = dx.view(ctx.shp)
        dw = raw.gemm(d2[:, :N] if npad != N else d2, x2, a_mn=True, b_mn=True, out_dtype=torch.float32)
        db = raw.colsum(d2, N) if ctx.has_bias else None
        return dx, dw, db, None, None


def torch_ops_enabled():
    """M3T_TORCH_OPS=1: Linear and the attention mix go through the dispatcher (torch.ops.m3t.*, custom_ops.py) instead
    of the autograd.Functions; same kernels, same results."""
    return os.environ.get("M3T_TORCH_OPS", "0") == "1"


def linear(x, w, b, relu=False, out_f32=False):
    if torch_ops_enabled():
        from . import custom_ops  # noqa: F401  (registers torch.ops.m3t)
        return torch.ops.m3t.linear(as_bf16(x), w, b, relu, out_f32)
    return LinearFn.apply(as_bf16(x), w, b, relu, out_f32)


def att_mix(x_a, x_v, s_a, s_v):
    if torch_ops_enabled():
        from . import custom_ops  # noqa: F401
        return torch.ops.m3t.att_mix(x_a.contiguous(), x_v.contiguous(), s_a.float(), s_v.float())
    return AttMixFn.apply(x_a, x_v, s_a, s_v)


def _gru_packed(w_ih, w_ih_r, w_hh, w_hh_r, I, H, biases):
    """(wih [6H, Ipad], whh [2,3H,H], whht [2,H,3H]) bf16 copies of one bidirectional layer and, given the four bias
    vectors, bias fp32 [2, 6H] = ([b_ih ; b_ih_r], [b_hh ; b_hh_r]): one launch, cached until the parameters change."""
    def make():
        if biases and all(b is not None for b in biases):
            _wish_gru((w_ih, w_ih_r, w_hh, w_hh_r) + tuple(biases), I, H)
        Ipad = (I + 7) // 8 * 8
        dev = w_ih.device
        wih = torch.empty((6 * H, Ipad), device=dev, dtype=torch.bfloat16)
        whh = torch.empty((2, 3 * H, H), device=dev, dtype=torch.bfloat16)
        whht = torch.empty((2, H, 3 * H), device=dev, dtype=torch.bfloat16)
        bias = torch.empty((2, 6 * H), device=dev, dtype=torch.float32)
        raw.gru_pack_weights(w_ih.detach(), w_ih_r.detach(), w_hh.detach(), w_hh_r.detach(), wih, whh, whht, I, Ipad, H,
                             [b.detach() for b in biases], bias)
        return wih, whh, whht, bias
    return _cached_multi((w_ih, w_ih_r, w_hh, w_hh_r) + tuple(biases), "gru_w", make)


class GRULayerFn(torch.autograd.Function):
    """One bidirectional nn.GRU layer (models/rnn.py:17,75): x bf16 [B,T,I] -> bf16 [B,T,2H].
    Parameters in nn.GRU layout: w_ih/w_hh/b_ih/b_hh for the forward and the reverse direction."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r, training):
        B, T, I = x.shape
        H = w_hh.shape[1]
        x2 = x.contiguous().view(B * T, I)

        wih, whh, _, bias = _gru_packed(w_ih, w_ih_r, w_hh, w_hh_r, I, H, (b_ih, b_ih_r, b_hh, b_hh_r))
        bih, bhh = bias[0], bias[1]
        gi = raw.gemm(x2, wih, shift=bih, out_dtype=torch.float32)
        out, _, saved = raw.gru_fwd(gi, whh, bhh, B, T, H, training)
        if training:
            ctx.save_for_backward(x2, out, saved, w_ih, w_hh, w_ih_r, w_hh_r, b_ih, b_ih_r, b_hh, b_hh_r)
            ctx.dims = (B, T, I, H)
        return out

    @staticmethod
    def backward(ctx, dout):
        x2, out, saved, w_ih, w_hh, w_ih_r, w_hh_r, b_ih, b_ih_r, b_hh, b_hh_r = ctx.saved_tensors
        B, T, I, H = ctx.dims

        wih, _, whht, _ = _gru_packed(w_ih, w_ih_r, w_hh, w_hh_r, I, H, (b_ih, b_ih_r, b_hh, b_hh_r))   # cache hit
        dgi, dgh, hprev, dbias = raw.gru_bwd(dout.contiguous(), out, saved, whht, B, T, H)
        dx = None
        if ctx.needs_input_grad[0]:
            dxp = raw.gemm(dgi, wih, b_mn=True, out_dtype=torch.bfloat16)              # [B*T, Ipad]
            dx = (dxp if dxp.shape[1] == I else dxp[:, :I].contiguous()).view(B, T, I)
        dwih = raw.gemm(dgi, x2, a_mn=True, b_mn=True, out_dtype=torch.float32)      # [6H, I]
        dbih, dbhh = dbias[0], dbias[1]      # reduced inside the BPTT kernel from the fp32 gate gradients
        dwhh = []
        for d in range(2):
            dwhh.append(raw.gemm(dgh[:, d * 3 * H:(d + 1) * 3 * H], hprev[:, d * H:(d + 1) * H], a_mn=True, b_mn=True,
                                 out_dtype=torch.float32))
        return (dx, dwih[:3 * H], dwhh[0], dbih[:3 * H], dbhh[:3 * H], dwih[3 * H:], dwhh[1], dbih[3 * H:],
                dbhh[3 * H:], None)


def _cached_multi(tensors, tag, fn):
    key = (tuple(id(t) for t in tensors), tag)
    ver = tuple((t._version, t.data_ptr()) for t in tensors)
    hit = _pack_cache.get(key)
    if hit is not None and hit[0] == ver and all(r() is t for r, t in zip(hit[1], tensors)):
        return hit[2]
    val = fn()
    if val is None:
        raise RuntimeError("derived weight cache miss for %s" % tag)
    if len(_pack_cache) > 4096:          # dead entries of discarded models
        for k in [k for k, v in _pack_cache.items() if any(r() is None for r in v[1])]:
            del _pack_cache[k]
    _pack_cache[key] = (ver, tuple(weakref.ref(t) for t in tensors), val)
    return val


class AVLossFn(torch.autograd.Function):
    """ccc / ccc_mtl training loss of the task module and dL/dy_hat in one launch (m3t_av_loss); returns the fp32
    vector (L, L_v, L_a, CE).  Only L is differentiable."""

    @staticmethod
    def forward(ctx, y_hat, label_v, label_a, cls, valid, idx_v, idx_a, n_logits, lam, w_ce):
        y2 = y_hat.contiguous().view(-1, y_hat.shape[-1]).float()
        N, C = y2.shape
        out = torch.empty(4, device=y2.device, dtype=torch.float32)
        dy = torch.empty_like(y2)
        raw.L.check(raw._lib().m3t_av_loss(
            raw.L.ptr(y2), raw.L.ptr(label_v.contiguous().view(-1).float()),
            raw.L.ptr(label_a.contiguous().view(-1).float()),
            raw.L.ptr(cls.contiguous().view(-1).long() if cls is not None else None),
            raw.L.ptr(valid.contiguous().view(-1).to(torch.uint8) if valid is not None else None), raw.L.i32(N),
            raw.L.i32(C), raw.L.i32(idx_v % C), raw.L.i32(idx_a % C), raw.L.i32(n_logits), raw.L.f32(lam),
            raw.L.f32(w_ce), raw.L.ptr(out), raw.L.ptr(dy), raw.L.stream_ptr()), "av_loss")
        ctx.save_for_backward(dy)
        ctx.shape = tuple(y_hat.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        (dy,) = ctx.saved_tensors
        return (dy * dout[0]).view(ctx.shape), None, None, None, None, None, None, None, None, None


class AttMixFn(torch.autograd.Function):
    """f = softmax(sigmoid(s_v), sigmoid(s_a)) . (x_v, x_a)   (models/att_fusion.py:21-25)."""

    @staticmethod
    def forward(ctx, x_a, x_v, s_a, s_v):
        x_a, x_v = x_a.contiguous(), x_v.contiguous()
        s_a, s_v = s_a.contiguous().float(), s_v.contiguous().float()
        ctx.save_for_backward(x_a, x_v, s_a, s_v)
        return raw.att_mix_fwd(x_a, x_v, s_a, s_v)

    @staticmethod
    def backward(ctx, df):
        x_a, x_v, s_a, s_v = ctx.saved_tensors
        dxa, dxv, dsa, dsv = raw.att_mix_bwd(df.contiguous(), x_a, x_v, s_a, s_v)
        return dxa, dxv, dsa, dsv


# ------------------------------------------------------------------------------------------------------------
# TCN pieces: Conv1d (+bias, +ReLU) with asymmetric (causal) padding, residual add + ReLU
# ------------------------------------------------------------------------------------------------------------
class Conv1dBiasAct(torch.autograd.Function):
    """y = act(conv1d(x, w, dilation) + b) on CL [B,T,C] bf16 with explicit lower/upper zero padding: the causal
    TemporalBlock conv is pad_lo=(k-1)*d, pad_hi=0, i.e. `Conv1d(padding=(k-1)d)` followed by Chomp1d
    (models/tcn.py:12-13,19-21) without ever computing the chomped tail."""

    @staticmethod
    def forward(ctx, x, w, b, dilation, pad_lo, pad_hi, relu):
        Cout, Cin, k = w.shape
        geom = _geom2d(x.shape, Cout, (k,), 1, (pad_lo,), (pad_hi,), dilation, nd=1)
        wf, _ = raw.pack_filter(w.detach().contiguous(), True)
        y = raw.conv_fprop(x, wf, geom, shift=b.detach() if b is not None else None, relu=relu)
        y = y.view(x.shape[0], -1, Cout)
        ctx.save_for_backward(x, w, y if relu else None)
        ctx.cfg = (geom, dilation, pad_lo, pad_hi, relu, b is not None)
        return y

    @staticmethod
    def backward(ctx, dout):
        x, w, y = ctx.saved_tensors
        geom, dil, pad_lo, pad_hi, relu, has_b = ctx.cfg
        Cout, Cin, k = w.shape
        dz = dout.contiguous()
        if relu:
            dz = raw.relu_bwd(dz, y)
        db = raw.colsum(dz.view(-1, Cout)) if has_b else None
        dw = raw.unpack_filter_grad(raw.conv_wgrad(x, dz, geom), (Cout, Cin, k))
        dx = None
        if ctx.needs_input_grad[0]:
            _, wd = raw.pack_filter(w.detach().contiguous(), True)
            span = dil * (k - 1)
            g2 = _geom2d(dz.shape, Cin, (k,), 1, (span - pad_lo,), (span - pad_hi,), dil, nd=1)
            dx = raw.conv_fprop(dz, wd, g2).view(x.shape)
        return dx, dw, db, None, None, None, None


class TCNConvFn(torch.autograd.Function):
    """One weight-normed dilated causal Conv1d of a TemporalBlock together with everything that follows it inside the
    block (models/tcn.py:19-33,43-46), ONE launch (m3t_tcn_conv_bf16):
        t = dropout_p(relu(conv1d(x, g * v / ||v||, dilation) + b)) ;  y = t                   (residual is None)
                                                                       y = relu(t + residual)  (second conv of the block)
    Weight-norm is folded into the epilogue: the tensor-core operand is the bf16 pack of `weight_v`, the per-channel
    factor g / ||v|| rides with the bias as the epilogue's scale.  The dropout mask is the counter-based generator of
    m3t_dropout_bf16 over the element index of y (seed drawn from torch's CPU generator; inside a captured CUDA graph
    plus a device-resident step counter, so that replays draw new masks), applied in the epilogue.
    Backward: ONE mask pass (m3t_tcn_epilogue_bwd_bf16: y > 0 selects the outer ReLU, t > 0 the inner ReLU AND the kept
    elements, so no mask is stored or re-derived), bias column sums, wgrad, dgrad; the weight-norm chain rule
    (dg = <dw, v> / ||v||, dv = (g / ||v||) (dw - <dw, v> v / ||v||^2)) runs on the parameter-sized tensors."""

    @staticmethod
    def forward(ctx, x, v, g, b, residual, dilation, pad_lo, drop_p, training):
        Cout, Cin, k = v.shape
        geom = _geom2d(x.shape, Cout, (k,), 1, (pad_lo,), (0,), dilation, nd=1)
        norm = v.detach().flatten(1).norm(dim=1)
        scale = (g.detach().flatten() / norm).contiguous()
        fold = os.environ.get("M3T_TCN_FOLD_WN", "1") == "1"
        if fold:
            wf, _ = packed_filter(v, True)
        else:       # diagnostic: round the EFFECTIVE weight to bf16 (what the bf16-emulating oracle does)
            wf, _ = raw.pack_filter((v.detach() * scale.view(-1, 1, 1)).contiguous(), False)
        p = float(drop_p) if training else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item()) if p > 0 else 0
        # under stream capture the host seed is frozen into the graph: the device-resident step counter, advanced by
        # the captured step itself, makes every replay draw a new mask (raw.dropout_counter)
        seed_dev = raw.dropout_counter(x.device) if (p > 0 and torch.cuda.is_current_stream_capturing()) else None
        y, t = raw.tcn_conv(x, wf, geom, scale if fold else None, b.detach() if b is not None else None,
                            residual=residual, drop_p=p, seed=seed, want_t=training, seed_dev=seed_dev)
        ctx.seed = seed
        ctx.save_for_backward(x, v, g, y, t, scale, norm)
        ctx.cfg = (geom, dilation, pad_lo, p, b is not None, residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, v, g, y, t, scale, norm = ctx.saved_tensors
        geom, dil, pad_lo, p, has_b, has_res = ctx.cfg
        Cout, Cin, k = v.shape
        dy = dy.contiguous()
        if has_res:
            dsum, da = raw.tcn_epilogue_bwd(dy, y, t, 1.0 / (1.0 - p))
        else:
            dsum, da = raw.tcn_epilogue_bwd(dy, None, y, 1.0 / (1.0 - p))
        db = raw.colsum(da.view(-1, Cout)) if has_b else None
        dw = raw.unpack_filter_grad(raw.conv_wgrad(x, da, geom), (Cout, Cin, k))      # gradient of the EFFECTIVE weight
        # the operand was bf16(v) and the epilogue multiplied by scale = g / ||v||:  w_eff = scale * v
        dot = (dw * v).flatten(1).sum(dim=1)                    # <dw, v> per output channel
        dg = (dot / norm).view_as(g)
        dv = scale.view(-1, 1, 1) * dw - (scale * dot / (norm * norm)).view(-1, 1, 1) * v
        dx = None
        if ctx.needs_input_grad[0]:
            w_eff = (v.detach() * scale.view(-1, 1, 1)).contiguous()
            _, wd = raw.pack_filter(w_eff, True)
            span = dil * (k - 1)
            g2 = _geom2d(da.shape, Cin, (k,), 1, (span - pad_lo,), (span,), dil, nd=1)
            dx = raw.conv_fprop(da, wd, g2).view(x.shape)
        return dx, dv, dg, db, dsum, None, None, None, None


class Conv2dBiasAct(torch.autograd.Function):
    """y = act(conv2d(x, w, padding) + b) on CL bf16 (F,H,W,Cin) -> (F,H,W,Cout), stride 1: the conv + ReLU units of
    VGGFace (reference models/vggface.py:41-50).  Bias and ReLU ride in the conv epilogue; backward = one mask pass,
    bias column sums, wgrad, dgrad.  w may have fewer input channels than x (x zero-padded to a multiple of 64)."""

    @staticmethod
    def forward(ctx, x, w, b, pad, relu):
        Cout, Cin_w, kh, kw = w.shape
        Cin = x.shape[-1]
        if Cin_w != Cin:        # zero filter planes for the padding channels of x
            wp = torch.zeros((Cout, Cin, kh, kw), device=w.device, dtype=w.dtype)
            wp[:, :Cin_w] = w.detach()
        else:
            wp = w.detach()
        geom = _geom2d(x.shape, Cout, (kh, kw), 1, (pad, pad), (pad, pad))
        wf, wd = raw.pack_filter(wp.contiguous(), True)
        y = raw.conv_fprop(x, wf, geom, shift=b.detach() if b is not None else None, relu=relu)
        y = y.view(y.shape[0], y.shape[2], y.shape[3], y.shape[4])
        ctx.save_for_backward(x, w, y if relu else None, wd)
        ctx.cfg = (geom, pad, relu, b is not None, Cin_w)
        return y

    @staticmethod
    def backward(ctx, dout):
        x, w, y, wd = ctx.saved_tensors
        geom, pad, relu, has_b, Cin_w = ctx.cfg
        Cout, _, kh, kw = w.shape
        Cin = x.shape[-1]
        dz = dout.contiguous()
        if relu:
            dz = raw.relu_bwd(dz, y)
        db = raw.colsum(dz.view(-1, Cout)) if has_b else None
        dw = raw.unpack_filter_grad(raw.conv_wgrad(x, dz, geom), (Cout, Cin, kh, kw))
        if Cin_w != Cin:
            dw = dw[:, :Cin_w].contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            g2 = _geom2d(dz.shape, Cin, (kh, kw), 1, (kh - 1 - pad, kw - 1 - pad), (kh - 1 - pad, kw - 1 - pad))
            dx = raw.conv_fprop(dz, wd, g2, tag="dgrad")
            dx = dx.view(x.shape)
        return dx, dw, db, None, None


class MaxPool2x2Ceil(torch.autograd.Function):
    """F.max_pool2d(x, 2, 2, 0, ceil_mode=True) on a NON-NEGATIVE CL bf16 map (it follows a ReLU in VGGFace,
    models/vggface.py:50): an odd extent is zero-padded by one row / column at the high end, which cannot change a
    maximum of non-negative values, then the 2x2/s2 pooling kernel runs (m3t_bn_relu_maxpool with unit scale)."""

    @staticmethod
    def forward(ctx, x):
        F_, H, W, C = x.shape
        ph, pw = H % 2, W % 2
        if ph or pw:
            x = torch.nn.functional.pad(x, (0, 0, 0, pw, 0, ph))
        one, zero = _unit(C, x.device)
        out, idx = raw.bn_relu_maxpool(x.contiguous(), one, zero, True, (2, 2, 0))
        ctx.save_for_backward(x, idx)
        ctx.crop = (H, W)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, idx = ctx.saved_tensors
        F_, Hp, Wp, C = x.shape
        one, zero = _unit(C, x.device)
        sums = raw.zeros_f32((2, C), x.device)
        fn = raw._lib().m3t_maxpool_bn_bwd
        dx = torch.empty_like(x)
        # mode 1 with zero statistics sums and unit scale: dx = route(dout) * relu'(x); x > 0 where a maximum was taken
        raw.L.check(fn(raw.L.i32(1), raw.L.ptr(dout.contiguous()), raw.L.ptr(idx), raw.L.ptr(x), raw.L.ptr(zero),
                       raw.L.ptr(one), raw.L.ptr(one), raw.L.ptr(zero), raw.L.ptr(sums), raw.ctypes_double(1.0),
                       raw.L.ptr(dx), raw.L.i32(F_), raw.L.i32(Hp), raw.L.i32(Wp), raw.L.i32(C), raw.L.i32(2),
                       raw.L.i32(2), raw.L.i32(0), raw.L.stream_ptr()), "maxpool_bwd")
        H, W = ctx.crop
        return dx[:, :H, :W].contiguous() if (H, W) != (Hp, Wp) else dx


_unit_affine = {}


def _unit(C, device):
    key = (C, str(device))
    if key not in _unit_affine:
        _unit_affine[key] = (torch.ones(C, device=device), torch.zeros(C, device=device))
    return _unit_affine[key]


class AddFn(torch.autograd.Function):
    """a + b on bf16 tensors (BasicBlockV2's residual join when a CBAM sits between conv2 and the add)."""

    @staticmethod
    def forward(ctx, a, b):
        return raw.add_bf16(a.contiguous(), b.contiguous())

    @staticmethod
    def backward(ctx, dout):
        return dout, dout


class AddReLU(torch.autograd.Function):
    """relu(a + b) on bf16 tensors (models/tcn.py:46)."""

    @staticmethod
    def forward(ctx, a, b):
        one, zero = _unit(a.shape[-1], a.device)
        out = raw.bn_act(a.contiguous(), one, zero, res=b.contiguous(), relu=True)
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, dout):
        (out,) = ctx.saved_tensors
        dz = raw.relu_bwd(dout.contiguous(), out)
        return dz, dz


class DropoutFn(torch.autograd.Function):
    """Inverted dropout on a bf16 tensor (nn.Dropout in the temporal blocks, models/tcn.py:23,29): one streaming pass
    forward, the same pass on the gradient backward — the mask is a pure function of (seed, element index), so it is
    re-derived instead of stored.  The seed is drawn from torch's CPU generator, so `torch.manual_seed` makes a run
    reproducible (the mask stream itself is this library's, not PyTorch's Philox)."""

    @staticmethod
    def forward(ctx, x, p):
        ctx.p = float(p)
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("DropoutFn: the stand-alone dropout pass takes a host seed, a captured step would replay "
                               "one mask; the fused TemporalBlock epilogue (M3T_TCN_FUSED=1, the default) does not")
        ctx.seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
        return raw.dropout_bf16(x.contiguous(), ctx.p, ctx.seed)

    @staticmethod
    def backward(ctx, dy):
        return raw.dropout_bf16(dy.contiguous(), ctx.p, ctx.seed), None


def fused_dropout_enabled():
    """Training-mode TCN dropout runs through DropoutFn / m3t_dropout_bf16 (mask bit-exact against oracle/dropout.py
    on a B200, tests/gpu_cases.py::case_dropout).  M3T_FUSED_DROPOUT=0 falls back to torch.nn.functional.dropout."""
    return os.environ.get("M3T_FUSED_DROPOUT", "1") == "1"


# ------------------------------------------------------------------------------------------------------------
# General conv + bias + BatchNorm (+ ReLU) (+ spatial max-pool) unit for the VGG-M 3-D stacks and tcn_simple
# ------------------------------------------------------------------------------------------------------------
_vggm_idx = {}


def vggm_s2d_index(device):
    """First VGG-M conv (Conv3d(3,64,3,stride=(1,2,2),padding=(1,0,0)), models/backbone.py:73) evaluated as a
    (3,2,2) stride-1 filter over the 2x2 space-to-depth image: idx[((kt*2+jh)*2+jw)*16 + (ph*2+pw)*3 + c] = flat
    offset of W[c, kt, kh=2jh+ph, kw=2jw+pw] inside one filter (kh or kw == 3 and the 4 pad channels -> -1)."""
    key = str(device)
    if key not in _vggm_idx:
        idx = torch.full((3, 2, 2, 16), -1, dtype=torch.int32)
        for kt in range(3):
            for jh in range(2):
                for jw in range(2):
                    for ph in range(2):
                        for pw in range(2):
                            kh, kw = 2 * jh + ph, 2 * jw + pw
                            if kh < 3 and kw < 3:
                                for c in range(3):
                                    idx[kt, jh, jw, (ph * 2 + pw) * 3 + c] = ((c * 3 + kt) * 3 + kh) * 3 + kw
        _vggm_idx[key] = idx.view(-1).to(device)
    return _vggm_idx[key]


class ConvNdBNAct(torch.autograd.Function):
    """out = [maxpool_(H,W)] relu( BN( convNd(x, w) + b ) ) on channels-last bf16, nd in {1, 3}, stride 1.
    cfg: dict(nd, k=(d,h,w), pad_lo, pad_hi, relu, pool=None|(K,S,PAD), s2d_first=bool)."""

    @staticmethod
    def forward(ctx, x, w, cbias, gamma, beta, running_mean, running_var, cfg, training):
        nd, k = cfg["nd"], cfg["k"]
        Cout = w.shape[0]
        if nd == 3:
            N, D, H, W, Cin = x.shape
        else:
            N, W, Cin = x.shape
            D = H = 1
        geom = raw.conv_geom(nd, N, D, H, W, Cin, Cout, k, (1, 1, 1), cfg["pad_lo"], cfg["pad_hi"], (1, 1, 1))
        if cfg.get("s2d_first"):
            idx = vggm_s2d_index(w.device)
            wf = _cached(w, "vggm1", lambda: raw.gather_pack(w.detach().view(Cout, -1), idx, 12 * 16))
            flops = 2.0 * N * D * (H - 1) * (W - 1) * Cout * 3 * 27
        else:
            wf, _ = packed_filter(w, True)
            flops = None
        pool = cfg.get("pool")
        relu = cfg["relu"]
        Z, P, Q = raw.conv_out_dims(geom)
        if not training:
            ss = raw.bn_fold(gamma.detach(), beta.detach(), running_mean, running_var,
                             cbias.detach() if cbias is not None else None, BN_EPS)
            if pool:
                y = raw.conv_fprop(x, wf, geom, algo_flops=flops)
                out, _ = raw.bn_relu_maxpool(y.view(N * Z, P, Q, Cout), ss[0], ss[1], False, pool)
                out = out.view(N, Z, out.shape[1], out.shape[2], Cout)
            else:
                out = raw.conv_fprop(x, wf, geom, scale=ss[0], shift=ss[1], relu=relu, algo_flops=flops)
            return out if nd == 3 else out.view(N, Q, Cout)
        if raw.DETERMINISTIC and cfg.get("s2d_first"):
            # the 16-channel im2col mode runs on the one-tile-per-CTA kernel (no slotted statistics): one deterministic
            # streaming pass over the stored conv output instead
            y = raw.conv_fprop(x, wf, geom, algo_flops=flops)
            zero, one = _unit(Cout, x.device)[1], _unit(Cout, x.device)[0]
            stats, _ = raw.bn_bwd_reduce(y.view(-1, Cout), None, y.view(-1, Cout), zero, one, False, False)
        else:
            stats = raw.new_stats(Cout, x.device)
            y = raw.conv_fprop(x, wf, geom, stats=stats, algo_flops=flops)
        count = y.numel() // Cout
        fin = raw.bn_finalize(stats, count, gamma.detach(), beta.detach(), BN_EPS, BN_MOMENTUM, running_mean,
                              running_var)
        if cbias is not None:   # BN(conv + b): the bias only shifts the batch mean that enters running_mean
            running_mean.add_(cbias.detach(), alpha=BN_MOMENTUM)
        pidx = ymax = None
        if pool:
            out, pidx, ymax = raw.bn_relu_maxpool(y.view(N * Z, P, Q, Cout), fin[2], fin[3], True, pool, want_ymax=True)
            out = out.view(N, Z, out.shape[1], out.shape[2], Cout)
        else:
            out = raw.bn_act(y, fin[2], fin[3], relu=relu)
        ctx.save_for_backward(x, w, y, pidx, ymax, fin)      # ReLU mask recomputed from y in backward
        ctx.cfg, ctx.geom, ctx.count, ctx.flops = cfg, geom, count, flops
        ctx.has_bias = cbias is not None
        return out if nd == 3 else out.view(N, Q, Cout)

    @staticmethod
    def backward(ctx, dout):
        x, w, y, pidx, ymax, fin = ctx.saved_tensors
        cfg, geom = ctx.cfg, ctx.geom
        nd, k, pool, relu = cfg["nd"], cfg["k"], cfg.get("pool"), cfg["relu"]
        Cout = w.shape[0]
        N, Z, P, Q = y.shape[0], y.shape[1], y.shape[2], y.shape[3]
        dout = dout.contiguous()
        if pool:
            dy, sums = raw.maxpool_bn_bwd(dout.view(N * Z, -1, dout.shape[-2], Cout) if nd == 3 else dout, pidx,
                                          y.view(N * Z, P, Q, Cout), fin[0], fin[1], fin[2], fin[3], ctx.count, pool,
                                          ymax=ymax)
            dy = dy.view(y.shape)
        else:
            d5 = dout.view(y.shape)
            sums, _ = raw.bn_bwd_reduce(d5, None, y, fin[0], fin[1], relu, False, scale=fin[2], shift=fin[3])
            dy = raw.bn_bwd_apply(d5, None, y, fin[0], fin[1], fin[2], sums, ctx.count, relu, shift=fin[3])
        dwp = raw.conv_wgrad(x, dy, geom, algo_flops=ctx.flops)
        if cfg.get("s2d_first"):
            dw = raw.scatter_unpack(dwp, vggm_s2d_index(w.device), (Cout, w[0].numel())).view(w.shape)
        else:
            dw = raw.unpack_filter_grad(dwp, tuple(w.shape))
        dx = None
        if ctx.needs_input_grad[0]:
            _, wd = packed_filter(w, True)
            lo = tuple(k[i] - 1 - cfg["pad_lo"][i] for i in range(3))
            hi = tuple(k[i] - 1 - cfg["pad_hi"][i] for i in range(3))
            Cin = x.shape[-1]
            g2 = raw.conv_geom(nd, N, Z, P, Q, Cout, Cin, k, (1, 1, 1), lo, hi, (1, 1, 1))
            dx = raw.conv_fprop(dy, wd, g2, tag="dgrad").view(x.shape)
        dcb = torch.zeros_like(fin[0]) if ctx.has_bias else None   # exact: batch-norm removes the bias
        return dx, dw, dcb, sums[1], sums[0], None, None, None, None


# ------------------------------------------------------------------------------------------------------------
# 1-channel 3x3 conv + BN + ReLU (stem of the audio ResNet composition, BASELINE config 2)
# ------------------------------------------------------------------------------------------------------------
class Conv3x3C1BNReLU(torch.autograd.Function):
    """relu(BN(conv2d(x, w, padding=1))) for a single input channel: x fp32 [N,H,W] -> CL bf16 [N,H,W,Cout].
    The 9-tap patches become 16-wide bf16 GEMM rows (m3t_patch3x3_c1) and the conv is one tcgen05 GEMM with K = 16;
    BN statistics come out of its epilogue."""

    @staticmethod
    def forward(ctx, x, w, gamma, beta, running_mean, running_var, training):
        N, H, W = x.shape
        Cout = w.shape[0]
        patches = raw.patch3x3_c1(x)
        wb = _cached(w, "c1", lambda: raw.cast_bf16(w.detach().view(Cout, 9), 16))
        if not training:
            ss = raw.bn_fold(gamma.detach(), beta.detach(), running_mean, running_var, None, BN_EPS)
            out = raw.gemm(patches, wb, scale=ss[0], shift=ss[1], relu=True)
            return out.view(N, H, W, Cout)
        if raw.DETERMINISTIC:     # the GEMM epilogue's statistics have no slotted form: one deterministic streaming pass
            y = raw.gemm(patches, wb)
            zero, one = _unit(Cout, x.device)[1], _unit(Cout, x.device)[0]
            stats, _ = raw.bn_bwd_reduce(y, None, y, zero, one, False, False)
        else:
            stats = raw.new_stats(Cout, x.device)
            y = raw.gemm(patches, wb, stats=stats)
        fin = raw.bn_finalize(stats, y.shape[0], gamma.detach(), beta.detach(), BN_EPS, BN_MOMENTUM, running_mean,
                              running_var)
        out = raw.bn_act(y, fin[2], fin[3], relu=True)
        ctx.save_for_backward(patches, w, y, None, fin)
        return out.view(N, H, W, Cout)

    @staticmethod
    def backward(ctx, dout):
        patches, w, y, out, fin = ctx.saved_tensors
        d2 = dout.contiguous().view(y.shape)
        sums, _ = raw.bn_bwd_reduce(d2, None, y, fin[0], fin[1], True, False, scale=fin[2], shift=fin[3])
        dy = raw.bn_bwd_apply(d2, None, y, fin[0], fin[1], fin[2], sums, y.shape[0], True, shift=fin[3])
        dwp = raw.gemm(dy, patches, a_mn=True, b_mn=True, out_dtype=torch.float32)     # [Cout, 16]
        dw = dwp[:, :9].reshape(w.shape)
        return None, dw, sums[1], sums[0], None, None, None
