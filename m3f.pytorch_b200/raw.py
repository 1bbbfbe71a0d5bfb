"""Raw tensor-level wrappers over the C ABI (no autograd).  Arguments are CUDA torch tensors already in the
layout the ABI wants (channels-last bf16 activations, packed bf16 filters); outputs are allocated here with
torch.empty so they live in PyTorch's caching allocator and on the current stream."""
import ctypes
import os

import torch

from . import lib as L


# Optional per-call CUDA-event timing of the tensor-core kernels (bench.py's roofline leg).  Events are recorded on
# the current stream, i.e. the stream the kernels are launched on.
_prof = None


class KernelProfiler:
    def __init__(self, track_hbm=True, only=None):
        self.records = []          # (key, algorithmic flops, algorithmic HBM bytes, start event, end event)
        self.track_hbm = track_hbm  # False: the ~60 streaming passes per step ("hbm ..." keys) run un-timed
        self.only = None if only is None else frozenset(only)   # a set of keys: every other launch runs un-timed

    def add(self, key, flops, e0, e1, nbytes=0.0):
        self.records.append((key, flops, nbytes, e0, e1))

    def summary(self):
        """key -> dict(calls, ms_total, flops_total, bytes_total, tflops, gbps) (call after torch.cuda.synchronize()).
        flops / bytes are ALGORITHMIC: unpadded 2*M*N*K for the tensor-core kernels; one read + one write
        of each tensor a streaming pass must touch for the HBM-bound ones."""
        out = {}
        for key, flops, nbytes, e0, e1 in self.records:
            d = out.setdefault(key, {"calls": 0, "ms_total": 0.0, "flops_total": 0.0, "bytes_total": 0.0})
            d["calls"] += 1
            d["ms_total"] += e0.elapsed_time(e1)
            d["flops_total"] += flops
            d["bytes_total"] += nbytes
        for d in out.values():
            d["tflops"] = d["flops_total"] / (d["ms_total"] * 1e-3) / 1e12 if d["ms_total"] > 0 else 0.0
            d["gbps"] = d["bytes_total"] / (d["ms_total"] * 1e-3) / 1e9 if d["ms_total"] > 0 else 0.0
        return out


def set_profiler(p):
    global _prof
    _prof = p


def _timed(key, flops, fn, nbytes=0.0):
    if _prof is None or (not _prof.track_hbm and key.startswith("hbm ")) \
            or (_prof.only is not None and key not in _prof.only):
        return fn()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    _prof.add(key, flops, e0, e1, nbytes)
    return r


class StepPool:
    """Zero-initialised fp32 scratch for ONE training step.  A step needs ~90 small zeroed accumulators (BatchNorm
    statistics, split-K weight-gradient tiles, bias-gradient sums): as separate `torch.zeros` calls they were 51+ fill
    kernels per step.  Here they are bump-allocated slices of one buffer that is cleared by ONE
    memset at the start of the next step, over the part that was handed out (the rest has never been written).
    Active only between begin_step() / end_step() (TrainEngine, Trainer.fit); elsewhere zeros_f32 is torch.zeros."""

    def __init__(self, device, nbytes=96 << 20):
        self.buf = torch.zeros(nbytes // 4, device=device, dtype=torch.float32)
        self.used = 0          # floats handed out in the current step
        self.dirty = 0         # high-water mark: everything ever handed out (a captured step graph clears exactly this
        self.active = False    # range on every replay, whatever ran eagerly in between)

    def begin(self):
        if self.dirty:
            self.buf[:self.dirty].zero_()
        self.used = 0
        self.active = True

    def end(self):
        self.active = False

    def take(self, numel):
        n = (numel + 63) // 64 * 64            # 256-byte slots
        if not self.active or self.used + n > self.buf.numel():
            return None
        t = self.buf[self.used:self.used + numel]
        self.used += n
        self.dirty = max(self.dirty, self.used)
        return t


_pools = {}


def begin_step(device):
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    pool = _pools.get(key)
    if pool is None:
        pool = _pools[key] = StepPool(torch.device("cuda", key))
    pool.begin()
    return pool


def end_step(device):
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key in _pools:
        _pools[key].end()


def zeros_f32(shape, device):
    """fp32 zeros: a slice of the active step pool, else torch.zeros."""
    pool = _pools.get(torch.device(device).index)
    if pool is not None and pool.active:
        numel = 1
        for d in shape:
            numel *= int(d)
        t = pool.take(numel)
        if t is not None:
            return t.view(shape)
    return torch.zeros(shape, device=device, dtype=torch.float32)


# ----------------------------------------------------------------------------------------------------------
# Deterministic training (reference train.py:17).  M3T_DETERMINISTIC=1 or set_deterministic(True): every fp32 atomic
# accumulation of the step runs in its slotted form (include/m3t_b200.h "Deterministic training") - destinations are
# the first of (1 + nslots) zero-filled copies, each CTA / warp / split owns one copy, m3t_det_reduce sums them in
# index order - so two runs from the same state give bit-identical losses, gradients and parameters.
# ----------------------------------------------------------------------------------------------------------
DETERMINISTIC = os.environ.get("M3T_DETERMINISTIC", "0") == "1"


def set_deterministic(flag):
    global DETERMINISTIC
    prev, DETERMINISTIC = DETERMINISTIC, bool(flag)
    return prev


def _det_slots(kind):
    lib = _lib()
    return int(lib.m3t_det_stats_slots() if kind == "stats" else lib.m3t_det_cta_slots())


def _det_accum(shape, nslots, device):
    """(1 + nslots) zero-filled copies of an fp32 accumulator; returns (copy 0, the whole buffer)."""
    buf = torch.zeros((1 + nslots,) + tuple(shape), device=device, dtype=torch.float32)
    return buf[0], buf


def _det_reduce(dst, nslots):
    """dst = copy 0 of a _det_accum buffer: dst += copy 1 + copy 2 + ... in order."""
    L.check(_lib().m3t_det_reduce(L.ptr(dst), L.i64(dst.numel()), L.i32(nslots), L.stream_ptr()), "det_reduce")


def new_stats(C, device):
    """fp32 [2, C] accumulator for BatchNorm statistics / BatchNorm-backward sums (zero-initialised).  In deterministic
    mode it is copy 0 of a slotted buffer (the view keeps the buffer alive)."""
    if DETERMINISTIC:
        return _det_accum((2, C), _det_slots("stats"), device)[0]
    return zeros_f32((2, C), device)


def _nb(*ts):
    """Bytes of the given tensors (None skipped): algorithmic traffic of a streaming pass = each operand once."""
    return float(sum(t.numel() * t.element_size() for t in ts if t is not None))


def _chk_bf16(*ts):
    for t in ts:
        if t is not None:
            assert t.is_cuda and t.dtype == torch.bfloat16 and t.stride(-1) == 1, (t.dtype, t.shape, t.stride())


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
    def run():
        rc = L.load().m3t_gemm_bf16(
            L.ptr(A), L.i64(A.stride(0)), L.i32(a_mn), L.ptr(B), L.i64(B.stride(0)), L.i32(b_mn), L.ptr(out),
            L.i64(out.stride(0)), L.i32(out.dtype == torch.float32), L.i32(M), L.i32(N), L.i32(K), L.ptr(scale),
            L.ptr(shift), L.ptr(residual), L.i64(residual.stride(0) if residual is not None else 0), L.i32(relu),
            L.ptr(stats), L.stream_ptr())
        L.check(rc, "m3t_gemm_bf16")

    # GEMM launches are recorded without a FLOP credit: bench.py's roofline / conv aggregate stay convolution-only
    _timed("gemm %dx%dx%d" % (M, N, K), 0.0, run)
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


def conv_key(kind, g):
    return "%s nd%d %dx%dx%d c%d->%d k%dx%dx%d s%d" % (kind, g[0], g[2], g[3], g[4], g[5], g[6], g[7], g[8], g[9],
                                                        g[12])


def halo_route(g, tile_hint=0):
    """True if conv_fprop sends this geometry to the persistent halo-tile kernel (64->64 3x3/s1/p1, W+2 <= 48)."""
    return (USE_HALO and tile_hint == 0 and g[0] == 2 and g[5] == 64 and g[6] == 64 and tuple(g[7:10]) == (1, 3, 3)
            and tuple(g[10:13]) == (1, 1, 1) and tuple(g[13:19]) == (0, 0, 1, 1, 1, 1)
            and tuple(g[19:22]) == (1, 1, 1) and g[4] + 2 <= 48)


USE_HALO128 = True   # route 128->128 3x3/s1/p1 convolutions over 14x14-like images to the image-per-tile halo kernel


def halo128_route(g, tile_hint=0):
    H, W = g[3], g[4]
    return (USE_HALO128 and tile_hint == 0 and g[0] == 2 and g[5] == 128 and g[6] == 128
            and tuple(g[7:10]) == (1, 3, 3) and tuple(g[10:13]) == (1, 1, 1) and tuple(g[13:19]) == (0, 0, 1, 1, 1, 1)
            and tuple(g[19:22]) == (1, 1, 1) and 128 <= H * (W + 2) <= 256 and (H + 2) * (W + 2) <= 256
            and ((H + 2) * (W + 2)) % 8 == 0)


def conv3x3_c128_halo(x, w_packed, scale=None, shift=None, residual=None, relu=False, stats=None, tag="fprop"):
    """x: bf16 [F,H,W,128]; w_packed bf16 [128, 1152] -> bf16 [F,H,W,128]."""
    _chk_bf16(x, w_packed, residual)
    F_, H, W, C = x.shape
    assert C == 128 and w_packed.shape == (128, 1152)
    y = torch.empty((F_, H, W, 128), device=x.device, dtype=torch.bfloat16)
    det = DETERMINISTIC and stats is not None

    def run():
        rc = L.load().m3t_conv3x3_c128_halo(L.ptr(x), L.ptr(w_packed), L.ptr(y), L.i32(F_), L.i32(H), L.i32(W),
                                            L.ptr(scale), L.ptr(shift), L.ptr(residual), L.i32(int(bool(relu)) | (256 if det else 0)),
                                            L.ptr(stats), L.stream_ptr())
        L.check(rc, "m3t_conv3x3_c128_halo")
        if det:
            _det_reduce(stats, _det_slots("stats"))

    g = conv_geom(2, F_, 1, H, W, 128, 128, (1, 3, 3), (1, 1, 1), (0, 1, 1), (0, 1, 1), (1, 1, 1))
    _timed(conv_key(tag + "-halo", g), 2.0 * F_ * H * W * 128 * 128 * 9, run)
    return y


def conv_fprop(x, w_packed, g, scale=None, shift=None, residual=None, relu=False, stats=None, tile_hint=0,
               algo_flops=None, tag="fprop"):
    """x: bf16 [N,D,H,W,Cin] contiguous (any view with that element order), w_packed: bf16 [Cout, taps*Cin].
    Returns bf16 [N,Z,P,Q,Cout].  algo_flops overrides the algorithmic FLOP count credited to this launch
    (e.g. the space-to-depth stem computes padded taps; a stride-2 dgrad runs over a zero-inserted map)."""
    _chk_bf16(x, w_packed, residual)
    Z, P, Q = conv_out_dims(g)
    N, Cout = g[1], g[6]
    if halo128_route(g, tile_hint):
        y = conv3x3_c128_halo(x.view(N, g[3], g[4], 128), w_packed, scale, shift,
                              residual.view(N, g[3], g[4], 128) if residual is not None else None, relu, stats, tag)
        return y.view(N, 1, g[3], g[4], 128)
    if halo_route(g, tile_hint):
        y = conv3x3_c64_halo(x.view(N, g[3], g[4], 64), w_packed, scale, shift,
                             residual.view(N, g[3], g[4], 64) if residual is not None else None, relu, stats, tag)
        return y.view(N, 1, g[3], g[4], 64)
    y = torch.empty((N, Z, P, Q, Cout), device=x.device, dtype=torch.bfloat16)

    def run():
        det = DETERMINISTIC and stats is not None
        rc = L.load().m3t_conv_fprop_bf16(L.ptr(x), L.ptr(w_packed), L.ptr(y), L.int_array(g), L.ptr(scale),
                                          L.ptr(shift), L.ptr(residual), L.i32(relu), L.ptr(stats),
                                          L.i32(tile_hint | (256 if det else 0)), L.stream_ptr())
        L.check(rc, "m3t_conv_fprop_bf16")
        if det:
            _det_reduce(stats, _det_slots("stats"))

    if _prof is not None and algo_flops is None:
        algo_flops = 2.0 * N * Z * P * Q * Cout * g[5] * g[7] * g[8] * g[9]
    _timed(conv_key(tag, g), algo_flops, run)
    return y


_DROPOUT_COUNTERS = {}
_DROPOUT_STRIDE = 0x2545F4914F6CDD1D          # odd 63-bit constant: the counter walks all of Z / 2^64


def dropout_counter(device):
    """64-bit device-resident counter added to every fused-dropout seed drawn while a CUDA graph is being captured
    (m3t_tcn_conv_bf16_dseed): a captured step freezes its launch arguments, so what changes between replays has to
    live in device memory.  advance_dropout_counter() is part of the captured step (engine._eager_step)."""
    dev = torch.device(device)
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    t = _DROPOUT_COUNTERS.get(key)
    if t is None:
        t = _DROPOUT_COUNTERS[key] = torch.zeros(1, device=dev, dtype=torch.int64)
    return t


def advance_dropout_counter(device):
    dropout_counter(device).add_(_DROPOUT_STRIDE)       # int64 wrap-around = arithmetic mod 2^64, as the kernel adds


def tcn_conv(x, w_packed, g, scale, shift, residual=None, drop_p=0.0, seed=0, want_t=False, seed_dev=None):
    """One TemporalBlock conv with its epilogue (m3t_tcn_conv_bf16): t = dropout(relu(conv * scale + shift));
    y = relu(t + residual) (t also returned when want_t) or y = t.  x: CL bf16 [B,T,Cin].
    seed_dev: optional int64[1] device tensor added to `seed` on the device (dropout_counter)."""
    _chk_bf16(x, w_packed, residual)
    Z, P, Q = conv_out_dims(g)
    N, Cout = x.shape[0], w_packed.shape[0]
    y = torch.empty((N, Q, Cout), device=x.device, dtype=torch.bfloat16)
    t = torch.empty_like(y) if (want_t and residual is not None) else None
    flops = 2.0 * N * Q * Cout * w_packed.shape[1]

    def run():
        if seed_dev is not None:
            L.check(L.load().m3t_tcn_conv_bf16_dseed(L.ptr(x), L.ptr(w_packed), L.ptr(y), L.ptr(t), L.int_array(g),
                                                     L.ptr(scale), L.ptr(shift), L.ptr(residual), L.f32(drop_p),
                                                     ctypes.c_ulonglong(int(seed)), L.ptr(seed_dev), L.stream_ptr()),
                    "tcn_conv_dseed")
            return
        L.check(L.load().m3t_tcn_conv_bf16(L.ptr(x), L.ptr(w_packed), L.ptr(y), L.ptr(t), L.int_array(g), L.ptr(scale),
                                           L.ptr(shift),
                                           L.ptr(residual), L.f32(drop_p), ctypes.c_ulonglong(int(seed)),
                                           L.stream_ptr()), "tcn_conv")
    _timed(conv_key("tcn", g), flops, run)
    return y, t


def tcn_epilogue_bwd(dy, y, t, scale):
    """(dsum, da) of tcn_conv's epilogue; y is None when no residual joined (then dsum is None and t is the output)."""
    _chk_bf16(dy, y, t)
    da = torch.empty_like(dy)
    dsum = torch.empty_like(dy) if y is not None else None
    L.check(_lib().m3t_tcn_epilogue_bwd_bf16(L.ptr(dy), L.ptr(y), L.ptr(t), L.ptr(dsum), L.ptr(da), L.f32(scale),
                                             L.i64(dy.numel()), L.stream_ptr()), "tcn_epilogue_bwd")
    return dsum, da


def conv_fprop_scatter(x, w_packed, g, out_base, img_pitch, row_pitch, px_pitch, accumulate=False, algo_flops=None,
                       tag="dgrad-s2"):
    """conv_fprop (2-D, no epilogue arithmetic) storing output pixel (n,p,q) at out_base + (n*img_pitch + p*row_pitch +
    q*px_pitch) * Cout elements.  out_base: bf16 tensor view whose data pointer is the address of pixel (0,0,0).
    accumulate: add to what is stored there instead of overwriting."""
    _chk_bf16(x, w_packed, out_base)
    Z, P, Q = conv_out_dims(g)
    N, Cout = g[1], g[6]

    def run():
        rc = L.load().m3t_conv_fprop_scatter_bf16(L.ptr(x), L.ptr(w_packed), L.ptr(out_base), L.int_array(g),
                                                  L.i64(img_pitch), L.i64(row_pitch), L.i64(px_pitch),
                                                  L.i32(1 if accumulate else 0), L.i32(0), L.stream_ptr())
        L.check(rc, "m3t_conv_fprop_scatter_bf16")

    if _prof is not None and algo_flops is None:
        algo_flops = 2.0 * N * Z * P * Q * Cout * g[5] * g[7] * g[8] * g[9]
    _timed(conv_key(tag, g), algo_flops, run)


USE_HALO = True   # route 64->64 3x3/s1/p1 convolutions to the persistent halo-tile kernel


def conv3x3_c64_halo(x, w_packed, scale=None, shift=None, residual=None, relu=False, stats=None, tag="fprop"):
    """x: bf16 [F,H,W,64]; w_packed bf16 [64, 576] -> bf16 [F,H,W,64]."""
    _chk_bf16(x, w_packed, residual)
    F_, H, W, C = x.shape
    assert C == 64 and w_packed.shape == (64, 576)
    y = torch.empty((F_, H, W, 64), device=x.device, dtype=torch.bfloat16)
    det = DETERMINISTIC and stats is not None

    def run():
        rc = L.load().m3t_conv3x3_c64_halo(L.ptr(x), L.ptr(w_packed), L.ptr(y), L.i32(F_), L.i32(H), L.i32(W),
                                           L.ptr(scale), L.ptr(shift), L.ptr(residual), L.i32(int(bool(relu)) | (256 if det else 0)),
                                           L.ptr(stats), L.stream_ptr())
        L.check(rc, "m3t_conv3x3_c64_halo")
        if det:
            _det_reduce(stats, _det_slots("stats"))

    _timed("%s-halo nd2 1x%dx%d c64->64 k1x3x3 s1" % (tag, H, W), 2.0 * F_ * H * W * 64 * 576, run)
    return y


USE_HALO_WGRAD = True
USE_HALO_STEM = True


def stem_fprop_halo(xs, w_packed, stats=None, algo_flops=None):
    """xs bf16 [B,T,H2,W2,64], w_packed bf16 [64,1280] -> raw conv output bf16 [B*T,H2,W2,64] (+ column stats)."""
    _chk_bf16(xs, w_packed)
    B, T, H2, W2, _ = xs.shape
    y = torch.empty((B * T, H2, W2, 64), device=xs.device, dtype=torch.bfloat16)

    def run():
        L.check(L.load().m3t_stem_fprop_halo(L.ptr(xs), L.ptr(w_packed), L.ptr(y), L.i32(B), L.i32(T), L.i32(H2),
                                             L.i32(W2), L.ptr(None), L.ptr(None),
                                             L.i32(256 if (DETERMINISTIC and stats is not None) else 0), L.ptr(stats),
                                             L.stream_ptr()), "m3t_stem_fprop_halo")
        if DETERMINISTIC and stats is not None:
            _det_reduce(stats, _det_slots("stats"))

    _timed("stem-halo %dx%dx%d" % (T, H2, W2), algo_flops or 2.0 * B * T * H2 * W2 * 64 * 1280, run)
    return y


def wgrad_stem_halo(xs, dy, algo_flops=None):
    """Stem weight gradient over the W-unrolled s2d image: xs bf16 [B,T,H2,W2,64], dy bf16 [B*T,H2,W2,64] ->
    fp32 [64, 1280] (packed (kt,jh,jw,ch))."""
    _chk_bf16(xs, dy)
    B, T, H2, W2, _ = xs.shape
    if DETERMINISTIC:
        dw, _ = _det_accum((64, 1280), _det_slots("cta"), xs.device)
    else:
        dw = zeros_f32((64, 1280), xs.device)

    def run():
        fn = L.load().m3t_wgrad_stem_halo_det if DETERMINISTIC else L.load().m3t_wgrad_stem_halo
        L.check(fn(L.ptr(xs), L.ptr(dy), L.ptr(dw), L.i32(B), L.i32(T), L.i32(H2), L.i32(W2), L.stream_ptr()),
                "m3t_wgrad_stem_halo")
        if DETERMINISTIC:
            _det_reduce(dw, _det_slots("cta"))

    _timed("wgrad-halo stem %dx%dx%d" % (T, H2, W2), algo_flops or 2.0 * B * T * H2 * W2 * 64 * 1280, run)
    return dw


def conv_wgrad(x, dy, g, splits=0, algo_flops=None):
    """Returns fp32 [Cout, taps*Cin] (packed, tap-major / channel-minor)."""
    _chk_bf16(x, dy)
    Cin, Cout = g[5], g[6]
    taps = g[7] * g[8] * g[9]
    halo = (USE_HALO_WGRAD and splits == 0 and g[0] == 2 and Cin == 64 and Cout == 64 and tuple(g[7:10]) == (1, 3, 3)
            and tuple(g[10:13]) == (1, 1, 1) and tuple(g[13:19]) == (0, 0, 1, 1, 1, 1)
            and tuple(g[19:22]) == (1, 1, 1) and g[4] + 2 <= 48)
    det_slots = 0
    if DETERMINISTIC:
        det_slots = _det_slots("cta") if halo else int(L.load().m3t_conv_wgrad_splits(L.int_array(g), L.i32(splits)))
        if det_slots <= 0:
            raise L.M3TError("m3t_conv_wgrad_splits -> %d" % det_slots)
        dw, _ = _det_accum((Cout, taps * Cin), det_slots, x.device)
    else:
        dw = zeros_f32((Cout, taps * Cin), x.device)
    if halo:
        N, H, W = g[1], g[3], g[4]

        def run_h():
            fn = L.load().m3t_wgrad3x3_c64_halo_det if DETERMINISTIC else L.load().m3t_wgrad3x3_c64_halo
            L.check(fn(L.ptr(x), L.ptr(dy), L.ptr(dw), L.i32(N), L.i32(H), L.i32(W), L.stream_ptr()),
                    "m3t_wgrad3x3_c64_halo")
            if DETERMINISTIC:
                _det_reduce(dw, det_slots)

        _timed("wgrad-halo nd2 1x%dx%d c64->64 k1x3x3 s1" % (H, W), 2.0 * N * H * W * 64 * 576, run_h)
        return dw

    def run():
        rc = L.load().m3t_conv_wgrad_bf16(L.ptr(x), L.ptr(dy), L.ptr(dw), L.int_array(g),
                                          L.i32(splits | ((1 << 29) if DETERMINISTIC else 0)), L.stream_ptr())
        L.check(rc, "m3t_conv_wgrad_bf16")
        if DETERMINISTIC:
            _det_reduce(dw, det_slots)

    flops = None
    if _prof is not None:
        Z, P, Q = conv_out_dims(g)
        flops = algo_flops if algo_flops is not None else 2.0 * g[1] * Z * P * Q * Cout * Cin * taps
    _timed(conv_key("wgrad", g), flops, run)
    return dw


# ----------------------------------------------------------------------------------------------------------
# HBM-bound passes (elementwise.cu), GRU recurrence (gru.cu), attention mix (fusion.cu)
# ----------------------------------------------------------------------------------------------------------
def _lib():
    return L.load()


def video_prep_s2d(video, normalise):
    """video: (B,3,T,H,W) float32 or uint8 -> bf16 (B,T,H/2,W/2,16) space-to-depth, channel (ph*2+pw)*3+c.
    normalise=True applies (x-127.5)/127.5 (models/model.py:106); False passes values through."""
    assert video.is_cuda and video.is_contiguous() and video.dtype in (torch.float32, torch.uint8)
    B, C, T, H, W = video.shape
    assert C == 3
    out = torch.empty((B, T, H // 2, W // 2, 16), device=video.device, dtype=torch.bfloat16)
    mul, add = (1.0 / 127.5, -1.0) if normalise else (1.0, 0.0)
    L.check(_lib().m3t_video_prep_s2d(L.ptr(video), L.i32(video.dtype == torch.uint8), L.ptr(out), L.i32(B), L.i32(T),
                                      L.i32(H), L.i32(W), L.f32(mul), L.f32(add), L.stream_ptr()), "video_prep_s2d")
    return out


def video_prep_s2d_w4(video, normalise):
    """As video_prep_s2d, unrolled over the 4 horizontal stem taps: bf16 (B,T,H/2,W/2,64)."""
    assert video.is_cuda and video.is_contiguous() and video.dtype in (torch.float32, torch.uint8)
    B, C, T, H, W = video.shape
    assert C == 3
    out = torch.empty((B, T, H // 2, W // 2, 64), device=video.device, dtype=torch.bfloat16)
    mul, add = (1.0 / 127.5, -1.0) if normalise else (1.0, 0.0)
    L.check(_lib().m3t_video_prep_s2d_w4(L.ptr(video), L.i32(video.dtype == torch.uint8), L.ptr(out), L.i32(B),
                                         L.i32(T), L.i32(H), L.i32(W), L.f32(mul), L.f32(add), L.stream_ptr()),
            "video_prep_s2d_w4")
    return out


def video_augment_prep_s2d_w4(frames_u8, params, H, W, normalise):
    """frames_u8: uint8 [B,T,Hs,Ws,3] decoded frames; params: int32 [B,8] (include/m3t_b200.h) -> bf16 (B,T,H/2,W/2,64)."""
    assert frames_u8.is_cuda and frames_u8.is_contiguous() and frames_u8.dtype == torch.uint8 and frames_u8.shape[-1] == 3
    assert params.is_cuda and params.dtype == torch.int32 and params.is_contiguous()
    B, T, Hs, Ws, _ = frames_u8.shape
    assert params.shape == (B, 8)
    out = torch.empty((B, T, H // 2, W // 2, 64), device=frames_u8.device, dtype=torch.bfloat16)
    mul, add = (1.0 / 127.5, -1.0) if normalise else (1.0, 0.0)
    L.check(_lib().m3t_video_augment_prep_s2d_w4(L.ptr(frames_u8), L.ptr(params), L.ptr(out), L.i32(B), L.i32(T),
                                                 L.i32(Hs), L.i32(Ws), L.i32(H), L.i32(W), L.f32(mul), L.f32(add),
                                                 L.stream_ptr()), "video_augment_prep_s2d_w4")
    return out


def bn_finalize(stats, count, gamma, beta, eps, momentum, running_mean, running_var):
    C = gamma.numel()
    buf = torch.empty((4, C), device=gamma.device, dtype=torch.float32)  # mean, invstd, scale, shift
    L.check(_lib().m3t_bn_finalize(L.ptr(stats), L.i32(C), ctypes_double(count), L.ptr(gamma), L.ptr(beta), L.f32(eps),
                                   L.f32(momentum), L.ptr(running_mean), L.ptr(running_var), L.ptr(buf[0]),
                                   L.ptr(buf[1]), L.ptr(buf[2]), L.ptr(buf[3]), L.stream_ptr()), "bn_finalize")
    return buf


def bn_fold(gamma, beta, running_mean, running_var, conv_bias, eps):
    C = gamma.numel()
    buf = torch.empty((2, C), device=gamma.device, dtype=torch.float32)  # scale, shift
    L.check(_lib().m3t_bn_fold(L.i32(C), L.ptr(gamma), L.ptr(beta), L.ptr(running_mean), L.ptr(running_var),
                               L.ptr(conv_bias), L.f32(eps), L.ptr(buf[0]), L.ptr(buf[1]), L.stream_ptr()), "bn_fold")
    return buf


def ctypes_double(v):
    import ctypes
    return ctypes.c_double(float(v))


def bn_act(y, scale, shift, res=None, res_scale=None, res_shift=None, relu=True):
    C = y.shape[-1]
    rows = y.numel() // C
    out = torch.empty_like(y)
    _timed("hbm bn_act C%d" % C, 0.0, lambda: L.check(
        _lib().m3t_bn_act(L.ptr(y), L.ptr(scale), L.ptr(shift), L.ptr(res), L.ptr(res_scale), L.ptr(res_shift),
                          L.i32(relu), L.ptr(out), L.i64(rows), L.i32(C), L.stream_ptr()), "bn_act"),
           _nb(y, res, out))
    return out


def _relu_mode(relu, out):
    """0 = no ReLU, 1 = mask from the stored activation, 2 = mask recomputed from y*scale+shift (out is None)."""
    if not relu:
        return 0
    return 1 if out is not None else 2


def bn_bwd_reduce(dout, out, y, mean, invstd, relu, want_dz, scale=None, shift=None):
    C = y.shape[-1]
    rows = y.numel() // C
    sums = new_stats(C, y.device)
    dz = torch.empty_like(y) if want_dz else None
    mode = _relu_mode(relu, out)
    assert mode != 2 or (scale is not None and shift is not None)
    det = DETERMINISTIC
    _timed("hbm bn_bwd_reduce C%d" % C, 0.0, lambda: L.check(
        _lib().m3t_bn_bwd_reduce(L.ptr(dout), L.ptr(out), L.ptr(y), L.ptr(mean), L.ptr(invstd), L.ptr(scale),
                                 L.ptr(shift), L.i32(mode | (256 if det else 0)), L.ptr(dz), L.ptr(sums), L.i64(rows),
                                 L.i32(C), L.stream_ptr()), "bn_bwd_reduce"), _nb(dout, out, y, dz))
    if det:
        _det_reduce(sums, _det_slots("stats"))
    return sums, dz


def bn_bwd_apply(dout, out, y, mean, invstd, scale, sums, count, relu, shift=None):
    C = y.shape[-1]
    rows = y.numel() // C
    dy = torch.empty_like(y)
    mode = _relu_mode(relu, out)
    assert mode != 2 or shift is not None
    _timed("hbm bn_bwd_apply C%d" % C, 0.0, lambda: L.check(
        _lib().m3t_bn_bwd_apply(L.ptr(dout), L.ptr(out), L.ptr(y), L.ptr(mean), L.ptr(invstd), L.ptr(scale),
                                L.ptr(shift), L.ptr(sums), ctypes_double(count), L.i32(mode), L.ptr(dy),
                                L.i64(rows), L.i32(C), L.stream_ptr()), "bn_bwd_apply"), _nb(dout, out, y, dy))
    return dy


def bn_relu_maxpool(y, scale, shift, want_idx, pool=(3, 2, 1), want_ymax=False):
    """want_ymax: also return the raw y at every window's arg-max (the BN-backward sums then need the pooled tensors
    only: maxpool_bn_bwd(..., ymax=...))."""
    F_, H, W, C = y.shape
    K, S, PAD = pool
    P, Q = (H + 2 * PAD - K) // S + 1, (W + 2 * PAD - K) // S + 1
    out = torch.empty((F_, P, Q, C), device=y.device, dtype=torch.bfloat16)
    idx = torch.empty((F_, P, Q, C), device=y.device, dtype=torch.uint8) if want_idx else None
    ymax = torch.empty_like(out) if want_ymax else None
    _timed("hbm bn_relu_maxpool %dx%d C%d" % (H, W, C), 0.0, lambda: L.check(
        _lib().m3t_bn_relu_maxpool_ymax(L.ptr(y), L.ptr(scale), L.ptr(shift), L.ptr(out), L.ptr(idx), L.ptr(ymax),
                                        L.i32(F_), L.i32(H), L.i32(W), L.i32(C), L.i32(K), L.i32(S), L.i32(PAD),
                                        L.stream_ptr()), "bn_relu_maxpool"), _nb(y, out, idx, ymax))
    if want_ymax:
        return out, idx, ymax
    return out, idx


def maxpool_bn_bwd(dout, idx, y, mean, invstd, scale, shift, count, pool=(3, 2, 1), ymax=None):
    F_, H, W, C = y.shape
    K, S, PAD = pool
    fn = _lib().m3t_maxpool_bn_bwd
    if ymax is not None:
        # every window sends its gradient to ONE position whose raw value is ymax: the sums over the full-resolution map
        # equal the sums over the pooled tensors (mask recomputed from ymax * scale + shift > 0)
        sums, _ = bn_bwd_reduce(dout, None, ymax, mean, invstd, True, False, scale=scale, shift=shift)
    else:
        if DETERMINISTIC:
            raise L.M3TError("deterministic mode: the pooled units take their BatchNorm sums from ymax (want_ymax=True)")
        sums = zeros_f32((2, C), y.device)
        _timed("hbm maxpool_bn_bwd(reduce) %dx%d C%d" % (H, W, C), 0.0, lambda: L.check(
            fn(L.i32(0), L.ptr(dout), L.ptr(idx), L.ptr(y), L.ptr(mean), L.ptr(invstd), L.ptr(scale), L.ptr(shift),
               L.ptr(sums), ctypes_double(count), L.ptr(None), L.i32(F_), L.i32(H), L.i32(W), L.i32(C), L.i32(K),
               L.i32(S), L.i32(PAD), L.stream_ptr()), "maxpool_bn_bwd(reduce)"), _nb(dout, idx, y))
    dy = torch.empty_like(y)
    _timed("hbm maxpool_bn_bwd(apply) %dx%d C%d" % (H, W, C), 0.0, lambda: L.check(
        fn(L.i32(1), L.ptr(dout), L.ptr(idx), L.ptr(y), L.ptr(mean), L.ptr(invstd), L.ptr(scale), L.ptr(shift),
           L.ptr(sums), ctypes_double(count), L.ptr(dy), L.i32(F_), L.i32(H), L.i32(W), L.i32(C), L.i32(K),
           L.i32(S), L.i32(PAD), L.stream_ptr()), "maxpool_bn_bwd(apply)"), _nb(dout, idx, y, dy))
    return dy, sums


def avgpool(x, out_f32=False):
    F_, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (F_ * C)
    ob = torch.empty((F_, C), device=x.device, dtype=torch.bfloat16)
    of = torch.empty((F_, C), device=x.device, dtype=torch.float32) if out_f32 else None
    L.check(_lib().m3t_avgpool(L.ptr(x), L.ptr(ob), L.ptr(of), L.i32(F_), L.i32(HW), L.i32(C), L.stream_ptr()),
            "avgpool")
    return ob, of


def avgpool_bwd(dout, shape):
    F_, C = shape[0], shape[-1]
    HW = 1
    for d in shape[1:-1]:
        HW *= d
    dx = torch.empty(shape, device=dout.device, dtype=torch.bfloat16)
    L.check(_lib().m3t_avgpool_bwd(L.ptr(dout), L.i32(dout.dtype == torch.float32), L.ptr(dx), L.i32(F_), L.i32(HW),
                                   L.i32(C), L.stream_ptr()), "avgpool_bwd")
    return dx


def ncs_to_nsc_bf16(x, cpad=None):
    """fp32 [N,C,*S] -> bf16 [N,*S,Cpad] (channels-last)."""
    N, C = x.shape[:2]
    S = x.numel() // (N * C)
    cpad = cpad or C
    out = torch.empty([N] + list(x.shape[2:]) + [cpad], device=x.device, dtype=torch.bfloat16)
    L.check(_lib().m3t_ncs_f32_to_nsc_bf16(L.ptr(x), L.ptr(out), L.i32(N), L.i32(C), L.i32(S), L.i32(cpad),
                                           L.stream_ptr()), "ncs_to_nsc")
    return out


def nsc_to_ncs_f32(x, C=None):
    """bf16/fp32 [N,*S,Cpad] -> fp32 [N,C,*S]."""
    N, cpad = x.shape[0], x.shape[-1]
    C = C or cpad
    S = x.numel() // (N * cpad)
    out = torch.empty([N, C] + list(x.shape[1:-1]), device=x.device, dtype=torch.float32)
    L.check(_lib().m3t_nsc_to_ncs_f32(L.ptr(x), L.i32(x.dtype == torch.float32), L.ptr(out), L.i32(N), L.i32(C),
                                      L.i32(S), L.i32(cpad), L.stream_ptr()), "nsc_to_ncs")
    return out


def cast_bf16(x2d, ld_out=None, out=None):
    """fp32 [rows, cols] (row stride = stride(0)) -> bf16 [rows, ld_out] (pad columns zero)."""
    rows, cols = x2d.shape
    assert x2d.dtype == torch.float32 and x2d.stride(1) == 1
    ld_out = ld_out or (cols + 7) // 8 * 8
    if out is None:
        out = torch.empty((rows, ld_out), device=x2d.device, dtype=torch.bfloat16)
    L.check(_lib().m3t_cast_f32_bf16(L.ptr(x2d), L.i64(x2d.stride(0)), L.ptr(out), L.i64(out.stride(0)), L.i64(rows),
                                     L.i32(cols), L.stream_ptr()), "cast_f32_bf16")
    return out


def cast_f32(x2d, cols=None):
    rows = x2d.shape[0]
    cols = cols or x2d.shape[1]
    out = torch.empty((rows, cols), device=x2d.device, dtype=torch.float32)
    L.check(_lib().m3t_cast_bf16_f32(L.ptr(x2d), L.i64(x2d.stride(0)), L.ptr(out), L.i64(cols), L.i64(rows),
                                     L.i32(cols), L.stream_ptr()), "cast_bf16_f32")
    return out


def pack_filter(w, want_dgrad):
    """fp32 [Cout, Cin, *k] -> bf16 [Cout, taps*Cin] (+ dgrad pack [Cin, taps(flipped)*Cout])."""
    Cout, Cin = w.shape[:2]
    taps = w.numel() // (Cout * Cin)
    wf = torch.empty((Cout, taps * Cin), device=w.device, dtype=torch.bfloat16)
    wd = torch.empty((Cin, taps * Cout), device=w.device, dtype=torch.bfloat16) if want_dgrad else None
    L.check(_lib().m3t_pack_filter(L.ptr(w), L.ptr(wf), L.ptr(wd), L.i32(Cout), L.i32(Cin), L.i32(taps),
                                   L.stream_ptr()), "pack_filter")
    return wf, wd


def gru_pack_weights(w_ih, w_ih_r, w_hh, w_hh_r, wih, whh, whht, I, Ipad, H, biases=None, bias_out=None):
    for t in (w_ih, w_ih_r, w_hh, w_hh_r) + tuple(biases or ()):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    b = list(biases) if biases else [None] * 4
    L.check(_lib().m3t_gru_pack_weights(L.ptr(w_ih), L.ptr(w_ih_r), L.ptr(w_hh), L.ptr(w_hh_r), L.ptr(wih), L.ptr(whh),
                                        L.ptr(whht), L.i32(I), L.i32(Ipad), L.i32(H), L.ptr(b[0]), L.ptr(b[1]),
                                        L.ptr(b[2]), L.ptr(b[3]), L.ptr(bias_out), L.stream_ptr()),
            "m3t_gru_pack_weights")


def unpack_filter_grad(dwp, shape):
    Cout, Cin = shape[:2]
    taps = 1
    for d in shape[2:]:
        taps *= d
    dw = torch.empty(shape, device=dwp.device, dtype=torch.float32)
    L.check(_lib().m3t_unpack_filter_grad(L.ptr(dwp), L.ptr(dw), L.i32(Cout), L.i32(Cin), L.i32(taps),
                                          L.stream_ptr()), "unpack_filter_grad")
    return dw


def zero_insert2(dy, H, W):
    N, P, Q, C = dy.shape
    up = torch.empty((N, H, W, C), device=dy.device, dtype=torch.bfloat16)
    L.check(_lib().m3t_zero_insert2(L.ptr(dy), L.ptr(up), L.i32(N), L.i32(P), L.i32(Q), L.i32(H), L.i32(W), L.i32(C),
                                    L.stream_ptr()), "zero_insert2")
    return up


def add_bf16(a, b):
    out = torch.empty_like(a)
    _timed("hbm add_bf16", 0.0, lambda: L.check(
        _lib().m3t_add_bf16(L.ptr(a), L.ptr(b), L.ptr(out), L.i64(a.numel()), L.stream_ptr()), "add_bf16"),
           _nb(a, b, out))
    return out


def gather_pack(w2d, idx, K):
    rows = w2d.shape[0]
    out = torch.empty((rows, K), device=w2d.device, dtype=torch.bfloat16)
    L.check(_lib().m3t_gather_pack_bf16(L.ptr(w2d), L.ptr(idx), L.ptr(out), L.i32(rows), L.i64(w2d.stride(0)), L.i32(K),
                                        L.stream_ptr()), "gather_pack")
    return out


def scatter_unpack(dwp, idx, shape2d):
    rows, K = dwp.shape
    dw = zeros_f32(shape2d, dwp.device)
    L.check(_lib().m3t_scatter_unpack_f32(L.ptr(dwp), L.ptr(idx), L.ptr(dw), L.i32(rows), L.i64(dw.stride(0)), L.i32(K),
                                          L.stream_ptr()), "scatter_unpack")
    return dw


def colsum(x2d, cols=None):
    rows = x2d.shape[0]
    cols = cols or x2d.shape[1]
    out = zeros_f32((cols,), x2d.device)
    L.check(_lib().m3t_colsum_bf16(L.ptr(x2d), L.i64(x2d.stride(0)), L.i64(rows),
                                   L.i32(cols | ((1 << 30) if DETERMINISTIC else 0)), L.ptr(out), L.stream_ptr()),
            "colsum")
    return out


def relu_bwd(dy, out):
    dz = torch.empty_like(dy)
    L.check(_lib().m3t_relu_bwd_bf16(L.ptr(dy), L.ptr(out), L.ptr(dz), L.i64(dy.numel()), L.stream_ptr()), "relu_bwd")
    return dz


def dropout_bf16(x, p, seed):
    y = torch.empty_like(x)
    L.check(_lib().m3t_dropout_bf16(L.ptr(x), L.ptr(y), L.i64(x.numel()), L.f32(p), ctypes.c_ulonglong(int(seed)),
                                    L.stream_ptr()), "dropout_bf16")
    return y


def gru_cluster_default():
    """Small-batch inference recurrences (B <= 64) run on m3t_gru_fwd_cluster (thread-block clusters + DSMEM): bit
    identical to m3t_gru_fwd and 4.0 vs 6.8 us per step on a B200 (tests/gpu_cases.py::case_gru_cluster,
    profiles/r2_next_session.md).  M3T_GRU_CLUSTER=0 switches back to the L2-polling kernel."""
    return os.environ.get("M3T_GRU_CLUSTER", "1") == "1"


def gru_fwd(gi, w_hh_bf16, b_hh, B, T, H, want_saved, want_f32=False, cluster=None):
    dev = gi.device
    out = torch.empty((B, T, 2 * H), device=dev, dtype=torch.bfloat16)
    out32 = torch.empty((B, T, 2 * H), device=dev, dtype=torch.float32) if want_f32 else None
    if cluster is None:
        cluster = gru_cluster_default()
    saved = torch.empty((B * T, 2, 4, H), device=dev, dtype=torch.float32) if want_saved else None
    if cluster and B <= 128 and H % 64 == 0 and H <= 512:
        # one cluster per 16 rows and direction, all resident at once: 64 rows always fit, up to 128 when the device
        # keeps that many clusters of this size resident (the call returns -3 otherwise -> L2 kernel below)
        box = [0]

        def launch():
            box[0] = _lib().m3t_gru_fwd_cluster(L.ptr(gi), L.ptr(w_hh_bf16), L.ptr(b_hh), L.ptr(out), L.ptr(out32),
                                                L.ptr(saved), L.i32(B), L.i32(T), L.i32(H), L.stream_ptr())
            if box[0] != -3:
                L.check(box[0], "gru_fwd_cluster")
        _timed("gru_fwd_cluster B%d T%d H%d" % (B, T, H), 0.0, launch, _nb(gi, w_hh_bf16, out, out32, saved))
        if box[0] != -3:
            return out, out32, saved
    counters = torch.empty((2 * ((B + 31) // 32) + 2,), device=dev, dtype=torch.int32)
    _timed("gru_fwd B%d T%d H%d" % (B, T, H), 0.0, lambda: L.check(
        _lib().m3t_gru_fwd(L.ptr(gi), L.ptr(w_hh_bf16), L.ptr(b_hh), L.ptr(out), L.ptr(out32), L.ptr(saved),
                           L.ptr(counters), L.i32(B), L.i32(T), L.i32(H), L.stream_ptr()), "gru_fwd"),
           _nb(gi, w_hh_bf16, out, out32, saved))
    return out, out32, saved


def gru_bwd_cluster_default():
    """BPTT on the cluster / DSMEM kernel (m3t_gru_bwd_cluster) wherever it applies; M3T_GRU_BWD_CLUSTER=0 switches back
    to the L2 / arrival-counter kernel."""
    return os.environ.get("M3T_GRU_BWD_CLUSTER", "1") == "1"


def gru_bwd(dout, out, saved, w_hh_t_bf16, B, T, H, cluster=None):
    dev = dout.device
    dgi = torch.empty((B * T, 6 * H), device=dev, dtype=torch.bfloat16)
    dgh = torch.empty((B * T, 6 * H), device=dev, dtype=torch.bfloat16)
    hprev = torch.empty((B * T, 2 * H), device=dev, dtype=torch.bfloat16)
    counters = torch.empty((2 * ((B + 31) // 32) + 2,), device=dev, dtype=torch.int32)
    # deterministic mode: the kernel's bias sums are atomics over batch slices and warps; take them as (deterministic)
    # column sums of the gate-gradient tensors instead
    dbias = None if DETERMINISTIC else zeros_f32((2, 6 * H), dev)     # [b_ih | b_hh] x [dir0 3H | dir1 3H]
    if cluster is None:
        cluster = gru_bwd_cluster_default()

    def launch():
        if cluster and H in (128, 256, 512):
            rc = _lib().m3t_gru_bwd_cluster(L.ptr(dout), L.ptr(out), L.ptr(saved), L.ptr(w_hh_t_bf16), L.ptr(dgi),
                                            L.ptr(dgh), L.ptr(hprev), L.ptr(dbias), L.i32(B), L.i32(T), L.i32(H),
                                            L.stream_ptr())
            if rc not in (-23, -3):      # -23: no room for the cluster on this device, -3: batch too large for it
                return L.check(rc, "gru_bwd_cluster")
        L.check(_lib().m3t_gru_bwd(L.ptr(dout), L.ptr(out), L.ptr(saved), L.ptr(w_hh_t_bf16), L.ptr(dgi), L.ptr(dgh),
                                   L.ptr(hprev), L.ptr(counters), L.ptr(dbias), L.i32(B), L.i32(T), L.i32(H),
                                   L.stream_ptr()), "gru_bwd")

    _timed("gru_bwd B%d T%d H%d" % (B, T, H), 0.0, launch, _nb(dout, out, saved, w_hh_t_bf16, dgi, dgh, hprev))
    if dbias is None:
        dbias = torch.stack((colsum(dgi, 6 * H), colsum(dgh, 6 * H)))
    return dgi, dgh, hprev, dbias


def att_mix_fwd(x_a, x_v, s_a, s_v):
    C = x_a.shape[-1]
    rows = x_a.numel() // C
    f = torch.empty_like(x_a)
    _timed("att_mix_fwd %dx%d" % (rows, C), 0.0, lambda: L.check(
        _lib().m3t_att_mix_fwd(L.ptr(x_a), L.ptr(x_v), L.ptr(s_a), L.ptr(s_v), L.ptr(f), L.ptr(None), L.i64(rows),
                               L.i32(C), L.stream_ptr()), "att_mix_fwd"), _nb(x_a, x_v, s_a, s_v, f))
    return f


def att_mix_bwd(df, x_a, x_v, s_a, s_v):
    C = x_a.shape[-1]
    rows = x_a.numel() // C
    dxa, dxv = torch.empty_like(x_a), torch.empty_like(x_v)
    dsa, dsv = torch.empty_like(s_a), torch.empty_like(s_v)
    _timed("att_mix_bwd %dx%d" % (rows, C), 0.0, lambda: L.check(
        _lib().m3t_att_mix_bwd(L.ptr(df), L.ptr(x_a), L.ptr(x_v), L.ptr(s_a), L.ptr(s_v), L.ptr(dxa), L.ptr(dxv),
                               L.ptr(dsa), L.ptr(dsv), L.i64(rows), L.i32(C), L.stream_ptr()), "att_mix_bwd"),
           _nb(df, x_a, x_v, s_a, s_v, dxa, dxv, dsa, dsv))
    return dxa, dxv, dsa, dsv


def patch3x3_c1(x):
    """fp32 [N,H,W] -> bf16 [N*H*W, 16] (3x3 / pad 1 patches, 9 taps + zero pad)."""
    N, H, W = x.shape
    out = torch.empty((N * H * W, 16), device=x.device, dtype=torch.bfloat16)
    L.check(_lib().m3t_patch3x3_c1(L.ptr(x.contiguous()), L.ptr(out), L.i32(N), L.i32(H), L.i32(W), L.stream_ptr()),
            "patch3x3_c1")
    return out


# ----------------------------------------------------------------------------------------------------------
# CBAM gates (cbam.cu): x is channels-last bf16 (F,H,W,C)
# ----------------------------------------------------------------------------------------------------------
def _fsc(x):
    F_, H, W, C = x.shape
    return F_, H * W, C


def cbam_pool_hw(x):
    _chk_bf16(x)
    F_, S, C = _fsc(x)
    avg = torch.empty((F_, C), device=x.device, dtype=torch.float32)
    mx = torch.empty_like(avg)
    arg = torch.empty((F_, C), device=x.device, dtype=torch.int32)
    L.check(_lib().m3t_cbam_pool_hw(L.ptr(x), L.ptr(avg), L.ptr(mx), L.ptr(arg), L.i32(F_), L.i32(S), L.i32(C),
                                    L.stream_ptr()), "cbam_pool_hw")
    return avg, mx, arg


def cbam_pool_hw_bwd(davg, dmx, arg, shape):
    F_, H, W, C = shape
    dx = torch.empty(shape, device=davg.device, dtype=torch.bfloat16)
    L.check(_lib().m3t_cbam_pool_hw_bwd(L.ptr(davg.float()), L.ptr(dmx.float()), L.ptr(arg), L.ptr(dx), L.i32(F_),
                                        L.i32(H * W), L.i32(C), L.stream_ptr()), "cbam_pool_hw_bwd")
    return dx


def cbam_scale_c(x, sc):
    _chk_bf16(x)
    F_, S, C = _fsc(x)
    assert sc.dtype == torch.float32 and sc.shape == (F_, C) and sc.is_contiguous()
    y = torch.empty_like(x)
    L.check(_lib().m3t_cbam_scale_c(L.ptr(x), L.ptr(sc), L.ptr(y), L.i32(F_), L.i32(S), L.i32(C), L.stream_ptr()),
            "cbam_scale_c")
    return y


def cbam_scale_c_bwd(dy, x, sc):
    F_, S, C = _fsc(x)
    dx = torch.empty_like(x)
    dsc = torch.empty_like(sc)
    L.check(_lib().m3t_cbam_scale_c_bwd(L.ptr(dy), L.ptr(x), L.ptr(sc), L.ptr(dx), L.ptr(dsc), L.i32(F_), L.i32(S),
                                        L.i32(C), L.stream_ptr()), "cbam_scale_c_bwd")
    return dx, dsc


def cbam_pool_c(x):
    _chk_bf16(x)
    F_, H, W, C = x.shape
    comp = torch.empty((F_, 2, H, W), device=x.device, dtype=torch.float32)
    carg = torch.empty((F_, H, W), device=x.device, dtype=torch.int32)
    L.check(_lib().m3t_cbam_pool_c(L.ptr(x), L.ptr(comp), L.ptr(carg), L.i32(F_), L.i32(H * W), L.i32(C),
                                   L.stream_ptr()), "cbam_pool_c")
    return comp, carg


def cbam_pool_c_bwd(dcomp, carg, shape):
    F_, H, W, C = shape
    dx = torch.empty(shape, device=dcomp.device, dtype=torch.bfloat16)
    L.check(_lib().m3t_cbam_pool_c_bwd(L.ptr(dcomp.float()), L.ptr(carg), L.ptr(dx), L.i32(F_), L.i32(H * W), L.i32(C),
                                       L.stream_ptr()), "cbam_pool_c_bwd")
    return dx


def cbam_scale_s(x, ss):
    _chk_bf16(x)
    F_, H, W, C = x.shape
    assert ss.dtype == torch.float32 and ss.numel() == F_ * H * W and ss.is_contiguous()
    y = torch.empty_like(x)
    L.check(_lib().m3t_cbam_scale_s(L.ptr(x), L.ptr(ss), L.ptr(y), L.i64(F_ * H * W), L.i32(C), L.stream_ptr()),
            "cbam_scale_s")
    return y


def cbam_scale_s_bwd(dy, x, ss):
    F_, H, W, C = x.shape
    dx = torch.empty_like(x)
    dss = torch.empty_like(ss)
    L.check(_lib().m3t_cbam_scale_s_bwd(L.ptr(dy), L.ptr(x), L.ptr(ss), L.ptr(dx), L.ptr(dss), L.i64(F_ * H * W),
                                        L.i32(C), L.stream_ptr()), "cbam_scale_s_bwd")
    return dx, dss


def cbam_conv5(comp, w):
    F_, two, H, W = comp.shape
    assert two == 2 and comp.dtype == torch.float32 and comp.is_contiguous() and tuple(w.shape) == (1, 2, 5, 5)
    out = torch.empty((F_, 1, H, W), device=comp.device, dtype=torch.float32)
    L.check(_lib().m3t_cbam_conv5(L.ptr(comp), L.ptr(w.detach().contiguous()), L.ptr(out), L.i32(F_), L.i32(H),
                                  L.i32(W), L.stream_ptr()), "cbam_conv5")
    return out


def cbam_conv5_bwd(dout, comp, w):
    F_, _, H, W = comp.shape
    din = torch.empty_like(comp)
    dw = torch.empty((1, 2, 5, 5), device=comp.device, dtype=torch.float32)
    L.check(_lib().m3t_cbam_conv5_bwd(L.ptr(dout.float()), L.ptr(comp), L.ptr(w.detach().contiguous()), L.ptr(din),
                                      L.ptr(dw), L.i32(F_), L.i32(H), L.i32(W), L.stream_ptr()), "cbam_conv5_bwd")
    return din, dw


def avgpool2x2(x):
    """CL bf16 (F,H,W,C) -> (F,H//2,W//2,C): mean over 2x2 windows (floor)."""
    _chk_bf16(x)
    F_, H, W, C = x.shape
    y = torch.empty((F_, H // 2, W // 2, C), device=x.device, dtype=torch.bfloat16)
    L.check(_lib().m3t_avgpool2x2(L.ptr(x), L.ptr(y), L.i32(F_), L.i32(H), L.i32(W), L.i32(C), L.stream_ptr()),
            "avgpool2x2")
    return y


def avgpool2x2_bwd(dy, shape):
    F_, H, W, C = shape
    dx = torch.empty(shape, device=dy.device, dtype=torch.bfloat16)
    L.check(_lib().m3t_avgpool2x2_bwd(L.ptr(dy), L.ptr(dx), L.i32(F_), L.i32(H), L.i32(W), L.i32(C), L.stream_ptr()),
            "avgpool2x2_bwd")
    return dx
