"""Data-parallel training step of the hot path: forward, loss, backward, gradient all-reduce, clip, Adam
(reference: models/model.py:146-218 `training_step`, :375-390 Adam(lr, weight_decay=1e-4); train.py:35
gradient_clip_val=1.0; Lightning 'ddp' = one process per GPU + NCCL all-reduce of the gradients).

Clips are independent, so ranks only meet in ONE collective per step: the all-reduce (sum) of the flat gradient
arena over NCCL/NVLink; the 1/world mean is folded into the optimiser kernel.  BatchNorm statistics and the CCC
loss stay per-rank, as in the reference (no SyncBN).

Parameters and gradients live in two flat fp32 arenas (every `p.data` / `p.grad` is a view), so the collective, the
global-norm clip and Adam each are ONE call over one buffer (optim.cu), with no host synchronisation.

Optional (overlap_allreduce=True / M3T_OVERLAP_ALLREDUCE=1): the all-reduce split into a few contiguous buckets of
the gradient arena and overlapped with backward: from the second step on every `p.grad` IS its arena
view (autograd accumulates in place into the zeroed arena), a post-accumulate hook counts the parameters of each
bucket, and the bucket's `all_reduce(async_op=True)` is issued the moment its last gradient lands.  It is OFF by
default because it measured slower on 2 x B200 (31.9-32.0 ms vs 31.4 ms per step; the single-GPU step is 31.2 ms)
and again on 8 x B200 in round 2 (31.15 vs 30.52 ms; with SMs reserved for NCCL, M3T_OVERLAP_SM_RESERVE, 31.5-31.8):
the conv kernels are persistent, one CTA per SM, so the SMs NCCL's kernels occupy while they overlap turn a
one-wave launch into a two-wave one, and leaving SMs free for them costs more than the 0.85 ms collective.
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import lib as L
from . import ops


class TrainEngine:
    def __init__(self, model, lr=5e-5, weight_decay=1e-4, clip=1.0, betas=(0.9, 0.999), eps=1e-8,
                 overlap_allreduce=None):
        self.model = model
        self.all_params = [p for p in model.parameters() if p.requires_grad]
        self.clip = float(clip) if clip else 0.0
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.steps = 0
        self.on_gpu = self.all_params[0].device.type == "cuda"
        self.params = None          # the arena is laid out after the first backward (see _build_arena)
        for p in self.all_params:
            p.grad = None
        self.nbt = [m.num_batches_tracked for m in model.modules()
                    if isinstance(m, torch.nn.modules.batchnorm._BatchNorm) and m.num_batches_tracked is not None]
        self.overlap = False        # armed after the arena exists (second step on) when world > 1
        self.num_buckets = int(os.environ.get("M3T_OVERLAP_BUCKETS", "4"))
        if overlap_allreduce is None:
            overlap_allreduce = os.environ.get("M3T_OVERLAP_ALLREDUCE", "0") == "1"
        self.want_overlap = bool(overlap_allreduce)

    def _build_arena(self):
        """Lay out the flat arenas over the parameters that actually receive a gradient.  Parameters autograd never
        reaches (e.g. `resnet.fc`, present in the state_dict but unused with agg_mode 'ap', models/resnet.py:72,117)
        stay outside, exactly as torch.optim.Adam skips parameters whose .grad is None (no weight decay on them)."""
        old = None
        if self.params is not None:      # re-layout: a parameter outside the arena received its first gradient
            old = {id(p): (o, p.numel()) for p, o in zip(self.params, self.offs)}
            old_m, old_v = (self.m, self.v) if self.on_gpu else (None, None)
            in_arena = set(old)
            self.params = [p for p in self.all_params if id(p) in in_arena or p.grad is not None]
        else:
            self.params = [p for p in self.all_params if p.grad is not None]
        dev = self.params[0].device
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4       # 16-byte aligned slots
        self.n, self.offs = n, offs
        self._in_arena = {id(p) for p in self.params}
        self._outside = [p for p in self.all_params if id(p) not in self._in_arena]
        if self.world > 1:
            # every rank must have laid out the same arena (same parameters reached by its first backward), or the
            # flat all-reduce would add misaligned buffers
            names = {id(p): i for i, p in enumerate(self.all_params)}
            sig = torch.tensor([len(self.params), n, sum((names[id(p)] + 1) * (o + 1) % 1000003
                                                         for p, o in zip(self.params, offs))],
                               dtype=torch.int64, device=dev)
            lo, hi = sig.clone(), sig.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            if not torch.equal(lo, hi):
                raise RuntimeError("TrainEngine: ranks built different gradient arenas (a parameter received a "
                                   "gradient on some ranks only): %s vs %s" % (lo.tolist(), hi.tolist()))
        self.flat_p = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        self.grad_views = []
        for p, o in zip(self.params, offs):
            self.flat_p[o:o + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.flat_p[o:o + p.numel()].view_as(p)
            self.grad_views.append(self.flat_g[o:o + p.numel()].view_as(p))
        if self.on_gpu:
            self.m = torch.zeros(n, device=dev, dtype=torch.float32)
            self.v = torch.zeros(n, device=dev, dtype=torch.float32)
            if old is not None:          # carry the Adam moments of the parameters that were already there
                for p, o in zip(self.params, offs):
                    if id(p) in old:
                        oo, k = old[id(p)]
                        self.m[o:o + k].copy_(old_m[oo:oo + k])
                        self.v[o:o + k].copy_(old_v[oo:oo + k])
            self.gnorm_sq = torch.zeros(1, device=dev, dtype=torch.float32)
            L.load().m3t_sumsq_workspace_floats.restype = ctypes.c_longlong
            self.sumsq_ws = torch.empty(int(L.load().m3t_sumsq_workspace_floats()), device=dev, dtype=torch.float32)
        else:  # host-side logic tests only (gloo); the product path is the CUDA one
            prev = getattr(self, "opt", None)
            self.opt = torch.optim.Adam(self.params, lr=self.lr, weight_decay=self.wd, betas=self.betas,
                                        eps=self.eps)
            if prev is not None:
                for p_, st in prev.state.items():
                    self.opt.state[p_] = st

    def _gather_grads(self):
        """autograd hands every parameter a fresh gradient tensor (p.grad is None before backward, so nothing is
        accumulated); one multi-tensor copy moves them into the flat arena."""
        if self.params is None:
            self._build_arena()
        elif any(p.grad is not None for p in self._outside):
            # torch.optim.Adam decides per step which parameters it updates; the arena is laid out once, so a parameter
            # that gets its first gradient later (conditional branch, requires_grad switched on) joins it now.
            # NOTE the fused kernel keeps ONE step counter: a late joiner's bias correction starts at the engine's step
            # count (its moments start at zero), torch.optim.Adam would start it at 1.
            if self.overlap:
                raise RuntimeError("TrainEngine: a parameter outside the gradient arena received a gradient while the "
                                   "bucketed all-reduce is armed; construct the engine after the graph is final")
            self._build_arena()
        dst, src = [], []
        for p, v in zip(self.params, self.grad_views):
            if p.grad is None:
                v.zero_()
            else:
                dst.append(v)
                src.append(p.grad)
        torch._foreach_copy_(dst, src)
        for p in self.params:
            p.grad = None

    # ---- bucketed all-reduce overlapped with backward (world > 1) ----
    def _arm_overlap(self):
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]
        target = (self.n + self.num_buckets - 1) // self.num_buckets
        self.buckets, start, acc, first = [], 0, 0, 0          # (offset_begin, offset_end, n_params)
        self.bucket_of = []
        for i, sz in enumerate(sizes):
            acc += sz
            self.bucket_of.append(len(self.buckets))
            if acc - start >= target or i == len(sizes) - 1:
                self.buckets.append((start, acc, i + 1 - first))
                start, first = acc, i + 1
        for i, (p, v) in enumerate(zip(self.params, self.grad_views)):
            p.grad = v
            p.register_post_accumulate_grad_hook(lambda _p, b=self.bucket_of[i]: self._on_grad(b))
        # M3T_OVERLAP_SM_RESERVE=k: the overlapped buckets run on their own communicator capped at k CTAs and the
        # library's persistent kernels leave k SMs free while a bucket is in flight (m3t_set_sm_reserve), so that
        # NCCL's CTAs do not turn 148-CTA one-wave launches into two-wave ones.  Measured at N = 8 (round 2, ms per
        # 256-clip step): blocking all-reduce 30.52; overlapped 31.15 (k = 0), 31.51 (k = 4), 31.82 (k = 8) - the
        # reserve costs more compute than the contention it removes, and the overlapped path's in-place gradient
        # accumulation into the zeroed arena costs more than the 0.85 ms collective it hides.  Off by default.
        self.sm_reserve = int(os.environ.get("M3T_OVERLAP_SM_RESERVE", "0")) if self.on_gpu else 0
        self._pg = None
        if self.on_gpu and self.sm_reserve > 0 and dist.get_backend() == "nccl":
            try:
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = self.sm_reserve
                opts.config.min_ctas = 1
                self._pg = dist.new_group(backend="nccl", pg_options=opts)
            except Exception:       # older bindings without per-communicator config: fall back to the default group
                self._pg = None
        self.overlap = True

    def _issue_bucket(self, b):
        lo, hi, _ = self.buckets[b]
        if self.on_gpu and self.sm_reserve > 0 and not self._reserved:
            L.load().m3t_set_sm_reserve(self.sm_reserve)
            self._reserved = True
        self._works.append(dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.SUM, async_op=True, group=self._pg))

    def _on_grad(self, b):
        if not self.overlap or self._pending is None:
            return
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._issue_bucket(b)

    def _backward_overlapped(self, loss):
        self.flat_g.zero_()
        self._pending = [n for _, _, n in self.buckets]
        self._works = []
        self._reserved = False
        try:
            loss.backward()
        finally:
            if self._reserved:
                L.load().m3t_set_sm_reserve(0)
                self._reserved = False
        for b, left in enumerate(self._pending):       # buckets holding a parameter that got no gradient this step
            if left > 0:
                self._issue_bucket(b)
        if self._reserved:
            L.load().m3t_set_sm_reserve(0)
            self._reserved = False
        self._pending = None
        for w in self._works:
            w.wait()
        if not self.on_gpu:
            self.flat_g.mul_(1.0 / self.world)

    def _allreduce_grads(self):
        if self.world > 1 and os.environ.get("M3T_DEBUG_SKIP_ALLREDUCE") != "1":   # debug knob: time a step without it
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)
            if not self.on_gpu:
                self.flat_g.mul_(1.0 / self.world)

    def _optimizer_step(self, count=True):
        if count:
            self.steps += 1
        if not self.on_gpu:
            for p, v in zip(self.params, self.grad_views):
                p.grad = v
            if self.clip:
                torch.nn.utils.clip_grad_norm_(self.params, self.clip)
            self.opt.step()
            if not self.overlap:
                for p in self.params:
                    p.grad = None
            return
        lib = L.load()
        st = L.stream_ptr()
        if self.clip:
            L.check(lib.m3t_sumsq_f32(L.ptr(self.flat_g), L.i64(self.n), L.ptr(self.gnorm_sq), L.ptr(self.sumsq_ws), st),
                    "sumsq")
        # lr / weight decay / step count live on the device (hyper): nothing a scheduler or the step counter changes is
        # a kernel argument, so the same launch sequence can be replayed from a CUDA graph
        self._sync_hyper()
        L.check(lib.m3t_adam_clip_step_dev(L.ptr(self.flat_p), L.ptr(self.flat_g), L.ptr(self.m), L.ptr(self.v),
                                           L.i64(self.n), L.ptr(self.hyper), L.f32(self.betas[0]),
                                           L.f32(self.betas[1]), L.f32(self.eps), L.f32(self.clip),
                                           L.f32(1.0 / self.world), L.ptr(self.gnorm_sq), st), "adam_clip_step_dev")
        # the kernel rewrote every parameter through the arena pointer: PyTorch's per-tensor version counters did not
        # move, so the packed / bf16 copies derived from the old values must be dropped explicitly
        ops.clear_caches()

    def _sync_hyper(self, force=False):
        """Mirror (lr, weight_decay[, step count]) into the device-resident optimiser state when they changed on the host
        (schedulers, checkpoint restore).  The copy is stream-ordered and OUTSIDE any captured graph."""
        if getattr(self, "hyper", None) is None:
            self.hyper = torch.zeros(8, device=self.flat_p.device, dtype=torch.float32)
            self._hyper_host = None
            force = True
        want = (float(self.lr), float(self.wd))
        if force or self._hyper_host != want:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("TrainEngine: lr / weight decay changed inside a graph capture")
            if force:       # (re)seed the device step counter from the host's: steps already taken = self.steps - 1
                self.hyper.copy_(torch.tensor([want[0], want[1], float(max(self.steps - 1, 0)), 0, 0, 0, 0, 0],
                                              dtype=torch.float32), non_blocking=False)
            else:
                self.hyper[:2].copy_(torch.tensor(want, dtype=torch.float32), non_blocking=False)
            self._hyper_host = want

    def set_step_count(self, steps):
        """Restore the optimiser's step count (checkpoint resume)."""
        self.steps = int(steps)
        if getattr(self, "hyper", None) is not None:
            self.steps += 1
            self._sync_hyper(force=True)
            self.steps -= 1

    # ---- the whole step as ONE CUDA-graph launch (launch-bound shards: strong scaling, small per-GPU batches) ----
    def capture(self, example_batch, warmup=3):
        """Capture forward + loss + backward + gradient all-reduce + clip + Adam for batches shaped like
        `example_batch` into a CUDA graph; afterwards `step(batch)` with matching shapes copies the batch into the
        static input buffers and replays.  At 32 clips per GPU (BASELINE config 4's global batch 256 on 8 GPUs) the
        eager step is bound by ~600 Python / ctypes launches, not by the GPU.
        What makes the step capturable: every kernel launches on the current stream and never synchronises; TMA
        descriptors are by-value kernel parameters; the optimiser scalars are device-resident (`hyper`); weight
        re-packing is part of the captured sequence (the derived-weight cache is empty at the start of every step);
        NCCL all-reduce is capturable."""
        if not self.on_gpu:
            raise RuntimeError("graph capture needs the CUDA path")
        if self.want_overlap:
            raise RuntimeError("graph capture and the bucketed all-reduce are alternatives")
        from . import streams
        self.graph = None
        self.static_in = {k: v.detach().clone() for k, v in example_batch.items()}
        nclips = max(int(v.shape[0]) for v in self.static_in.values() if v.dim() > 0)
        prev_overlap = streams.set_train_overlap(
            nclips <= streams.MAX_CLIPS and os.environ.get("M3T_TRAIN_STREAMS", "1") != "0")
        # programmatic dependent launch (csrc/common.cuh): in the launch-bound regime the next kernel's scheduling and
        # prologue overlap the previous kernel's tail (measured, 32 clips: 6.77 -> 6.65 ms per replay); at 256 clips it
        # costs 0.4 ms (early-resident CTAs of ~400 boundaries), so it is on only for the captured small-shard graph
        prev_pdl = L.set_pdl(None)
        if "M3T_PDL" not in os.environ:
            L.set_pdl(nclips <= streams.MAX_CLIPS)
        try:
            return self._capture(warmup)
        finally:
            streams.set_train_overlap(prev_overlap)
            L.set_pdl(prev_pdl)

    def _capture(self, warmup):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1 if self.params is not None else 2)):
                self._eager_step(self.static_in)      # also lays out the arena and allocates hyper
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._sync_hyper()
        from . import raw
        raw.dropout_counter(self.all_params[0].device)      # allocated before the capture starts
        g = torch.cuda.CUDAGraph()
        ops.clear_caches()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            self.static_loss = self._eager_step(self.static_in, count=False)
        self.graph = g
        self._graph_sig = {k: (tuple(v.shape), v.dtype) for k, v in self.static_in.items()}
        return self

    def _graph_matches(self, batch):
        sig = self._graph_sig
        return len(batch) == len(sig) and all(
            k in sig and torch.is_tensor(v) and (tuple(v.shape), v.dtype) == sig[k] for k, v in batch.items())

    def step(self, batch):
        """One optimisation step on this rank's shard; returns the (detached) loss tensor, no host sync."""
        if getattr(self, "graph", None) is not None and self._graph_matches(batch):
            self._sync_hyper()
            for k, v in batch.items():
                if v.data_ptr() != self.static_in[k].data_ptr():
                    self.static_in[k].copy_(v, non_blocking=True)
            self.graph.replay()
            self.steps += 1
            ops.clear_caches()          # the replay rewrote the parameters through the arena
            return self.static_loss
        return self._eager_step(batch)

    def _eager_step(self, batch, count=True):
        ops.DEFER_NUM_BATCHES_TRACKED = True
        if self.on_gpu:
            from . import raw
            raw.begin_step(self.all_params[0].device)     # one memset instead of ~90 zero-fill kernels (raw.StepPool)
            if torch.cuda.is_current_stream_capturing():   # part of the graph: fused dropout draws new masks per replay
                raw.advance_dropout_counter(self.all_params[0].device)
            if self.steps >= 1 and os.environ.get("M3T_PREPACK", "1") == "1":
                if not hasattr(self, "_param_ids"):
                    self._param_ids = {id(p) for p in self.all_params}
                ops.prepack(self._param_ids)              # every conv filter re-packed by one launch
        try:
            return self._eager_step_body(batch, count)
        finally:
            ops.DEFER_NUM_BATCHES_TRACKED = False
            if self.on_gpu:
                raw.end_step(self.all_params[0].device)

    def _eager_step_body(self, batch, count):
        y = self.model(batch)
        if self.nbt and self.model.training:
            torch._foreach_add_(self.nbt, 1)
        loss, _ = self.model.compute_loss(y, batch, sync_free=True)
        if self.overlap:
            self._backward_overlapped(loss)
        else:
            loss.backward()
            self._gather_grads()
            self._allreduce_grads()
        self._optimizer_step(count)
        if self.world > 1 and self.want_overlap and not self.overlap and self.params is not None:
            self._arm_overlap()
        return loss.detach()
