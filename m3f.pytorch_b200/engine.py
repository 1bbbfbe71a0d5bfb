"""Data-parallel training step of the hot path: forward, loss, backward, gradient all-reduce, clip, Adam
(reference: models/model.py:146-218 `training_step`, :375-390 Adam(lr, weight_decay=1e-4); train.py:35
gradient_clip_val=1.0; Lightning 'ddp' = one process per GPU + NCCL all-reduce of the gradients).

Clips are independent, so ranks only meet in ONE collective per step: the all-reduce (sum) of the flat gradient
arena over NCCL/NVLink; the 1/world mean is folded into the optimiser kernel.  BatchNorm statistics and the CCC
loss stay per-rank, as in the reference (no SyncBN).

Parameters and gradients live in two flat fp32 arenas (every `p.data` / `p.grad` is a view), so the collective, the
global-norm clip and Adam each are ONE call over one buffer (optim.cu), with no host synchronisation.
"""
import torch
import torch.distributed as dist

from . import lib as L
from . import ops


class TrainEngine:
    def __init__(self, model, lr=5e-5, weight_decay=1e-4, clip=1.0, betas=(0.9, 0.999), eps=1e-8):
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

    def _build_arena(self):
        """Lay out the flat arenas over the parameters that actually receive a gradient.  Parameters autograd never
        reaches (e.g. `resnet.fc`, present in the state_dict but unused with agg_mode 'ap', models/resnet.py:72,117)
        stay outside, exactly as torch.optim.Adam skips parameters whose .grad is None (no weight decay on them)."""
        self.params = [p for p in self.all_params if p.grad is not None]
        dev = self.params[0].device
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4       # 16-byte aligned slots
        self.n = n
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
            self.gnorm_sq = torch.zeros(1, device=dev, dtype=torch.float32)
        else:  # host-side logic tests only (gloo); the product path is the CUDA one
            self.opt = torch.optim.Adam(self.params, lr=self.lr, weight_decay=self.wd, betas=self.betas,
                                        eps=self.eps)

    def _gather_grads(self):
        """autograd hands every parameter a fresh gradient tensor (p.grad is None before backward, so nothing is
        accumulated); one multi-tensor copy moves them into the flat arena."""
        if self.params is None:
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

    def _allreduce_grads(self):
        if self.world > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)
            if not self.on_gpu:
                self.flat_g.mul_(1.0 / self.world)

    def _optimizer_step(self):
        self.steps += 1
        if not self.on_gpu:
            for p, v in zip(self.params, self.grad_views):
                p.grad = v
            if self.clip:
                torch.nn.utils.clip_grad_norm_(self.params, self.clip)
            self.opt.step()
            for p in self.params:
                p.grad = None
            return
        lib = L.load()
        st = L.stream_ptr()
        if self.clip:
            L.check(lib.m3t_sumsq_f32(L.ptr(self.flat_g), L.i64(self.n), L.ptr(self.gnorm_sq), st), "sumsq")
        L.check(lib.m3t_adam_clip_step(L.ptr(self.flat_p), L.ptr(self.flat_g), L.ptr(self.m), L.ptr(self.v),
                                       L.i64(self.n), L.f32(self.lr), L.f32(self.betas[0]), L.f32(self.betas[1]),
                                       L.f32(self.eps), L.f32(self.wd), L.i32(self.steps), L.f32(self.clip),
                                       L.f32(1.0 / self.world), L.ptr(self.gnorm_sq), st), "adam_clip_step")

    def step(self, batch):
        """One optimisation step on this rank's shard; returns the (detached) loss tensor, no host sync."""
        ops.DEFER_NUM_BATCHES_TRACKED = True
        try:
            y = self.model(batch)
        finally:
            ops.DEFER_NUM_BATCHES_TRACKED = False
        if self.nbt and self.model.training:
            torch._foreach_add_(self.nbt, 1)
        loss, _ = self.model.compute_loss(y, batch, sync_free=True)
        loss.backward()
        self._gather_grads()
        self._allreduce_grads()
        self._optimizer_step()
        return loss.detach()
