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


class TrainEngine:
    def __init__(self, model, lr=5e-5, weight_decay=1e-4, clip=1.0, betas=(0.9, 0.999), eps=1e-8):
        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.clip = float(clip) if clip else 0.0
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.steps = 0
        dev = self.params[0].device
        self.on_gpu = dev.type == "cuda"
        # 16-byte aligned slots so every view is vector-load friendly
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.n = n
        self.flat_p = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        for p, o in zip(self.params, offs):
            self.flat_p[o:o + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.flat_p[o:o + p.numel()].view_as(p)
            p.grad = self.flat_g[o:o + p.numel()].view_as(p)
        if self.on_gpu:
            self.m = torch.zeros(n, device=dev, dtype=torch.float32)
            self.v = torch.zeros(n, device=dev, dtype=torch.float32)
            self.gnorm_sq = torch.zeros(1, device=dev, dtype=torch.float32)
        else:  # host-side logic tests only (gloo); the product path is the CUDA one
            self.opt = torch.optim.Adam(self.params, lr=lr, weight_decay=weight_decay, betas=betas, eps=eps)

    def _allreduce_grads(self):
        if self.world > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)
            if not self.on_gpu:
                self.flat_g.mul_(1.0 / self.world)

    def _optimizer_step(self):
        self.steps += 1
        if not self.on_gpu:
            if self.clip:
                torch.nn.utils.clip_grad_norm_(self.params, self.clip)
            self.opt.step()
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
        y = self.model(batch)
        loss, _ = self.model.compute_loss(y, batch, sync_free=True)
        loss.backward()
        self._allreduce_grads()
        self._optimizer_step()
        self.flat_g.zero_()
        return loss.detach()
