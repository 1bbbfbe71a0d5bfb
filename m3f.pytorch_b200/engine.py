"""Data-parallel training step of the hot path: forward, loss, backward, gradient all-reduce, clip, Adam
(reference: models/model.py:146-218 `training_step`, :375-390 Adam(lr, weight_decay=1e-4); train.py:35
gradient_clip_val=1.0; Lightning 'ddp' = one process per GPU + NCCL all-reduce of the gradients).

Clips are independent, so ranks only meet in ONE collective per step: the all-reduce (mean) of a flat gradient
buffer over NCCL/NVLink.  BatchNorm statistics and the CCC loss stay per-rank, as in the reference (no SyncBN).
"""
import torch
import torch.distributed as dist


class TrainEngine:
    def __init__(self, model, lr=5e-5, weight_decay=1e-4, clip=1.0, grad_dtype=torch.float32):
        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.clip = clip
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.opt = torch.optim.Adam(self.params, lr=lr, weight_decay=weight_decay,
                                    fused=self.params[0].is_cuda)
        self.grad_dtype = grad_dtype
        self._flat = None

    def _allreduce_grads(self):
        grads = [p.grad for p in self.params if p.grad is not None]
        if self._flat is None or self._flat.numel() != sum(g.numel() for g in grads):
            self._flat = torch.empty(sum(g.numel() for g in grads), device=grads[0].device, dtype=self.grad_dtype)
        flat = self._flat
        torch.cat([g.reshape(-1) for g in grads], out=flat) if flat.dtype == grads[0].dtype else \
            flat.copy_(torch.cat([g.reshape(-1) for g in grads]))
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.mul_(1.0 / self.world)
        off = 0
        views = []
        for g in grads:
            views.append(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        torch._foreach_copy_(grads, views)

    def step(self, batch):
        """One optimisation step on this rank's shard; returns the (detached) loss tensor, no host sync."""
        y = self.model(batch)
        loss, _ = self.model.compute_loss(y, batch, sync_free=True)
        loss.backward()
        if self.world > 1:
            self._allreduce_grads()
        if self.clip:
            torch.nn.utils.clip_grad_norm_(self.params, self.clip, foreach=True)
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        return loss.detach()
