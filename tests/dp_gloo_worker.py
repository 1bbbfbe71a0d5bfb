"""Worker for test_data_parallel_gradient_allreduce_gloo: exercises TrainEngine's flat-gradient all-reduce on a small
CPU model (the reduction logic is device-agnostic; NCCL replaces gloo on the GPU box)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from m3t_b200.engine import TrainEngine  # noqa: E402


class Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(8, 16)
        self.b = torch.nn.Linear(16, 3)

    def forward(self, batch):
        return self.b(torch.tanh(self.a(batch["x"])))

    def compute_loss(self, y, batch, sync_free=False):
        return ((y - batch["t"]) ** 2).mean(), {}


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.manual_seed(0)
    m = Tiny()
    ref = Tiny()
    ref.load_state_dict(m.state_dict())
    eng = TrainEngine(m, lr=0.0, weight_decay=0.0, clip=None)
    g = torch.Generator().manual_seed(100 + rank)
    batch = {"x": torch.randn(5, 8, generator=g), "t": torch.randn(5, 3, generator=g)}
    # expected: mean over ranks of the per-rank gradients
    exp = None
    for r in range(world):
        gr = torch.Generator().manual_seed(100 + r)
        b = {"x": torch.randn(5, 8, generator=gr), "t": torch.randn(5, 3, generator=gr)}
        ref.zero_grad()
        ref.compute_loss(ref(b), b)[0].backward()
        gs = [p.grad.clone() for p in ref.parameters()]
        exp = gs if exp is None else [a + c for a, c in zip(exp, gs)]
    exp = [e / world for e in exp]
    y = m(batch)
    loss, _ = m.compute_loss(y, batch)
    loss.backward()
    eng._gather_grads()
    eng._allreduce_grads()
    for v, e in zip(eng.grad_views, exp):
        assert torch.allclose(v, e, atol=1e-6), (rank, (v - e).abs().max())
    # the overlapped path (second step on): bucketed async all-reduce issued from post-accumulate hooks; with lr = 0 the
    # parameters do not move, so every step must reproduce the same mean gradient in the arena
    eng.num_buckets = 2
    eng.want_overlap = True
    for _ in range(3):
        eng.step(batch)
        assert eng.overlap and len(eng.buckets) == 2
        for v, e in zip(eng.grad_views, exp):
            assert torch.allclose(v, e, atol=1e-6), (rank, "overlap", (v - e).abs().max())
    print("DP_OK rank %d" % rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
