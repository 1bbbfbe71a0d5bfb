"""Worker of test_trainer_ddp_gloo_world2 (launched by torch.distributed.run with 2 ranks, gloo, CPU): the Trainer's
'ddp' mode on a small module — sharded batches, flat gradient all-reduce, gathered validation, rank-0 checkpoint."""
import glob
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from m3t_b200.lightning import Trainer  # noqa: E402
from tests.test_cpu_trainer import _manual_fit, _toy_hparams, _toy_module  # noqa: E402


def main():
    out_dir = sys.argv[1]
    Toy = _toy_module()
    for optimizer in ("adam", "sgd"):
        # each rank sees 4 of every 8 consecutive samples (DistributedSampler without shuffle interleaves ranks), so
        # a global batch of 2 x 4 holds the same samples as the single-process batch of 8: the mean of the two shard
        # gradients of an MSE loss is the full-batch gradient
        hp = _toy_hparams(optimizer=optimizer, scheduler="exp", distributed=True, batch_size=4)
        m = Toy(hp)
        tr = Trainer(gradient_clip_val=0.05, default_save_path=os.path.join(out_dir, optimizer), max_epochs=2,
                     gpus=None, distributed_backend="ddp", nb_sanity_val_steps=0, show_progress_bar=False)
        tr.fit(m)
        assert tr.world == 2 and dist.get_world_size() == 2
        ref = _manual_fit(Toy, _toy_hparams(optimizer=optimizer, scheduler="exp", batch_size=8), 2, 0.05)
        for (k, a), b in zip(m.state_dict().items(), ref.state_dict().values()):
            assert torch.allclose(a, b, atol=3e-6, rtol=1e-5), (optimizer, k, float((a - b).abs().max()))
        # both ranks hold identical parameters
        flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
        both = [torch.empty_like(flat) for _ in range(2)]
        dist.all_gather(both, flat)
        assert torch.equal(both[0], both[1])
        if tr.rank == 0:
            assert m.ended and all(e == list(range(hp.n)) for e in m.ended), "rank 0 must see every rank's outputs"
        else:
            assert not m.ended
        assert "val_loss" in tr.callback_metrics            # broadcast from rank 0
        dist.barrier()
        ck = glob.glob(os.path.join(out_dir, optimizer, "lightning_logs", "version_0", "checkpoints", "*.ckpt"))
        assert len(ck) == 1, ck
    print("TRAINER_DDP_OK rank %d" % dist.get_rank(), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
