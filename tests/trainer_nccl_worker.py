"""2-GPU check of the Trainer's 'ddp' mode on NCCL with the real audio-visual model (launched by torchrun through
`m3t_b200.run`, which gives this file the same import surface as the reference's train.py): one epoch on the synthetic
Aff-Wild2 tree, then both ranks must hold bit-identical parameters (one all-reduce of the flat gradient arena + the same
fused Adam step), rank 0's validation_end must have seen the windows of both ranks, and exactly one checkpoint exists.

  torchrun --nproc-per-node 2 --master-addr 127.0.0.1 -m m3t_b200.run tests/trainer_nccl_worker.py <tmp dir>
"""
import glob
import os
import sys
from argparse import ArgumentParser

import torch
import torch.distributed as dist
from pytorch_lightning import Trainer

from models.model import AffWild2VA

if __name__ == "__main__":
    tmp = sys.argv[1]
    rank = int(os.environ["RANK"])
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tests import synth_affwild
    data = os.path.join(tmp, "data")
    if rank == 0:
        synth_affwild.build(data, tmp, input_size=128)
    os.chdir(tmp)
    parser = AffWild2VA.add_model_specific_args(ArgumentParser(add_help=False))
    hp = parser.parse_args(["--modality", "audiovisual", "--backbone", "resnet", "--fusion_type", "attention",
                            "--split_layer", "5", "--window", "8", "--windows_per_epoch", "4", "--batch_size", "2",
                            "--dataset_path", data, "--release", "vipl", "--input_size", "128", "--workers", "0",
                            "--checkpoint_path", tmp, "--device_augment", "--cutout", "--distributed",
                            "--max_nb_epochs", "1"])
    # M3T_TRAINER_BACKEND=gloo: both ranks on GPU 0 with gloo collectives (a one-GPU box) / on the CPU (dry run of the
    # host logic up to the first kernel)
    backend = os.environ.get("M3T_TRAINER_BACKEND", "nccl")
    gpus = ("0,1" if backend == "nccl" else "0,0") if torch.cuda.is_available() else None
    torch.manual_seed(12345)
    model = AffWild2VA(hp)
    seen = []
    inner = model.validation_end
    model.validation_end = lambda outputs: (seen.append(sum(len(o["vid_names"]) for o in outputs)), inner(outputs))[1]
    tr = Trainer(early_stop_callback=None, check_val_every_n_epoch=1, gradient_clip_val=1.0, default_save_path=tmp,
                 max_epochs=1, gpus=gpus, nb_gpu_nodes=1,
                 distributed_backend="ddp", nb_sanity_val_steps=0)
    # rank 0 wrote the tree: nobody may read it before that is done.  Trainer joins torchrun's group in fit(); the
    # barrier here needs it earlier.
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    elif gpus:
        torch.cuda.set_device(0)
    dist.init_process_group(backend)
    dist.barrier()
    tr.fit(model)
    assert tr.world == 2 and tr.engine is not None
    assert tr.device.index == (int(os.environ["LOCAL_RANK"]) if backend == "nccl" else 0)
    flat = torch.cat([p.detach().reshape(-1).float() for p in model.parameters()])
    if backend != "nccl":
        flat = flat.cpu()                          # gloo gathers host tensors only
    both = [torch.empty_like(flat) for _ in range(2)]
    dist.all_gather(both, flat)
    assert torch.equal(both[0], both[1]), "ranks diverged: %g" % float((both[0] - both[1]).abs().max())
    assert torch.isfinite(flat).all()
    n_val_windows = 4 + 3                          # vidC: 27 frames, vidD: 19 frames, window 8 (DistributedSampler pads to 8)
    if rank == 0:
        assert seen and seen[-1] >= n_val_windows, seen
    else:
        assert not seen
    assert "val_loss" in tr.callback_metrics
    dist.barrier()
    assert len(glob.glob(os.path.join(tmp, "lightning_logs", "version_0", "checkpoints", "*.ckpt"))) == 1
    print("TRAINER_NCCL_OK rank %d val_loss %.4f" % (rank, tr.callback_metrics["val_loss"]), flush=True)
