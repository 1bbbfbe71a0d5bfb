"""Throughput of the training step when it is driven by the Lightning-0.6 call surface (`Trainer.fit` ->
`training_step` with the reference's `.item()` syncs -> backward -> engine) next to `TrainEngine.step` (what bench.py
times), same model and batch shape (BASELINE config 4 per GPU: 256 clips x 16 frames).  Not a pytest file; run on a
B200:   python tests/bench_trainer.py [clips] [steps]
Prints one JSON line per arm.  Batches come from pinned host memory through a DataLoader (batch_size=None), so the
Trainer arm includes the host->device copy of every step."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as BN  # noqa: E402
from m3t_b200 import lightning as pl  # noqa: E402
from m3t_b200.engine import TrainEngine  # noqa: E402
from m3t_b200.models.model import AffWild2VA  # noqa: E402


class _Steps(torch.utils.data.Dataset):
    def __init__(self, batches, n):
        self.batches, self.n = batches, n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return self.batches[i % len(self.batches)]


def main():
    clips = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    hp = BN.hparams()
    hp.scheduler, hp.freeze_enc, hp.test_lr = "none", False, False
    batches = [BN.synth_batch(clips, 100 + i, pin=True) for i in range(2)]
    frames = clips * BN.T_FRAMES

    class Timed(AffWild2VA):
        """Records a CUDA event when each batch ends; the first `warm` steps are not timed."""
        def __init__(self, hparams):
            super().__init__(hparams)
            self.events = []

        def on_batch_end(self):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.events.append(e)

        @pl.data_loader
        def train_dataloader(self):
            return torch.utils.data.DataLoader(_Steps(batches, steps + 3), batch_size=None, pin_memory=False)

        @pl.data_loader
        def val_dataloader(self):
            return None

    for prefetch in ("0", "1"):
        os.environ["M3T_TRAINER_PREFETCH"] = prefetch
        torch.manual_seed(12345)
        m = Timed(hp)
        BN.randomise_bn(m, 7)
        tr = pl.Trainer(gradient_clip_val=1.0, max_epochs=1, gpus="0", nb_sanity_val_steps=0,
                        checkpoint_callback=False, early_stop_callback=False, show_progress_bar=False,
                        distributed_backend="dp")
        tr.fit(m)
        torch.cuda.synchronize()
        ms = m.events[2].elapsed_time(m.events[-1]) / (len(m.events) - 3)
        print(json.dumps({"arm": "Trainer.fit (training_step + host syncs + H2D per step), prefetch=" + prefetch,
                          "ms_per_step": ms, "frames_per_s": frames / ms * 1e3, "clips": clips,
                          "steps": len(m.events) - 3}), flush=True)
        del m, tr

    torch.manual_seed(12345)
    m2 = AffWild2VA(hp)
    BN.randomise_bn(m2, 7)
    m2 = m2.cuda().train()
    eng = TrainEngine(m2, lr=hp.learning_rate, weight_decay=1e-4, clip=1.0)
    dev = [{k: v.cuda() for k, v in b.items()} for b in batches]
    for i in range(3):
        eng.step(dev[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        eng.step(dev[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"arm": "TrainEngine.step (resident batch, sync-free)", "ms_per_step": ms,
                      "frames_per_s": frames / ms * 1e3, "clips": clips, "steps": steps}), flush=True)


if __name__ == "__main__":
    main()
