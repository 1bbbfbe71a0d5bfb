"""Started as ONE plain process (`python relaunch_script.py <out dir>`): Trainer(distributed_backend='ddp',
num_processes=2) re-launches this command line as rank 1 and both ranks train the small CPU module of
tests/test_cpu_trainer.py over gloo; every rank leaves a checksum of its parameters."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from m3t_b200.lightning import Trainer  # noqa: E402
from tests.test_cpu_trainer import _toy_hparams, _toy_module  # noqa: E402

if __name__ == "__main__":
    out = sys.argv[1]
    with open(os.path.join(out, "started_%d" % os.getpid()), "w") as f:
        f.write(os.environ.get("RANK", "none"))
    model = _toy_module()(_toy_hparams(distributed=True, batch_size=4))
    tr = Trainer(gradient_clip_val=0.05, default_save_path=out, max_epochs=2, gpus=None, distributed_backend="ddp",
                 num_processes=2, nb_sanity_val_steps=0, show_progress_bar=False)
    tr.fit(model)
    tr.test(model)                 # a second entry must not launch ranks again
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    with open(os.path.join(out, "rank_%d" % tr.rank), "w") as f:
        f.write("%d %.10f" % (tr.world, float(flat.double().sum())))
