"""ncu target for the launch-bound regime's own kernels (one rank's 32-clip share of BASELINE config 4): the cluster
GRU forward / BPTT kernels at the model's three hidden sizes and the one-launch weight re-pack of the AV model.
   ncu --set full --clock-control none --import-source on -k regex:"gru_|pack_filters" -o gpurun_out/prof_rec \
       python tests/prof_recurrence.py [clips]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from m3t_b200 import ops, raw  # noqa: E402
from m3t_b200.engine import TrainEngine  # noqa: E402
from m3t_b200.models.model import AffWild2VA  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    T, dev = 16, "cuda"
    g = torch.Generator().manual_seed(0)
    for H in (512, 256, 128):
        gi = (torch.randn((B * T, 6 * H), generator=g) * 0.8).to(dev)
        w = (torch.randn((2, 3 * H, H), generator=g) / H ** 0.5).bfloat16().to(dev)
        bh = (torch.randn((2, 3 * H), generator=g) * 0.1).to(dev)
        out, _, saved = raw.gru_fwd(gi, w, bh, B, T, H, True, cluster=B <= 64)
        dout = torch.randn((B, T, 2 * H), generator=g).bfloat16().to(dev)
        raw.gru_bwd(dout, out, saved, w.transpose(1, 2).contiguous(), B, T, H, cluster=True)
        raw.gru_bwd(dout, out, saved, w.transpose(1, 2).contiguous(), B, T, H, cluster=False)
    # the AV model's re-pack: two engine steps (the second one starts with ops.prepack)
    torch.manual_seed(0)
    m = AffWild2VA(bench.hparams()).cuda().train()
    eng = TrainEngine(m, lr=1e-4, weight_decay=1e-4, clip=1.0)
    batch = {k: v.cuda() for k, v in bench.synth_batch(4, 1, pin=False).items()}
    for _ in range(3):
        eng.step(batch)
    torch.cuda.synchronize()
    print("done", ops.CACHE_GENERATION)


if __name__ == "__main__":
    main()
