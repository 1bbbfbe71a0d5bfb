"""Secondary measurements for the other BASELINE.json configs (the driver's headline line is bench.py = config 4):
  1  VA_3DResNet visual-only forward, 2 x 16 x 112 x 112            (the reference's CPU-runnable case)
  2  audio ResNet over log-Mel windows + TCN head, forward+backward, batch 64 (builder-declared composition, A2)
  3  full AV inference, batch 32 x T=32: (i) --backbone v2p_split (what model.py runs as-is), (ii) --backbone resnet
  5  long-sequence AV eval, T in {64,128,256}, 16 clips per GPU (= batch 128 on 8 GPUs)
Prints one JSON line per config: frames/s, ms per pass (CUDA events, 3 warm-up + 10 timed, max not needed: 1 GPU)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from m3t_b200 import lib  # noqa: E402


def timed(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lib.launch_count()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (lib.launch_count() - n0) // iters


def hp(**kw):
    d = dict(backbone="resnet", backend="gru", modality="audiovisual", fusion_type="attention", window=32,
             loss="ccc_mtl", loss_lambda=0.5, num_hidden=512, split_layer=5, num_fc_layers=2, learning_rate=5e-5,
             optimizer="adam")
    d.update(kw)
    return argparse.Namespace(**d)


def av_batch(B, T):
    g = torch.Generator().manual_seed(0)
    return {"video": torch.randint(0, 256, (B, 3, T, 112, 112), generator=g, dtype=torch.uint8).float().cuda(),
            "audio": (torch.randn((B, T, 200), generator=g) * 20 - 40).cuda(),
            "se_features": torch.randn((B, 512, T), generator=g).cuda()}


def main():
    from bench import randomise_bn
    from m3t_b200.models.audio_resnet import AudioResNetTCN
    from m3t_b200.models.backbone import VA_3DResNet
    from m3t_b200.models.model import AffWild2VA
    out = []
    # config 1
    torch.manual_seed(12345)
    m = VA_3DResNet(hiddenDim=512, frameLen=16, backend="gru", resnet_ver="v1", nClasses=9, nFCs=2)
    randomise_bn(m, 7)
    m = m.cuda().eval()
    x = (torch.randint(0, 256, (2, 3, 16, 112, 112)).float().cuda() - 127.5) / 127.5
    with torch.no_grad():
        ms, nl = timed(lambda: m(x))
    out.append({"config": 1, "what": "VA_3DResNet eval fwd 2x16", "ms": ms, "frames_per_s": 32 / ms * 1e3, "launches": nl})
    # config 2
    m = AudioResNetTCN(dropout=0.0)
    randomise_bn(m, 7)
    m = m.cuda().train()
    a = (torch.randn(64, 32, 200) * 20 - 40).cuda()

    def step2():
        m.zero_grad(set_to_none=True)
        m(a).square().mean().backward()

    ms, nl = timed(step2)
    out.append({"config": 2, "what": "audio ResNet+TCN fwd+bwd 64x32", "ms": ms, "frames_per_s": 64 * 32 / ms * 1e3,
                "launches": nl})
    # config 3
    for name, h in (("v2p_split", hp(backbone="v2p_split", split_layer=3)), ("resnet", hp())):
        torch.manual_seed(12345)
        m = AffWild2VA(h)
        randomise_bn(m, 7)
        m = m.cuda().eval()
        b = av_batch(32, 32)
        with torch.no_grad():
            ms, nl = timed(lambda: m(b))
        out.append({"config": 3, "what": "AV attention inference 32x32, backbone " + name, "ms": ms,
                    "frames_per_s": 1024 / ms * 1e3, "launches": nl})
    # config 5
    for T in (64, 128, 256):
        torch.manual_seed(12345)
        m = AffWild2VA(hp(window=T))
        randomise_bn(m, 7)
        m = m.cuda().eval()
        b = av_batch(16, T)
        with torch.no_grad():
            ms, nl = timed(lambda: m(b), iters=5)
        out.append({"config": 5, "what": "AV attention eval 16 clips x T=%d (resnet backbone)" % T, "ms": ms,
                    "frames_per_s": 16 * T / ms * 1e3, "launches": nl})
        del m, b
    for o in out:
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
