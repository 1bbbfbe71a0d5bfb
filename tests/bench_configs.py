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


def postproc_line():
    """overlap-add + Wiener(35) + masked CCC for a validation-set-sized problem (70 videos, 2000-9000
    frames, window 32 at half stride, V/A) on the device, next to scipy.signal.wiener + numpy (what the reference
    calls) on ONE host core for the same tracks."""
    import time

    import numpy as np
    from m3t_b200.process import postproc as PP
    rng = np.random.default_rng(0)
    window, C = 32, 2
    lengths = [int(x) for x in rng.integers(2000, 9000, size=70)]
    segs = []
    for v, n in enumerate(lengths):
        st = 0
        while True:
            ln = min(window, n - st)
            segs.append((v, st, ln))
            if st + ln >= n:
                break
            st += window // 2
    S = len(segs)
    preds = np.tanh(rng.standard_normal((S, window, C)).cumsum(1) * 0.2).astype(np.float32)
    gts = np.clip(preds * 0.8 + rng.standard_normal((S, window, C)) * 0.2, -1, 1).astype(np.float32)
    vids, starts, lens = [s[0] for s in segs], [s[1] for s in segs], [s[2] for s in segs]
    dp, dg = torch.from_numpy(preds).cuda(), torch.from_numpy(gts).cuda()

    def run():
        tp = PP.overlap_add(dp, starts, vids, lens, window, len(lengths))
        tg = PP.overlap_add(dg, starts, vids, lens, window, len(lengths))
        return PP.concordance_cc2_np(PP.smooth_predictions(tp, 35), tg)

    ms, nl = timed(run)
    total = sum(lengths)
    algo_bytes = total * C * (4 + 2 * 8 + 8 + 4 + 8 + 8) + 2 * S * window * C * 4   # wiener in/out + workspaces + ccc reads
    cpu_ms = None
    try:
        from scipy.signal import wiener
        tp = PP.overlap_add(dp, starts, vids, lens, window, len(lengths))
        tg = PP.overlap_add(dg, starts, vids, lens, window, len(lengths))
        tps, tgs = [t.cpu().numpy() for t in tp.split()], [t.cpu().numpy() for t in tg.split()]
        t0 = time.perf_counter()
        allp, allg = [], []
        for p_, g_ in zip(tps, tgs):
            sm = np.stack([wiener(p_[:, c], 35) for c in range(C)], 1)
            valid = np.all(g_ >= -1, axis=1)
            allp.append(sm[valid])
            allg.append(g_[valid])
        P_, G_ = np.concatenate(allp), np.concatenate(allg)
        for c in range(C):
            mcp = ((P_[:, c] - P_[:, c].mean()) * (G_[:, c] - G_[:, c].mean())).mean()
            _ = 2 * mcp / (P_[:, c].var() + G_[:, c].var() + (P_[:, c].mean() - G_[:, c].mean()) ** 2)
        cpu_ms = (time.perf_counter() - t0) * 1e3
    except ImportError:
        pass
    return {"config": "N3", "what": "eval post-processing: overlap-add + Wiener(35) + masked CCC, 70 videos / %d frames"
            % total, "ms": ms, "frames_per_s": total / ms * 1e3, "launches": nl,
            "achieved_GBps": algo_bytes / ms / 1e6, "cpu_scipy_numpy_ms_1core_smooth_and_ccc_only": cpu_ms}


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
            # per-class breakdown of one forward (temporal / fusion kernels reported separately)
            from m3t_b200 import raw
            prof = raw.KernelProfiler()
            raw.set_profiler(prof)
            m(b)
            torch.cuda.synchronize()
            raw.set_profiler(None)
        cls = {"conv": 0.0, "gemm": 0.0, "gru_fwd": 0.0, "att_mix": 0.0}
        for k, v in prof.summary().items():
            key = "gru_fwd" if k.startswith("gru_fwd") else "att_mix" if k.startswith("att_mix") else \
                "gemm" if k.startswith("gemm") else "conv"
            cls[key] += v["ms_total"]
        out.append({"config": 5, "what": "AV attention eval 16 clips x T=%d (resnet backbone)" % T, "ms": ms,
                    "frames_per_s": 16 * T / ms * 1e3, "launches": nl,
                    "ms_by_class": {k: round(v, 3) for k, v in cls.items()},
                    "gru_us_per_step": round(cls["gru_fwd"] * 1e3 / (8 * T), 2)})
        del m, b
    # CUDA-graph replay of the launch-bound small-batch forwards (m3t_b200.graphs)
    from m3t_b200.graphs import GraphedInference
    torch.manual_seed(12345)
    m = VA_3DResNet(hiddenDim=512, frameLen=16, backend='gru', resnet_ver='v1', nClasses=9, nFCs=2)
    randomise_bn(m, 7)
    m = m.cuda().eval()
    x = (torch.randint(0, 256, (2, 3, 16, 112, 112)).float().cuda() - 127.5) / 127.5
    g = GraphedInference(m, x)
    ms, nl = timed(lambda: g(x))
    out.append({"config": 1, "what": "VA_3DResNet eval fwd 2x16, CUDA-graph replay", "ms": ms,
                "frames_per_s": 32 / ms * 1e3, "launches": nl})
    torch.manual_seed(12345)
    m = AffWild2VA(hp())
    randomise_bn(m, 7)
    m = m.cuda().eval()
    b = av_batch(32, 32)
    g = GraphedInference(m, b)
    ms, nl = timed(lambda: g(b))
    out.append({"config": 3, "what": "AV attention inference 32x32, backbone resnet, CUDA-graph replay", "ms": ms,
                "frames_per_s": 1024 / ms * 1e3, "launches": nl})
    del m, b, g
    # fp32-parity mode (m3t_b200.fp32): configs 1 and 3 (resnet backbone) again with float32 activations and
    # split-operand tensor-core launches
    from m3t_b200 import fp32
    torch.manual_seed(12345)
    m = VA_3DResNet(hiddenDim=512, frameLen=16, backend='gru', resnet_ver='v1', nClasses=9, nFCs=2)
    randomise_bn(m, 7)
    m = m.cuda().eval()
    x = (torch.randint(0, 256, (2, 3, 16, 112, 112)).float().cuda() - 127.5) / 127.5
    with torch.no_grad(), fp32.parity_mode():
        ms, nl = timed(lambda: m(x))
    out.append({"config": 1, "what": "VA_3DResNet eval fwd 2x16, fp32-parity mode", "ms": ms,
                "frames_per_s": 32 / ms * 1e3, "launches": nl})
    torch.manual_seed(12345)
    m = AffWild2VA(hp())
    randomise_bn(m, 7)
    m = m.cuda().eval()
    b = av_batch(32, 32)
    with torch.no_grad(), fp32.parity_mode():
        ms, nl = timed(lambda: m(b))
    out.append({"config": 3, "what": "AV attention inference 32x32, backbone resnet, fp32-parity mode", "ms": ms,
                "frames_per_s": 1024 / ms * 1e3, "launches": nl})
    del m, b
    out.append(postproc_line())
    for o in out:
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
