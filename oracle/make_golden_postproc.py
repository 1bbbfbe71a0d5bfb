"""Generate tests/golden/postproc.pt by running the UNMODIFIED reference code on seeded synthetic window predictions:
  * overlap-add: AffWild2VA.validation_end (models/model.py:247-310, test_on_val branch) called with a stub `self`
    inside a temporary directory (it torch.save()s 'predictions_val.pt' into the cwd, which is read back);
  * smoothing + CCC: models/utils.py smooth_predictions (-> scipy.signal.wiener) and concordance_cc2_np, driven exactly
    as get_smoothed_ccc.py:10-30 does.
Run on the build box only:  python -m oracle.make_golden_postproc
"""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import _refload  # noqa: E402


def main():
    torch.manual_seed(7)
    window, C = 32, 2
    lengths = [150, 97, 64, 233]                 # frames per video (last windows are shorter)
    segs = []                                    # (video, start, n)
    for v, n in enumerate(lengths):
        st = 0
        while True:
            ln = min(window, n - st)
            segs.append((v, st, ln))
            if st + ln >= n:
                break
            st += window // 2
    S = len(segs)
    preds = torch.tanh(torch.randn(S, window, C) * 0.7)
    gts = (preds * 0.7 + torch.randn(S, window, C) * 0.25).clamp(-1, 1)   # correlated with the predictions
    # a few invalid annotations (-5) as in Aff-Wild2
    gts[1, 3:9, 0] = -5.0
    gts[6, 10:12, 1] = -5.0
    perm = torch.randperm(S, generator=torch.Generator().manual_seed(3)).tolist()     # batches arrive unordered
    model_mod = _refload.load("model")
    utils_mod = _refload.load("utils")
    # ---- reference overlap-add through validation_end ----
    outputs, bs = [], 5
    for i in range(0, S, bs):
        idx = perm[i:i + bs]
        outputs.append({
            "v_gt": [gts[s, :segs[s][2], 0] for s in idx], "a_gt": [gts[s, :segs[s][2], 1] for s in idx],
            "v_pred": [preds[s, :segs[s][2], 0] for s in idx], "a_pred": [preds[s, :segs[s][2], 1] for s in idx],
            "vid_names": ["vid%02d" % segs[s][0] for s in idx],
            "start_frames": torch.tensor([segs[s][1] for s in idx]),
        })
    stub = argparse.Namespace(hparams=argparse.Namespace(test_on_val=True, window=window))
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        try:
            model_mod.AffWild2VA.validation_end(stub, outputs)
            saved = torch.load("predictions_val.pt")
        finally:
            os.chdir(cwd)
    names = ["vid%02d" % v for v in range(len(lengths))]
    track_pred = [torch.stack([saved["valence_pred"][k], saved["arousal_pred"][k]], 1) for k in names]
    track_gt = [torch.stack([saved["valence_gt"][k], saved["arousal_gt"][k]], 1) for k in names]
    # ---- reference smoothing + CCC, as get_smoothed_ccc.py drives them ----
    smooth, per_video, all_p, all_g = [], [], [], []
    for p, g in zip(track_pred, track_gt):
        # numpy inputs: with the numpy of the reference's era np.apply_along_axis returned an ndarray for a tensor
        # argument, so concordance_cc2_np saw ndarrays (biased .var()).  numpy >= 2 wraps the result back into a
        # torch.Tensor, whose .var() is unbiased - an environment artefact, not the reference's arithmetic.
        pv = utils_mod.smooth_predictions(p[:, 0].numpy(), 35, mode="wiener")
        pa = utils_mod.smooth_predictions(p[:, 1].numpy(), 35, mode="wiener")
        gv, ga = g[:, 0].numpy(), g[:, 1].numpy()
        valid = (gv >= -1) & (ga >= -1)
        per_video.append([utils_mod.concordance_cc2_np(pv[valid], gv[valid]),
                          utils_mod.concordance_cc2_np(pa[valid], ga[valid])])
        smooth.append(torch.from_numpy(np.stack([pv, pa], 1)))
        all_p.append(np.stack([pv[valid], pa[valid]], 1))
        all_g.append(np.stack([gv[valid], ga[valid]], 1))
    P, G = np.concatenate(all_p), np.concatenate(all_g)
    overall = [utils_mod.concordance_cc2_np(P[:, c], G[:, c]) for c in range(C)]
    fx = {"window": window, "segs": segs, "preds": preds, "gts": gts, "lengths": lengths,
          "track_pred": track_pred, "track_gt": track_gt, "smooth": smooth,
          "ccc_per_video": torch.tensor(per_video, dtype=torch.float64),
          "ccc_overall": torch.tensor(overall, dtype=torch.float64)}
    out = os.path.join(ROOT, "tests", "golden", "postproc.pt")
    torch.save(fx, out)
    print("wrote", out, os.path.getsize(out), "bytes; overall CCC", overall)


if __name__ == "__main__":
    main()
