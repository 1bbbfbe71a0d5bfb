"""ORACLE (test infrastructure, never imported by the product): CPU restatement of the reference's evaluation
post-processing, numpy float64 where the reference's numpy / scipy calls compute in float64.

  overlap_add      models/model.py:281-297 (validation_end, test_on_val) and :358-366 (test_end)
  wiener           scipy.signal.wiener as called by models/utils.py:29-33 smooth_predictions(mode='wiener') from
                   get_smoothed_ccc.py:15-16 (window 35).  scipy is a third-party dependency of the reference
                   (requirements.txt, unpinned); the restatement follows scipy 1.18.1's source and is pinned against
                   scipy.signal.wiener itself and the reference's smooth_predictions (tests/golden/postproc.pt).
  ccc              models/utils.py:19-21 concordance_cc2_np + the validity mask of get_smoothed_ccc.py:21-26
"""
import numpy as np


def overlap_add(preds, starts, vid_of_seg, seg_lens, window, n_videos):
    """preds [S][L][C] float32 -> list over videos of [nframes][C] float32 tracks."""
    out = []
    for v in range(n_videos):
        segs = sorted((int(starts[s]), s) for s in range(len(starts)) if vid_of_seg[s] == v)
        last = segs[-1][1]
        nframes = int(starts[last]) + int(seg_lens[last])
        track = np.zeros((nframes, preds.shape[2]), dtype=np.float32)
        for st, s in segs:
            n = int(seg_lens[s])
            track[st:st + n] += preds[s, :n]
        track[window // 2:] /= np.float32(2.0)
        out.append(track)
    return out


def wiener(x, window):
    """x: 1-D float32.  Returns float64 (scipy promotes through its float64 ones() kernel; x**2 is taken in float32)."""
    x = np.asarray(x)
    n, half = len(x), window // 2
    x64 = x.astype(np.float64)
    sq64 = (x ** 2).astype(np.float64)
    lmean = np.empty(n)
    lvar = np.empty(n)
    for i in range(n):
        a, b = max(0, i - half), min(n, i + half + 1)
        lmean[i] = x64[a:b].sum() / float(window)
        lvar[i] = sq64[a:b].sum() / float(window) - lmean[i] ** 2
    noise = lvar.mean()
    with np.errstate(divide="ignore", invalid="ignore"):
        res = (x64 - lmean) * (1 - noise / lvar) + lmean
    return np.where(lvar < noise, lmean, res)


def ccc(r1, r2):
    """As the reference evaluates it: r1 = smoothed prediction (float64), r2 = ground truth (float32), so r2's mean
    and (biased) variance are float32 numpy reductions; the CUDA path takes them in float64 (difference ~1e-8)."""
    mcp = ((r1 - r1.mean()) * (r2 - r2.mean())).mean()
    return (2 * mcp) / (r1.var() + r2.var() + (r1.mean() - r2.mean()) ** 2)


def smoothed_ccc(tracks_pred, tracks_gt, window=35):
    """get_smoothed_ccc.py main: per-video and global CCC of Wiener-smoothed predictions; channel 0 = V, 1 = A."""
    per_video, all_p, all_g = [], [], []
    for p, g in zip(tracks_pred, tracks_gt):
        sm = np.stack([wiener(p[:, c], window) for c in range(p.shape[1])], axis=1)
        valid = np.all(g >= -1, axis=1)
        per_video.append([ccc(sm[valid, c], g[valid, c]) for c in range(p.shape[1])])
        all_p.append(sm[valid])
        all_g.append(g[valid])
    P, G = np.concatenate(all_p), np.concatenate(all_g)
    return np.array(per_video), np.array([ccc(P[:, c], G[:, c]) for c in range(P.shape[1])])
