"""Evaluation post-processing on the device: overlap-add of half-stride windows, Wiener smoothing, masked CCC.

Mirrors the reference's host-side steps after the model (models/model.py:281-297,358-366; models/utils.py:19-33;
get_smoothed_ccc.py) with the same function names where the reference has one, on ragged per-video tracks kept in
ONE flat device tensor.  All arithmetic runs in csrc/postproc.cu through the C ABI; there is no host fallback.
"""
import numpy as np
import torch

from .. import lib as L


def _lib():
    return L.load()


class Tracks:
    """Per-video frame tracks as one flat tensor: data [total_frames, C], seq_off int64 [V+1] (device), lengths (host)."""

    def __init__(self, data, seq_off, lengths):
        self.data, self.seq_off, self.lengths = data, seq_off, list(lengths)

    def split(self):
        return list(torch.split(self.data, self.lengths))


def overlap_add(preds, starts, vid_of_seg, seg_lens, window, n_videos):
    """preds: float32 [S, L, C] window predictions (device); starts / vid_of_seg / seg_lens: per-segment start frame,
    video index and number of valid frames (host sequences).  Returns Tracks (float32).
    (models/model.py:281-297: nframes = start of the last window + its length; frames >= window//2 are halved.)"""
    assert preds.is_cuda and preds.dtype == torch.float32 and preds.dim() == 3
    S, Lw, C = preds.shape
    # per-segment bookkeeping is vectorised numpy on the host (tens of thousands of windows per validation set)
    starts = np.asarray(starts, dtype=np.int64)
    vids = np.asarray(vid_of_seg, dtype=np.int64)
    lens = np.asarray(seg_lens, dtype=np.int64)
    lengths = np.zeros(n_videos, dtype=np.int64)
    np.maximum.at(lengths, vids, starts + lens)
    off = np.concatenate([[0], np.cumsum(lengths)])
    dev = preds.device
    host = torch.from_numpy(np.concatenate([off, off[vids]]))                      # one H2D copy for both int64 arrays
    dev64 = host.to(dev)
    seq_off, seg_base = dev64[:n_videos + 1], dev64[n_videos + 1:]
    dev32 = torch.from_numpy(np.concatenate([starts, lens]).astype(np.int32)).to(dev)
    seg_start, seg_len = dev32[:S], dev32[S:]
    lengths = [int(x) for x in lengths]
    off = [int(x) for x in off]
    out = torch.empty((off[-1], C), dtype=torch.float32, device=dev)
    L.check(_lib().m3t_overlap_add_f32(L.ptr(preds.contiguous()), L.ptr(seg_start), L.ptr(seg_len), L.ptr(seg_base),
                                       L.ptr(seq_off), L.ptr(out), L.i64(S), L.i32(Lw), L.i32(C), L.i32(n_videos),
                                       L.i64(off[-1]), L.i32(window), L.stream_ptr()), "m3t_overlap_add_f32")
    return Tracks(out, seq_off, lengths)


def smooth_predictions(tracks, window=13, mode="wiener"):
    """models/utils.py:29-33 on every video of `tracks` (Tracks, float32) at once; returns Tracks (float64)."""
    if mode != "wiener":
        raise NotImplementedError("only mode='wiener' is used by the reference's evaluation (get_smoothed_ccc.py:15)")
    x = tracks.data
    assert x.is_cuda and x.dtype == torch.float32
    F_, C = x.shape
    V = len(tracks.lengths)
    lmean = torch.empty((F_, C), dtype=torch.float64, device=x.device)
    lvar = torch.empty_like(lmean)
    out = torch.empty_like(lmean)
    noise = torch.empty((V, C), dtype=torch.float64, device=x.device)
    L.check(_lib().m3t_wiener1d_f64(L.ptr(x), L.ptr(tracks.seq_off), L.i32(V), L.i32(C), L.i32(window),
                                    L.i64(max(tracks.lengths)), L.ptr(lmean), L.ptr(lvar), L.ptr(noise), L.ptr(out),
                                    L.stream_ptr()), "m3t_wiener1d_f64")
    return Tracks(out, tracks.seq_off, tracks.lengths)


def _ccc_from_moments(m):
    n, sa, sb, saa, sbb, sab = m.unbind(-1)
    ma, mb = sa / n, sb / n
    cov = sab / n - ma * mb
    va, vb = saa / n - ma * ma, sbb / n - mb * mb
    return 2 * cov / (va + vb + (ma - mb) ** 2)


def concordance_cc2_np(pred_tracks, gt_tracks):
    """Masked CCC of models/utils.py:19-21 / get_smoothed_ccc.py:17-30 for every video and channel and over all videos.
    pred_tracks: Tracks float64; gt_tracks: Tracks float32 (frames with any ground-truth channel < -1 are skipped).
    Returns (per_video [V, C], overall [C]) float64 device tensors."""
    p, g = pred_tracks.data, gt_tracks.data
    assert p.dtype == torch.float64 and g.dtype == torch.float32 and p.shape == g.shape
    V, C = len(pred_tracks.lengths), p.shape[1]
    mom = torch.empty((V, C, 6), dtype=torch.float64, device=p.device)
    L.check(_lib().m3t_ccc_moments_f64(L.ptr(p), L.ptr(g), L.ptr(pred_tracks.seq_off), L.i32(V), L.i32(C), L.ptr(mom),
                                       L.stream_ptr()), "m3t_ccc_moments_f64")
    return _ccc_from_moments(mom), _ccc_from_moments(mom.sum(0))
