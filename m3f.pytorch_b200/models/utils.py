"""Metrics and post-processing helpers of the task module (reference: models/utils.py:6-45), same names and call
conventions, so `get_smoothed_ccc.py` / `create_submission.py` run unchanged through `m3t_b200.run`."""
import numpy as np


def concordance_cc2(r1, r2, reduction='mean'):
    """Batch sequence-wise CCC: biased covariance over unbiased variances (reference semantics)."""
    m1 = r1.mean(dim=-1, keepdim=True)
    m2 = r2.mean(dim=-1, keepdim=True)
    cov = ((r1 - m1) * (r2 - m2)).mean(dim=-1, keepdim=True)
    ccc = 2 * cov / (r1.var(dim=-1, keepdim=True) + r2.var(dim=-1, keepdim=True) + (m1 - m2) ** 2)
    if reduction == 'none':
        return ccc
    return ccc.mean()


def concordance_cc2_np(r1, r2):
    cov = ((r1 - r1.mean()) * (r2 - r2.mean())).mean()
    return 2 * cov / (r1.var() + r2.var() + (r1.mean() - r2.mean()) ** 2)


def mse(preds, labels):
    return sum((preds - labels) ** 2) / len(labels)


def smooth_predictions(preds, window=13, mode='wiener'):
    """Wiener smoothing of one prediction track along axis 0 (reference :29-33 -> scipy.signal.wiener): float32
    ndarray / tensor (T,) or (T, C) in, float64 ndarray out, computed by `m3t_wiener1d_f64` on the current CUDA device
    (no host fallback).  For whole validation sets use process.postproc.smooth_predictions, which filters every video
    in one launch; this wrapper keeps the reference's one-track-per-call convention."""
    import torch

    from ..process import postproc
    if mode != 'wiener':
        raise NotImplementedError("mode=%r: the reference's scripts only smooth with mode='wiener'" % (mode,))
    x = preds.detach().cpu() if torch.is_tensor(preds) else torch.from_numpy(np.ascontiguousarray(preds))
    if x.dtype != torch.float32:
        raise NotImplementedError("smooth_predictions takes float32 tracks (got %s)" % x.dtype)
    if not torch.cuda.is_available():
        raise RuntimeError("smooth_predictions runs on the device (m3t_wiener1d_f64); no CUDA device is visible")
    flat = x.reshape(x.shape[0], -1).cuda()
    if flat.shape[0] == 0:
        return np.zeros(tuple(x.shape), dtype=np.float64)
    seq_off = torch.tensor([0, flat.shape[0]], dtype=torch.int64, device=flat.device)
    out = postproc.smooth_predictions(postproc.Tracks(flat.contiguous(), seq_off, [flat.shape[0]]), window, mode)
    return out.data.cpu().numpy().reshape(tuple(x.shape))


def plot_results(base_path, y1, y2, index):
    """Ground truth against prediction over the frames of one video (reference :36-45)."""
    from matplotlib import pyplot as plt
    frames = np.arange(len(y1))
    plt.plot(frames, y1, label="Actual " + index)
    plt.plot(frames, y2, label="Predicted " + index)
    plt.xlabel('Frames')
    plt.ylabel(index)
    plt.title("Aff-Wild2 predictions")
    plt.legend()
    plt.show()
