"""ORACLE (test infrastructure) - numpy restatement of the reference's log-Mel extraction
(process/extract_melspec.py:13-20).

PARITY UNPINNED: the arithmetic lives in librosa (requirements.txt:3, unpinned, not installed here, and the
reference ships no vectors for it).  This file restates librosa's published algorithm for the exact call the
reference makes -
    melspectrogram(y, sr=16000, n_fft=512, hop_length=hop, win_length=400, n_mels=40)   [center=True, hann,
    power=2.0, filters.mel(htk=False, norm='slaney', fmin=0, fmax=sr/2)] ;  power_to_db(ref=1.0, amin=1e-10, top_db=80)
- in float64 numpy; `pad_mode` is 'constant' for librosa >= 0.10 and 'reflect' before.  Pinned on the second-best
anchor available: features recorded from torchaudio's MelSpectrogram(norm='slaney', mel_scale='slaney') +
AmplitudeToDB (tests/golden/logmel_torchaudio.pt, oracle/make_golden_logmel.py; agreement < 1e-3 dB), an independent
implementation of the same published algorithm - not librosa itself, hence still "unpinned" in the strict sense.
"""
import math

import numpy as np


def hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) / (math.log(6.4) / 27.0), 3.0 * f / 200.0)


def mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= 15.0, 1000.0 * np.exp((math.log(6.4) / 27.0) * (m - 15.0)), 200.0 * m / 3.0)


def mel_filters(sr=16000, n_fft=512, n_mels=40):
    freqs = np.arange(n_fft // 2 + 1) * (sr / n_fft)
    pts = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2.0), n_mels + 2))
    fb = np.zeros((n_mels, len(freqs)))
    for i in range(n_mels):
        lo, ce, hi = pts[i], pts[i + 1], pts[i + 2]
        up = (freqs - lo) / (ce - lo)
        down = (hi - freqs) / (hi - ce)
        fb[i] = np.maximum(0.0, np.minimum(up, down)) * (2.0 / (hi - lo))
    return fb


def logmel(y, fps, pad_mode="constant", top_db=80.0, sr=16000, n_fft=512, win=400, n_mels=40):
    y = np.asarray(y, dtype=np.float64)
    hop = int(1 / 3 * 1 / fps * sr)
    yp = np.pad(y, n_fft // 2, mode=pad_mode)
    n_frames = 1 + len(y) // hop
    k = np.arange(win)
    w = np.zeros(n_fft)
    w[(n_fft - win) // 2:(n_fft - win) // 2 + win] = 0.5 - 0.5 * np.cos(2 * np.pi * k / win)
    frames = np.stack([yp[i * hop:i * hop + n_fft] for i in range(n_frames)])
    spec = np.abs(np.fft.rfft(frames * w, axis=1)) ** 2
    mel = spec @ mel_filters(sr, n_fft, n_mels).T
    db = 10.0 * np.log10(np.maximum(mel, 1e-10))
    if top_db:
        db = np.maximum(db, db.max() - top_db)
    return db            # (n_frames, n_mels) == spec.transpose() of the reference


def stack_windows(mel, start_idx, w_len):
    """models/dataset.py:83-95."""
    out = np.zeros((w_len, 5 * mel.shape[1]), dtype=mel.dtype)
    for i in range(w_len):
        fr = mel[(start_idx + i) * 3:(start_idx + i) * 3 + 5]
        out[i, :fr.size] = fr.reshape(-1)
    return out
