"""GPU log-Mel extraction (reference: process/extract_melspec.py:8-25).

`extract_melspec((fps, src_wav, dst_npy))` keeps the reference's task signature; the arithmetic
(librosa.feature.melspectrogram(n_fft=512, hop_length=int(16000/(3*fps)), win_length=400, n_mels=40) +
librosa.power_to_db) runs as one fused CUDA kernel per STFT frame.  The input must already be 16 kHz mono: the
resampling inside `librosa.load(sr=16000)` is outside the hot path.
"""
import math
import os

import numpy as np
import torch

from .. import lib as L

SR = 16000
N_FFT = 512
WIN = 400
N_MELS = 40


def _hz_to_mel(f):
    """Slaney mel scale (librosa default, htk=False)."""
    f = np.asarray(f, dtype=np.float64)
    lin = f / (200.0 / 3)
    logstep = math.log(6.4) / 27.0
    return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) / logstep, lin)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    logstep = math.log(6.4) / 27.0
    return np.where(m >= 15.0, 1000.0 * np.exp(logstep * (m - 15.0)), m * (200.0 / 3))


def mel_filterbank(sr=SR, n_fft=N_FFT, n_mels=N_MELS, fmin=0.0, fmax=None):
    """librosa.filters.mel(..., htk=False, norm='slaney'): triangular filters, area-normalised."""
    fmax = fmax if fmax is not None else sr / 2.0
    fftfreqs = np.linspace(0, sr / 2.0, n_fft // 2 + 1)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, n_fft // 2 + 1))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


_fb_cache = {}


# librosa.stft's centre padding changed from 'reflect' (< 0.10, the version the reference script ran on and the released
# checkpoints' mel_spec/*.npy were made with: requirements.txt pins nothing, the script dates from early 2020 =
# librosa 0.7) to 'constant' (>= 0.10).  'reflect' is therefore the default; only the first / last frames differ.
DEFAULT_PAD_MODE = "reflect"


def melspectrogram_db(y, fps, pad_mode=DEFAULT_PAD_MODE, top_db=80.0):
    """y: 1-D float32 CUDA tensor (16 kHz).  Returns (n_frames, 40) float32 dB log-Mel, n_frames = 1 + len(y)//hop.
    pad_mode: 'reflect' (librosa < 0.10, the reference's era) or 'constant' (librosa >= 0.10)."""
    if pad_mode not in ("reflect", "constant"):
        raise ValueError("pad_mode must be 'reflect' or 'constant', got %r" % (pad_mode,))
    assert y.is_cuda and y.dtype == torch.float32 and y.dim() == 1
    y = y.contiguous()
    hop = int(1 / 3 * 1 / fps * SR)
    key = str(y.device)
    if key not in _fb_cache:
        _fb_cache[key] = torch.from_numpy(mel_filterbank()).to(y.device)
    n_frames = 1 + y.numel() // hop
    out = torch.empty((n_frames, N_MELS), device=y.device, dtype=torch.float32)
    scratch = torch.empty(1, device=y.device, dtype=torch.int32)
    rc = L.load().m3t_logmel(L.ptr(y), L.i64(y.numel()), L.i32(hop), L.i32(WIN), L.i32(N_MELS), L.ptr(_fb_cache[key]),
                             L.i32(1 if pad_mode == "reflect" else 0), L.f32(top_db if top_db else 0.0), L.ptr(out),
                             L.ptr(scratch), L.stream_ptr())
    L.check(rc, "m3t_logmel")
    return out


def stack_audio_windows(mel, start_idx, w_len):
    """models/dataset.py:83-95 `load_audio`: (w_len, 200) = 5 consecutive 40-bin frames at stride 3, zero tail."""
    out = torch.empty((w_len, 5 * mel.shape[1]), device=mel.device, dtype=torch.float32)
    rc = L.load().m3t_mel_stack(L.ptr(mel.contiguous()), L.i64(mel.shape[0]), L.i32(mel.shape[1]), L.i64(start_idx),
                                L.i32(w_len), L.ptr(out), L.stream_ptr())
    L.check(rc, "m3t_mel_stack")
    return out


def _read_wav_16k(path):
    from scipy.io import wavfile
    sr, data = wavfile.read(path)
    if sr != SR:
        raise RuntimeError("%s: expected a 16 kHz file (resampling is outside the hot path), got %d Hz" % (path, sr))
    if data.dtype.kind == "i":
        data = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)
    if data.ndim == 2:
        data = data.mean(axis=1)
    return data.astype(np.float32)


def extract_melspec(task, pad_mode=DEFAULT_PAD_MODE):
    """Same contract as the reference: returns 1 if the output exists, 0 on success, -1 on error."""
    fps, src_wav, dst_npy = task
    src_wav = src_wav.replace('_left', '').replace('_right', '')
    if os.path.exists(dst_npy):
        return 1
    try:
        y = torch.from_numpy(_read_wav_16k(src_wav)).cuda()
        spec = melspectrogram_db(y, fps, pad_mode=pad_mode)
        np.save(dst_npy, spec.cpu().numpy())        # (time, channels), as the reference stores it
        return 0
    except Exception as e:  # noqa: BLE001
        print('Exception on {}: {}'.format(src_wav, e))
        return -1


def build_tasks(src_dir, dst_dir, frames_fps_csv='splits/frames_fps.csv'):
    """(fps, wav path, npy path) for every video of at least 15 fps listed in splits/frames_fps.csv
    (reference :31-39; slower videos get all-zero audio features in models/dataset.py:277-278)."""
    tasks = []
    with open(frames_fps_csv) as f:
        for line in f.read().splitlines():
            vid_name, _, fps = line.split(',')
            if float(fps) >= 15:
                tasks.append((float(fps), os.path.join(src_dir, vid_name + '.wav'),
                              os.path.join(dst_dir, vid_name + '.npy')))
    return tasks


def main(argv=None):
    """`python -m m3t_b200.process.extract_melspec <wav dir> <npy dir>`: the reference script's command line
    (:27-46).  One process feeds the GPU file by file instead of a 16-process librosa pool."""
    import sys
    argv = list(sys.argv[1:] if argv is None else argv)
    pad_mode = DEFAULT_PAD_MODE
    for a in [a for a in argv if a.startswith('--pad_mode')]:      # optional; the reference's two positionals stay
        i = argv.index(a)
        pad_mode = a.split('=', 1)[1] if '=' in a else argv.pop(i + 1)
        argv.remove(a)
    src_dir, dst_dir = argv[0], argv[1]
    tasks = build_tasks(src_dir, dst_dir)
    for done, task in enumerate(tasks, 1):
        result = extract_melspec(task, pad_mode)
        if result <= 0:
            print('Finished {}, result: {}, progress: {}/{}'.format(task[1], result, done, len(tasks)))
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
