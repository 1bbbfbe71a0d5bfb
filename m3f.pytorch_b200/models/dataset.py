"""Aff-Wild2 window dataset feeding the hot path (reference: models/dataset.py `AffWild2SequenceDataset`, `load_video`,
`load_audio`, `sequence_cutout`, `one_runs`), SURVEY 8(f) N1/N2: the caller right in front of `AffWild2VA.forward`.

Same constructor arguments, file layout (`splits/*.csv` relative to the working directory, `<path>/face_<size>` or
`cropped_aligned`, `annotations/{VA,EXPR}_Set`, `se101_feats`, `AU_feats`, `mel_spec`), sample dictionary keys,
padding rules and — so that a seeded run visits the same windows with the same augmentation — the same consumption
order of `random` / `numpy.random` (checked against the reference class on a synthetic tree by
tests/test_cpu_trainer.py).  This is host-side Python by nature (JPEG decode + numpy); the device-side half of the
input pipeline is `process/video_input.py` + `m3t_video_augment_prep_s2d_w4`.

`emit_u8=True` (not in the reference) hands the decoded uint8 frames and the augmentation draws to that kernel instead
of a float32 clip: batch['video_u8'] (T,S,S,3) + batch['video_aug'] (8,) — 3x less host->device traffic.
"""
import os
import pickle
import random

import cv2
import numpy as np
import torch
from torch.utils.data import Dataset

from ..process.video_input import draw_params

_FOLD = {'train': 'Training_Set', 'val': 'Validation_Set'}


def sequence_cutout(seq, n_holes=1, fill_value=127.5):
    """Cutout over a (C,T,H,W) clip: one square hole of side H (clipped at the border) shared by all frames
    (reference :16-31).  127.5 becomes 0 after normalisation."""
    h, w = seq.shape[-2:]
    half = h // 2
    for _ in range(n_holes):
        cy = np.random.randint(h)
        cx = np.random.randint(w)
        ys = slice(max(cy - half, 0), min(cy + half, h))
        xs = slice(max(cx - half, 0), min(cx + half, w))
        seq[..., ys, xs] = fill_value
    return seq


def one_runs(a):
    """[start, end) index pairs of the runs of ones in a 1-D 0/1 array (reference :36-43)."""
    flags = np.concatenate(([False], np.asarray(a) == 1, [False]))
    edges = np.flatnonzero(flags[1:] != flags[:-1])
    return edges.reshape(-1, 2)


def _read_frames(path, start, length, crop=None, resize=False, flip=False):
    """Decoded frames start+1 .. start+length (file names are 1-based); a missing file repeats the previous frame,
    or is a black 112x112 image at the head of the window (reference :62-66)."""
    frames = []
    for idx in range(start, start + length):
        img = cv2.imread(os.path.join(path, '%05d.jpg' % (idx + 1)))
        if img is None:
            img = frames[-1] if frames else np.zeros((112, 112, 3), dtype=np.uint8)
        else:
            if crop is not None:
                x0, y0, side = crop
                img = img[y0:y0 + side, x0:x0 + side]
                if resize:
                    img = cv2.resize(img, (112, 112))
            if flip:
                img = cv2.flip(img, 1)
        frames.append(img)
    return frames


def load_video(path, start, length, is_training=False, mirror_augment=False, crop_augment=False,
               cutout_augment=False, input_size=256):
    """float32 (3,T,112,112) clip with values 0..255 in cv2's BGR order (reference :46-80).  One crop window per
    clip; random when training, centred otherwise; mirror / cutout only when training."""
    crop, resize = None, False
    if crop_augment:
        side = input_size * 7 // 8
        if is_training:
            x0 = random.randint(0, input_size // 8)
            y0 = random.randint(0, input_size // 8)
        else:
            x0 = y0 = input_size // 16
        crop, resize = (x0, y0, side), input_size > 128
    frames = _read_frames(path, start, length, crop, resize, mirror_augment and is_training)
    seq = np.stack(frames).transpose(3, 0, 1, 2).astype(np.float32)
    if cutout_augment and is_training:
        seq = sequence_cutout(seq)
    return seq


def load_video_u8(path, start, length, is_training=False, mirror_augment=False, crop_augment=False,
                  cutout_augment=False, input_size=128):
    """Same draws as `load_video`, but crop / mirror / cutout / normalisation are left to the stem's input kernel:
    returns (uint8 [T,S,S,3] decoded frames, int32 [8] parameter row).  Frames of 256-pixel tracks are resized to
    128 first (the window then maps to the same pixels up to interpolation; the reference resizes after cropping)."""
    row = draw_params(is_training, mirror_augment, crop_augment, cutout_augment, min(input_size, 128))
    frames = _read_frames(path, start, length)
    side = 128 if crop_augment else 112
    frames = [f if f.shape[0] == side else cv2.resize(f, (side, side)) for f in frames]
    return np.stack(frames), np.asarray(row, dtype=np.int32)


def load_audio(audio_path, start_idx, w_len):
    """(w_len, 200): for video frame t the five log-Mel frames 3t .. 3t+4, zero-padded past the end of the file
    (reference :83-95)."""
    mel = np.load(audio_path)
    rows = 3 * (start_idx + np.arange(w_len))[:, None] + np.arange(5)[None, :]
    ok = rows < len(mel)
    out = np.zeros((w_len, 5, mel.shape[1]), dtype=mel.dtype)
    out[ok] = mel[rows[ok]]
    return out.reshape(w_len, -1)


def _pad_edge(a, n, axis):
    if n <= 0:
        return a
    width = [(0, 0)] * a.ndim
    width[axis] = (0, n)
    return np.pad(a, width, 'edge')


class AffWild2SequenceDataset(Dataset):
    """Windows of `window_len` frames of the Aff-Wild2 face tracks (reference :98-346).

    split 'train': `windows_per_epoch` random windows per video and epoch, drawn from the windows in which every frame
    has an image and a valence/arousal label in [-1,1]; 'val' / 'test': consecutive windows with stride
    window_len // inv_test_stride, the last one padded with edge values (`length` tells how many frames are real).
    """

    def __init__(self, split, path, window_len=16, windows_per_epoch=20, apply_cutout=True, release='ibug',
                 input_size=112, modality='visual', noise_and_balance=False, inv_test_stride=1, emit_u8=False):
        self.split, self.path = split, path
        self.window_len, self.windows_per_epoch = window_len, windows_per_epoch
        self.apply_cutout, self.release, self.input_size, self.modality = apply_cutout, release, input_size, modality
        self.emit_u8 = emit_u8
        self.base = os.path.join(path, 'cropped_aligned' if release == 'ibug' else 'face_%d' % input_size)

        self.nb_frames, self.fps = {}, {}
        with open('splits/frames_fps.csv') as f:
            for line in f.read().splitlines():
                name, n, fps = line.split(',')
                self.nb_frames[name], self.fps[name] = int(n), float(fps)
        with open('splits/%s.csv' % split) as f:
            self.files = f.read().splitlines()
        if modality == 'audio':     # the audio features of < 15 fps videos are unusable
            self.files = [v for v in self.files if self.fps[v] >= 15.0]

        if split == 'train':
            self.sample_src = list(range(len(self.files))) * windows_per_epoch
            random.shuffle(self.sample_src)
        else:
            stride = window_len // inv_test_stride
            self.sample_src = [(i, s) for i, v in enumerate(self.files) for s in range(0, self.nb_frames[v], stride)]

        if split != 'test':
            self._read_labels()
        if split == 'train':
            self.avail_windows = (self.get_noisy_balanced_windows() if noise_and_balance
                                  else self.get_available_windows())
        print('Loaded partition {}: {} files, {} windows'.format(split, len(self.files), len(self.sample_src)))

    # ------------------------------------------------------------------ annotations
    def _annotation(self, kind, fold, vid):
        with open(os.path.join(self.path, 'annotations', kind, fold, vid + '.txt')) as f:
            return f.read().splitlines()

    def _read_labels(self):
        self.labels_va, self.labels_expr, self.labels_au = {}, {}, {}
        for vid in self.files:
            self.labels_va[vid] = np.loadtxt(self._annotation('VA_Set', _FOLD[self.split], vid), delimiter=',',
                                             skiprows=1, dtype=np.float32)
        known = set(self.files)
        with open('splits/expr.csv') as f:
            for line in f.read().splitlines():
                vid, fold = line.split(',')
                if vid in known:
                    self.labels_expr[vid] = np.loadtxt(self._annotation('EXPR_Set', fold, vid), skiprows=1,
                                                       dtype=np.int64)

    def _has_label(self, vid):
        return np.abs(self.labels_va[vid]).max(axis=1) <= 1

    def _has_image(self, vid):
        fold = os.path.join(self.base, vid)
        return np.array([os.path.exists(os.path.join(fold, '%05d.jpg' % (i + 1)))
                         for i in range(len(self.labels_va[vid]))])

    def _cached(self, cache_path, scan):
        """Window lists are cached in the working directory under the reference's file names.  Unlike the reference
        the file is written atomically and an unreadable one is rescanned: with one process per GPU every rank
        builds the dataset at the same moment, and a rank must never read another rank's half-written file."""
        if os.path.exists(cache_path):
            try:
                with open(cache_path, 'rb') as f:
                    return pickle.load(f)
            except (EOFError, pickle.UnpicklingError):
                pass
        windows = scan()
        part = '%s.%d.part' % (cache_path, os.getpid())
        with open(part, 'wb') as f:
            pickle.dump(windows, f)
        os.replace(part, cache_path)
        return windows

    def _run_starts(self, ok):
        starts = []
        for lo, hi in one_runs(ok):
            starts.extend(range(lo, hi - self.window_len + 1))
        return starts

    def get_available_windows(self):
        """Start frames of the windows lying wholly inside a run of usable frames (reference :201-224); cached in the
        working directory under the reference's file name."""
        def scan():
            windows = {}
            for vid in self.files:
                ok = self._has_label(vid)
                if self.modality != 'audio':
                    ok = self._has_image(vid) & ok
                windows[vid] = self._run_starts(ok)
                if self.modality != 'audio':
                    assert len(windows[vid]) > 0, 'no available windows for {}'.format(vid)
            return windows
        return self._cached('{}_{}_window{}_{}.pkl'.format(self.release, self.split, self.window_len, self.modality),
                            scan)

    def get_noisy_balanced_windows(self):
        """Variant tolerating 25 % missing frames / labels per window and listing windows of negative mean valence
        twice (reference :157-199)."""
        W = self.window_len

        def scan():
            windows = {}
            for vid in self.files:
                va = self.labels_va[vid]
                ok_label = self._has_label(vid)
                if self.modality == 'audio':
                    runs = one_runs(ok_label)
                    starts = self._run_starts(ok_label)
                    # the reference pairs the i-th listed window with the i-th run (zip), not with its own start
                    scored = sorted(((w, va[lo:lo + W, 0].mean()) for w, (lo, _) in zip(starts, runs)),
                                    key=lambda t: t[1])
                    n_neg = int(np.searchsorted([s for _, s in scored], 0))
                    windows[vid] = starts + [w for w, _ in scored[:n_neg]]
                    continue
                ok_image = self._has_image(vid)
                va[~ok_label] = 0
                starts = []
                for s in range(0, len(va) - W + 1):
                    missing = max(1 - ok_image[s:s + W].sum() / W, 1 - ok_label[s:s + W].sum() / W)
                    if missing > 0.25:
                        continue
                    starts.extend([s, s] if va[s:s + W, 0].mean() < 0 else [s])
                assert len(starts) > 0, 'no available windows for {}'.format(vid)
                windows[vid] = starts
            return windows
        return self._cached('{}_{}_noisybalancedwindow{}_{}.pkl'.format(self.release, self.split, W, self.modality),
                            scan)

    # ------------------------------------------------------------------ samples
    def __len__(self):
        return len(self.sample_src)

    def _features(self, folder, vid, start, n, width=None):
        """(C, n) slice of a per-video (frames, C) feature file, edge-padded when the file is shorter than the
        annotation (reference :253-266)."""
        feats = np.load(os.path.join(self.path, folder, vid + '.npy'))[start:start + n]
        if width is not None:
            feats = feats[:, :width]
        feats = feats.transpose()
        return _pad_edge(feats, n - feats.shape[-1], 1)

    def __getitem__(self, i):
        W = self.window_len
        training = self.split == 'train'
        if training:
            vid = self.files[self.sample_src[i]]
            n = W
            start = random.choice(self.avail_windows[vid])
        else:
            vi, start = self.sample_src[i]
            vid = self.files[vi]
            n = min(W, self.nb_frames[vid] - start)
        pad = W - n
        batch = {'vid_name': vid, 'start': start, 'length': n}

        if 'visual' in self.modality:
            loader = load_video_u8 if self.emit_u8 else load_video
            clip = loader(os.path.join(self.base, vid), start, n, training, random.random() > 0.5,
                          self.release == 'vipl', self.apply_cutout, self.input_size)
            if self.emit_u8:
                batch['video_u8'] = torch.from_numpy(_pad_edge(clip[0], pad, 0))
                batch['video_aug'] = torch.from_numpy(clip[1])
            else:
                batch['video'] = torch.from_numpy(_pad_edge(clip, pad, 1))
            batch['se_features'] = torch.from_numpy(_pad_edge(self._features('se101_feats', vid, start, n), pad, 1))
            batch['au_features'] = torch.from_numpy(_pad_edge(self._features('AU_feats', vid, start, n, 256), pad, 1))

        if 'audio' in self.modality:
            if self.fps[vid] < 15:
                audio = np.zeros((W, 200), dtype=np.float32)
            else:
                audio = load_audio(os.path.join(self.path, 'mel_spec', vid + '.npy'), start, n)
            batch['audio'] = torch.from_numpy(_pad_edge(audio, W - len(audio), 0)[:W])

        if self.split != 'test':
            va = _pad_edge(self.labels_va[vid][start:start + n], pad, 0)
            if vid in self.labels_expr:
                expr = self.labels_expr[vid][start:start + n]
                valid = expr >= 0
            else:
                expr = np.zeros(n, dtype=np.int64)
                valid = np.zeros(n, dtype=bool)
            batch['label_valence'] = torch.from_numpy(va[..., 0])
            batch['class_expr'] = _pad_edge(np.clip(expr, 0, 6), pad, 0)      # class ids outside 0..6 are invalid
            batch['expr_valid'] = _pad_edge(valid, pad, 0)
            batch['label_arousal'] = torch.from_numpy(va[..., 1])
        return batch
