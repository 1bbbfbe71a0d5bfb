"""Generate tests/golden/video_input.pt with the UNMODIFIED reference `load_video` (models/dataset.py:46-80): synthetic
128x128 frames are written as JPEGs into a temporary directory, load_video reads them back with its own cv2.imread,
crop / mirror / cutout under seeded `random` and `numpy.random`; the same files decoded with cv2.imread are stored as
the uint8 input, and the draws are replayed through m3t_b200.process.video_input.draw_params with the same seeds.
Run on the build box only:  python -m oracle.make_golden_video_input
"""
import os
import random
import sys
import tempfile

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import _refload  # noqa: E402


def main():
    ds = _refload.load("dataset")
    sys.path.insert(0, ROOT)
    from m3t_b200.process.video_input import draw_params
    T, size = 2, 128
    rng = np.random.default_rng(11)
    clips = []
    for ci, (train, mirror, crop, cut) in enumerate(((True, True, True, True), (True, False, True, True),
                                                    (False, False, True, False))):
        with tempfile.TemporaryDirectory() as td:
            yy, xx = np.mgrid[0:size, 0:size]
            for t in range(T):
                img = np.stack([(yy * 2 + t * 9 + ci * 31) % 256, (xx * 2 + t * 5) % 256, ((yy + xx) + t * 17) % 256], -1)
                img = (img + rng.integers(0, 40, size=img.shape)).clip(0, 255).astype(np.uint8)
                cv2.imwrite(os.path.join(td, "%05d.jpg" % (t + 1)), img)
            decoded = np.stack([cv2.imread(os.path.join(td, "%05d.jpg" % (t + 1))) for t in range(T)])
            random.seed(100 + ci)
            np.random.seed(200 + ci)
            seq = ds.load_video(td, 0, T, is_training=train, mirror_augment=mirror, crop_augment=crop,
                                cutout_augment=cut, input_size=size)
        random.seed(100 + ci)
        np.random.seed(200 + ci)
        row = draw_params(train, mirror, crop, cut, size, 112)
        clips.append({"frames": torch.from_numpy(decoded), "params": row, "seq": torch.from_numpy(seq).half(),   # 0..255 integers and 127.5: exact in fp16
                      "flags": (train, mirror, crop, cut)})
    out = os.path.join(ROOT, "tests", "golden", "video_input.pt")
    torch.save(clips, out)
    print("wrote", out, os.path.getsize(out), "bytes;", [c["params"] for c in clips])


if __name__ == "__main__":
    main()
