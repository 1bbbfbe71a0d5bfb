"""A miniature Aff-Wild2 tree in the layout models/dataset.py reads (test infrastructure): a `splits/` directory for
the working directory plus `<root>/{face_<size>|cropped_aligned, annotations/{VA,EXPR}_Set, se101_feats, AU_feats,
mel_spec}`.  Frames are smooth random JPEGs; some frames are missing, some labels are the dataset's "not annotated"
markers (-5 for valence/arousal, -1 for expressions) so that the window scan, the missing-frame rule and the masks
are all exercised."""
import os

import cv2
import numpy as np


def _frame(rng, size):
    low = rng.integers(0, 256, (size // 8, size // 8, 3)).astype(np.uint8)
    return cv2.resize(low, (size, size), interpolation=cv2.INTER_CUBIC)


def build(root, cwd, videos=None, release="vipl", input_size=128, seed=0):
    """videos: {name: dict(split=train|val|test, frames=int, fps=float, expr=bool, missing=[frame idx], bad=[label idx])}
    Writes the data tree under `root` and `splits/*.csv` under `cwd`; returns the videos dict."""
    rng = np.random.default_rng(seed)
    if videos is None:
        videos = {
            "vidA": dict(split="train", frames=40, fps=30.0, expr=True, missing=[17], bad=[5, 6]),
            "vidB_left": dict(split="train", frames=30, fps=25.0, expr=False, missing=[], bad=[]),
            "vidC": dict(split="val", frames=27, fps=30.0, expr=True, missing=[3], bad=[20]),
            "vidD": dict(split="val", frames=19, fps=12.0, expr=False, missing=[], bad=[]),
            "vidE": dict(split="test", frames=22, fps=30.0, expr=False, missing=[0], bad=[]),
        }
    os.makedirs(os.path.join(cwd, "splits"), exist_ok=True)
    with open(os.path.join(cwd, "splits", "frames_fps.csv"), "w") as f:
        f.write("".join("%s,%d,%s\n" % (n, v["frames"], v["fps"]) for n, v in videos.items()))
    for split in ("train", "val", "test"):
        with open(os.path.join(cwd, "splits", "%s.csv" % split), "w") as f:
            f.write("".join(n + "\n" for n, v in videos.items() if v["split"] == split))
    fold = {"train": "Training_Set", "val": "Validation_Set"}
    with open(os.path.join(cwd, "splits", "expr.csv"), "w") as f:
        f.write("".join("%s,%s\n" % (n, fold[v["split"]]) for n, v in videos.items()
                        if v["expr"] and v["split"] in fold))
    base = os.path.join(root, "cropped_aligned" if release == "ibug" else "face_%d" % input_size)
    for d in ("se101_feats", "AU_feats", "mel_spec"):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    for name, v in videos.items():
        n = v["frames"]
        os.makedirs(os.path.join(base, name), exist_ok=True)
        for i in range(n):
            if i in v["missing"]:
                continue
            cv2.imwrite(os.path.join(base, name, "%05d.jpg" % (i + 1)), _frame(rng, input_size))
        np.save(os.path.join(root, "se101_feats", name + ".npy"), rng.standard_normal((n - 1, 512)).astype(np.float32))
        np.save(os.path.join(root, "AU_feats", name + ".npy"), rng.standard_normal((n, 268)).astype(np.float32))
        np.save(os.path.join(root, "mel_spec", name + ".npy"),
                (rng.standard_normal((3 * n - 4, 40)) * 20 - 40).astype(np.float32))
        if v["split"] in fold:
            va = rng.uniform(-1, 1, (n, 2))
            va[v["bad"]] = -5
            d = os.path.join(root, "annotations", "VA_Set", fold[v["split"]])
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, name + ".txt"), "w") as f:
                f.write("valence,arousal\n" + "".join("%.3f,%.3f\n" % tuple(r) for r in va))
            if v["expr"]:
                ex = rng.integers(-1, 8, n)         # -1 = not annotated; 7 is out of range and gets clipped
                d = os.path.join(root, "annotations", "EXPR_Set", fold[v["split"]])
                os.makedirs(d, exist_ok=True)
                with open(os.path.join(d, name + ".txt"), "w") as f:
                    f.write("Neutral,Anger,Disgust,Fear,Happiness,Sadness,Surprise\n" + "".join("%d\n" % e for e in ex))
    return videos
