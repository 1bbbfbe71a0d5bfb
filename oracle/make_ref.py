"""ORACLE support (test infrastructure): vendor the UNMODIFIED reference modules into oracle/_ref/ so that
`bench.py --impl reference` and `cpu_baseline` can time the reference ITSELF on the GPU box's host cores
(/root/reference does not exist there).

    python -m oracle.make_ref          (also run by __graft_entry__.build() whenever /root/reference is present)

oracle/_ref/ is build output: git-ignored (no reference source enters the history), not gpurun-ignored (it travels to
the GPU box with the snapshot, like the built .so).  Files are byte-for-byte copies of `models/*.py`; nothing is
edited.  The two packages the reference imports that are absent from this image (pytorch_lightning, matplotlib) are
stubbed at import time by oracle/_refload.py, exactly as for the golden-fixture generators.
"""
import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("M3T_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")
FILES = ["models/%s.py" % n for n in ("att_fusion", "backbone", "cbam", "densenet", "model", "resnet", "rnn", "tcn",
                                      "utils", "vggface", "dataset", "cv_augment", "lr_finder")]


def main():
    if not os.path.isdir(os.path.join(SRC, "models")):
        return False
    os.makedirs(os.path.join(DST, "models"), exist_ok=True)
    manifest = {}
    for rel in FILES:
        src = os.path.join(SRC, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(DST, rel)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    init = os.path.join(SRC, "models", "__init__.py")
    if os.path.exists(init):
        shutil.copyfile(init, os.path.join(DST, "models", "__init__.py"))
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    print("vendored" if main() else "reference tree not found at %s" % SRC)
