"""ORACLE support (test infrastructure): import the UNMODIFIED reference modules from /root/reference.

On the build box this is /root/reference; on the GPU box (no /root/reference) it is the byte-for-byte vendored copy
oracle/_ref written by oracle/make_ref.py (git-ignored build output that travels with the snapshot).  `pytorch_lightning` and `matplotlib` are not
installed, so tiny stand-ins are registered before `models.model` / `models.utils` are imported;
nothing in the reference tree is changed or copied.
"""
import importlib
import os
import sys
import types

import torch.nn as nn

def _find_root():
    """$M3T_REFERENCE, /root/reference (build box), else the vendored copy oracle/_ref (GPU box; oracle/make_ref.py)."""
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in (os.environ.get("M3T_REFERENCE"), "/root/reference", os.path.join(here, "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "models", "model.py")):
            return cand
    return os.environ.get("M3T_REFERENCE", "/root/reference")


REFERENCE_ROOT = _find_root()


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "model.py"))


def _install_stubs():
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            pass

        pl.LightningModule = LightningModule
        pl.data_loader = lambda f: f
        pl.Trainer = object
        sys.modules["pytorch_lightning"] = pl
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


def load(name):
    """Import `models.<name>` from the reference tree (e.g. 'backbone', 'rnn', 'tcn', 'att_fusion', 'model')."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # our own package also has a `models` sub-package; the reference's top-level `models` must win here
    mod = sys.modules.get("models")
    if mod is not None and not any(str(q).startswith(REFERENCE_ROOT) for q in getattr(mod, "__path__", [])):
        raise RuntimeError("a different top-level `models` package is already imported")
    return importlib.import_module("models." + name)


def hparams(**over):
    """The reference's argparse defaults (models/model.py:448-493, train.py:46-57) as a Namespace."""
    import argparse
    d = dict(backbone="v2p_split", backend="gru", modality="visual", fusion_type="concat", freeze_enc=False,
             resample=False, mode="video", window=32, windows_per_epoch=200, learning_rate=5e-5, min_lr=1e-8,
             decay_factor=0.5, batch_size=96, optimizer="adam", scheduler="plateau", test_lr=False,
             test_on_val=False, loss="ccc_mtl", loss_lambda=0.5, num_hidden=512, split_layer=3, num_fc_layers=2,
             cutout=False, distributed=False, dataset_path="", release="vipl", input_size=256, checkpoint_path=".",
             workers=8, max_nb_epochs=30, gpus="1", nodes=1, seed=12345, fusion_checkpoint="", checkpoint="")
    d.update(over)
    return argparse.Namespace(**d)
