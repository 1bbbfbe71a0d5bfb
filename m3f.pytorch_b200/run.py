"""Run one of the reference's scripts UNCHANGED on this implementation:

    python -m m3t_b200.run /path/to/m3f.pytorch/train.py --gpus 0 --modality audiovisual --backbone resnet ...
    torchrun --nproc-per-node 8 -m m3t_b200.run /path/to/m3f.pytorch/train.py --gpus 0,1,2,3,4,5,6,7 --distributed ...
    python -m m3t_b200.run /path/to/m3f.pytorch/eval.py --gpus 0 --checkpoint X --test_on_val ...

The script's two imports are redirected before it starts: `pytorch_lightning` -> m3t_b200.lightning (the 0.6 API the
scripts were written against; train.py:4, eval.py:3) and `models` / `models.*` -> m3t_b200.models (train.py:6,
eval.py:5).  `matplotlib` is only touched for its backend switch (train.py:1-2); when it is not installed an inert
stand-in is registered.  The script's own directory is NOT put on sys.path (its `models/` package must not win), but
stays the place `splits/*.csv` are read from when it is the working directory, as with the reference.
"""
import importlib
import os
import runpy
import sys
import types


def install_aliases():
    """Idempotent: register the module aliases the reference's scripts resolve their imports through."""
    from . import lightning
    sys.modules["pytorch_lightning"] = lightning
    models = importlib.import_module(__package__ + ".models")
    sys.modules["models"] = models
    for sub in ("model", "backbone", "vggm", "resnet", "rnn", "tcn", "att_fusion", "utils", "dataset", "lr_finder"):
        sys.modules["models." + sub] = importlib.import_module("%s.models.%s" % (__package__, sub))
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        mpl.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, mpl.pyplot


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 2
    script = os.path.abspath(argv[0])
    if not os.path.isfile(script):
        raise SystemExit("m3t_b200.run: no such script: %s" % script)
    install_aliases()
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
