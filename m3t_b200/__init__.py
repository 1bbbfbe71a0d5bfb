"""Import alias for the product package.

The product lives in the directory ``m3f.pytorch_b200/`` (a name Python cannot import because of the dot);
this shim package points its ``__path__`` there, so ``import m3t_b200.models.backbone`` loads
``m3f.pytorch_b200/models/backbone.py``.
"""
import os as _os

_real = _os.path.normpath(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "m3f.pytorch_b200"))
if not _os.path.isdir(_real):  # pragma: no cover
    raise ImportError("m3t_b200: product directory %s is missing" % _real)
__path__ = [_real]
PACKAGE_DIR = _real
