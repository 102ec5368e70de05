"""Flat-module drop-in: put this directory first on sys.path and the reference's `import correspondences` resolves here."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
import tsbb15_b200 as _pkg  # noqa: E402
from importlib import import_module as _imp  # noqa: E402

_mod = _imp(_pkg.__name__ + ".correspondences")
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
