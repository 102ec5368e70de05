"""Importable alias of the package directory ``tsbb15-3d-reconstruction-project_b200`` (hyphens are not valid in an
``import`` statement).  ``import tsbb15_b200`` gives the package object itself."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("tsbb15-3d-reconstruction-project_b200")
sys.modules[__name__] = _pkg
