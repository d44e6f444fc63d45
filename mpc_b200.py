"""Importable alias of the package directory
``mpc-for-dynamic-locomotion-in-the-mit-cheetah-3_b200`` (whose name contains hyphens)."""
import importlib as _importlib
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
if _here not in _sys.path:
    _sys.path.insert(0, _here)
_pkg = _importlib.import_module("mpc-for-dynamic-locomotion-in-the-mit-cheetah-3_b200")
_sys.modules[__name__] = _pkg
