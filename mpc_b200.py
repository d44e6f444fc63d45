"""Importable alias of the package directory
``mpc-for-dynamic-locomotion-in-the-mit-cheetah-3_b200`` (whose name contains hyphens).

``import mpc_b200`` and ``import mpc_b200.<sub>`` / ``from mpc_b200.<sub> import ...`` resolve
to the SAME module objects as the hyphen-named package and its submodules (one copy of every
class, one loaded ``libcmpc.so``): a meta-path finder maps the ``mpc_b200.`` prefix."""
import importlib as _importlib
import importlib.abc as _abc
import importlib.util as _util
import os as _os
import sys as _sys

_REAL = "mpc-for-dynamic-locomotion-in-the-mit-cheetah-3_b200"
_ALIAS = __name__

_here = _os.path.dirname(_os.path.abspath(__file__))
if _here not in _sys.path:
    _sys.path.insert(0, _here)


class _AliasLoader(_abc.Loader):
    def __init__(self, real):
        self._real = real

    def create_module(self, spec):
        return _importlib.import_module(self._real)       # the real module object itself

    def exec_module(self, module):                        # already executed under its real name
        pass


class _AliasFinder(_abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(_ALIAS + "."):
            return None
        real = _REAL + fullname[len(_ALIAS):]
        try:
            if _util.find_spec(real) is None:
                return None
        except ModuleNotFoundError:
            return None
        return _util.spec_from_loader(fullname, _AliasLoader(real))


if not any(isinstance(f, _AliasFinder) for f in _sys.meta_path):
    _sys.meta_path.insert(0, _AliasFinder())
_pkg = _importlib.import_module(_REAL)
_sys.modules[_ALIAS] = _pkg
for _name, _mod in list(_sys.modules.items()):             # submodules the package already imported
    if _name.startswith(_REAL + "."):
        _sys.modules[_ALIAS + _name[len(_REAL):]] = _mod
