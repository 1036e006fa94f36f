"""Importable alias of the package directory `3d_vit_ensemble_b200/` (a Python identifier cannot
start with a digit).  `import vit3d_b200` IS that package: every `vit3d_b200.x.y` resolves to the
same module object as `3d_vit_ensemble_b200.x.y` (no duplicate module copies)."""
import importlib
import importlib.abc
import importlib.util
import os
import sys

_REAL = "3d_vit_ensemble_b200"
_ALIAS = __name__

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, real_name):
        self.real_name = real_name

    def create_module(self, spec):
        return importlib.import_module(self.real_name)

    def exec_module(self, module):
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname == _ALIAS or fullname.startswith(_ALIAS + "."):
            real = _REAL + fullname[len(_ALIAS):]
            return importlib.util.spec_from_loader(fullname, _AliasLoader(real))
        return None


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
_pkg = importlib.import_module(_REAL)
for _k, _m in list(sys.modules.items()):
    if _k == _REAL or _k.startswith(_REAL + "."):
        sys.modules[_ALIAS + _k[len(_REAL):]] = _m
sys.modules[_ALIAS] = _pkg
