"""Alias: `import bhsr` == the package in super-resolution-building-height-estimation_b200/
(the directory name the build contract fixes is not a valid Python identifier).  Every submodule
is registered under both names so `bhsr.x` and the real package share one module object."""
import importlib
import os
import pkgutil
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_REAL = "super-resolution-building-height-estimation_b200"
_pkg = importlib.import_module(_REAL)
for _m in pkgutil.iter_modules(_pkg.__path__):
    if _m.name in ("build",):
        continue
    _sub = importlib.import_module(f"{_REAL}.{_m.name}")
    sys.modules[f"{__name__}.{_m.name}"] = _sub
sys.modules[__name__] = _pkg
