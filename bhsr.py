"""Alias: `import bhsr` == the package in super-resolution-building-height-estimation_b200/
(the directory name the build contract fixes is not a valid Python identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("super-resolution-building-height-estimation_b200")
sys.modules[__name__] = _pkg
