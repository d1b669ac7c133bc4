"""Import alias for the package directory ``3d-reconstruction-from-multi-view-exp_b200/``.

The mandated directory name is not a Python identifier, so ``import`` statements cannot
spell it; this module loads it with importlib and re-exports it under a short name:

    import ba_b200
    ba_b200.BundleAdjuster(...)
"""
import importlib as _importlib
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.abspath(__file__))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)

PACKAGE_NAME = "3d-reconstruction-from-multi-view-exp_b200"
PACKAGE_DIR = _os.path.join(_ROOT, PACKAGE_NAME)
_pkg = _importlib.import_module(PACKAGE_NAME)


def submodule(name: str):
    """``submodule('scenes')`` -> the package's ``scenes`` module."""
    return _importlib.import_module(f"{PACKAGE_NAME}.{name}")


def __getattr__(name):
    return getattr(_pkg, name)
