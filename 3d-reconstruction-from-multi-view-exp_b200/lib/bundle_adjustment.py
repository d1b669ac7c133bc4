"""Shadow module for the reference's ``lib/bundle_adjustment.py``.

The reference's ``lib`` directory has no ``__init__.py`` (a PEP 420 namespace package), so
putting this package's directory *before* the reference checkout on ``sys.path`` makes
``from lib.bundle_adjustment import BundleAdjuster`` resolve here while ``lib.camera``,
``lib.utils`` ... still come from the reference: ``euclidiean_reconstruction.py`` runs
unchanged on the B200 engine (``run_reference_script.py``).
"""
import importlib
import os
import sys

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ROOT = os.path.dirname(_PKG_DIR)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

BundleAdjuster = importlib.import_module(
    os.path.basename(_PKG_DIR) + ".bundle_adjuster").BundleAdjuster

__all__ = ["BundleAdjuster"]
