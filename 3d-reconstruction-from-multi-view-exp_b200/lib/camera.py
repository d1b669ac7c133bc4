"""Shadow module for the reference's ``lib/camera.py``: everything is the reference's own code
(loaded from the next ``lib/camera.py`` on ``sys.path``) except ``calc_projected_points``
(``lib/camera.py:74-81``), which runs as one CUDA kernel.  Only meaningful when the reference
checkout follows this package's directory on ``sys.path``."""
import importlib
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG_DIR = os.path.dirname(_HERE)
_ROOT = os.path.dirname(_PKG_DIR)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)


def _load_reference_module():
    for entry in sys.path:
        cand = os.path.join(entry or ".", "lib", "camera.py")
        if os.path.isfile(cand) and os.path.dirname(os.path.abspath(cand)) != _HERE:
            # keep the package context (`lib.`) so the reference's relative imports resolve
            spec = importlib.util.spec_from_file_location("lib._reference_camera", cand)
            mod = importlib.util.module_from_spec(spec)
            sys.modules["lib._reference_camera"] = mod
            spec.loader.exec_module(mod)
            return mod
    raise ImportError("the reference's lib/camera.py is not on sys.path behind this shadow module")


_ref = _load_reference_module()
for _name in dir(_ref):
    if not _name.startswith("__"):
        globals()[_name] = getattr(_ref, _name)

calc_projected_points = importlib.import_module(
    os.path.basename(_PKG_DIR) + ".projection").calc_projected_points
