"""Shadow module for the reference's ``lib/perspective_camera_calibration.py``: everything is the
reference's own code (loaded from the next ``lib/perspective_camera_calibration.py`` on
``sys.path``) except ``_compute_projective_depth_primary_method`` (``:61-144``), which runs on the
GPU.  The replacement is installed in the reference module's own namespace, so that its
``perspective_self_calibration(..., method="primary")`` (``:513-540``) picks it up.  Only meaningful
when the reference checkout follows this package's directory on ``sys.path``."""
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
        cand = os.path.join(entry or ".", "lib", "perspective_camera_calibration.py")
        if os.path.isfile(cand) and os.path.dirname(os.path.abspath(cand)) != _HERE:
            spec = importlib.util.spec_from_file_location("lib._reference_perspective_camera_calibration", cand)
            mod = importlib.util.module_from_spec(spec)
            sys.modules["lib._reference_perspective_camera_calibration"] = mod
            spec.loader.exec_module(mod)
            return mod
    raise ImportError("the reference's lib/perspective_camera_calibration.py is not on sys.path behind this "
                      "shadow module")


_ref = _load_reference_module()
_ref._compute_projective_depth_primary_method = importlib.import_module(
    os.path.basename(_PKG_DIR) + ".projective_depth").compute_projective_depth_primary_method
for _name in dir(_ref):
    if not _name.startswith("__"):
        globals()[_name] = getattr(_ref, _name)
