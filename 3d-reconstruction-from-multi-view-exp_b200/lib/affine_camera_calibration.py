"""Shadow module for the reference's ``lib/affine_camera_calibration.py``: the three
self-calibrations (``:7``, ``:59``, ``:137``) factorise the observation matrix on the GPU
(``affine_calibration.py``); every other name is the reference's own (loaded from the next
``lib/affine_camera_calibration.py`` on ``sys.path``).  Inputs the kernels do not take (more than
64 images, fewer than 4 points) go to the reference's implementation.  Only meaningful when the
reference checkout follows this package's directory on ``sys.path``."""
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
        cand = os.path.join(entry or ".", "lib", "affine_camera_calibration.py")
        if os.path.isfile(cand) and os.path.dirname(os.path.abspath(cand)) != _HERE:
            spec = importlib.util.spec_from_file_location("lib._reference_affine_camera_calibration", cand)
            mod = importlib.util.module_from_spec(spec)
            sys.modules["lib._reference_affine_camera_calibration"] = mod
            spec.loader.exec_module(mod)
            return mod
    raise ImportError("the reference's lib/affine_camera_calibration.py is not on sys.path behind this "
                      "shadow module")


_ref = _load_reference_module()
_gpu = importlib.import_module(os.path.basename(_PKG_DIR) + ".affine_calibration")
for _name in dir(_ref):
    if not _name.startswith("__"):
        globals()[_name] = getattr(_ref, _name)


def _fits(data_list):
    try:
        return 2 <= len(data_list) <= _gpu.MAX_IMAGES and len(data_list[0]) >= 4
    except TypeError:
        return False


def orthographic_self_calibration(data_list):
    if not _fits(data_list):
        return _ref.orthographic_self_calibration(data_list)
    return _gpu.orthographic_self_calibration(data_list)


def symmetric_affine_self_calibration(data_list):
    if not _fits(data_list):
        return _ref.symmetric_affine_self_calibration(data_list)
    return _gpu.symmetric_affine_self_calibration(data_list)


def paraperspective_self_calibration(data_list, f):
    if not _fits(data_list):
        return _ref.paraperspective_self_calibration(data_list, f)
    return _gpu.paraperspective_self_calibration(data_list, f)
