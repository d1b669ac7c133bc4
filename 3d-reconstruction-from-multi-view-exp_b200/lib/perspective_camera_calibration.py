"""Shadow module for the reference's ``lib/perspective_camera_calibration.py``: everything is the
reference's own code (loaded from the next ``lib/perspective_camera_calibration.py`` on
``sys.path``) except the heavy stages of ``perspective_self_calibration`` (``:513-540``), which run
on the GPU:

  * ``_compute_projective_depth_primary_method`` (``:61-144``) and
    ``_compute_projective_depth_dual_method`` (``:147-235`` -- the script's choice, an
    (n_images, n_points, n_points) array in the reference, O(n_images x n_points) here);
  * ``factorization_method`` (``lib/factorization.py:5-15``, a full SVD with an n_points x n_points
    factor in the reference) for the rank-4 case.

The O(n_images) Euclidean upgrade (``:238-411``) and the O(n_points) NumPy lines of
``_reconstruct_3d`` / ``correct_world_coordinates`` stay the reference's own code.  The replacements
are installed in the reference module's own namespace, so that its ``perspective_self_calibration``
picks them up.  Inputs the kernels do not take (more than 64 images, fewer than 4 points, another
rank) go to the reference's implementation.  Only meaningful when the reference checkout follows
this package's directory on ``sys.path``."""
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
_gpu = importlib.import_module(os.path.basename(_PKG_DIR) + ".projective_depth")
_ref_primary = getattr(_ref, "_compute_projective_depth_primary_method", None)
_ref_dual = getattr(_ref, "_compute_projective_depth_dual_method", None)
_ref_factorization = getattr(_ref, "factorization_method", None)


def _fits(x):
    return x.ndim == 3 and x.shape[0] >= 4 and 2 <= x.shape[1] <= _gpu.MAX_IMAGES


def _primary(x, f0, tolerance, max_iter=200):
    if not _fits(x):
        return _ref_primary(x, f0, tolerance, max_iter)
    return _gpu.compute_projective_depth_primary_method(x, f0, tolerance, max(int(max_iter), 1))


def _dual(x, f0, tolerance, max_iter=50):
    if not _fits(x):
        return _ref_dual(x, f0, tolerance, max_iter)
    return _gpu.compute_projective_depth_dual_method(x, f0, tolerance, max_iter)


def _factorization(W, n_rank=4):
    if n_rank != 4 or W.ndim != 2 or not (4 <= W.shape[0] <= 3 * _gpu.MAX_IMAGES) or W.shape[1] < 4:
        return _ref_factorization(W, n_rank)
    return _gpu.factorization_method(W, n_rank)


if _ref_primary is not None:
    _ref._compute_projective_depth_primary_method = _primary
if _ref_dual is not None:
    _ref._compute_projective_depth_dual_method = _dual
if _ref_factorization is not None:
    _ref.factorization_method = _factorization
for _name in dir(_ref):
    if not _name.startswith("__"):
        globals()[_name] = getattr(_ref, _name)
