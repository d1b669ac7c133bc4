"""Projective-depth iteration (primary method) on the GPU: mirror of the reference's
``_compute_projective_depth_primary_method`` (``lib/perspective_camera_calibration.py:61-144``),
the first stage of ``perspective_self_calibration`` -- the step before bundle adjustment in the
perspective pipeline.

Same arguments, return value and printed lines; the SVD and the per-point eigenproblems become a
Gram matrix on the FP64 tensor cores, one Jacobi eigen-solve and a per-point 4 x 4 problem behind
``ba_projective_depth_primary`` of the C ABI (``csrc/k7_projective_depth.cu``).  No CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from .bundle_adjuster import _default_device


def projective_depth_primary(x, f0: float, tolerance: float, max_iter: int = 200, device: int | None = None):
    """``(z, errors)``: depths ``(n_points, n_images)`` and the reprojection error of every iteration."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    if x.ndim != 3 or x.shape[2] != 3:
        raise ValueError("x must be (n_points, n_images, 3)")
    N, M = x.shape[:2]
    z = np.empty((N, M), dtype=np.float64)
    errors = np.empty(int(max_iter), dtype=np.float64)
    n_iter = C.c_int(0)
    lib = _cabi.load()
    dev = _default_device() if device is None else int(device)
    _cabi.check(lib.ba_projective_depth_primary(dev, N, M, x.ctypes.data, float(f0), float(tolerance), int(max_iter),
                                                z.ctypes.data, errors.ctypes.data, C.byref(n_iter),
                                                _cabi.BA_MEM_HOST, None))
    return z, errors[: n_iter.value].copy()


def compute_projective_depth_primary_method(x, f0: float, tolerance: float, max_iter: int = 200):
    """Reference signature and side effects (:61-63, :141, :144-145): prints one line per iteration."""
    z, errors = projective_depth_primary(x, f0, tolerance, max_iter)
    for count, E in enumerate(errors, start=1):
        print(f"Iteration {count}: reprojection_error = {E:.8}")
    if len(errors) >= max_iter:
        print("Did not converge because the maximum number of iterations was reached.")
    return z
