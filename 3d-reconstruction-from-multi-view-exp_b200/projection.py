"""Batched re-projection on the GPU: mirror of the reference's ``calc_projected_points``
(``lib/camera.py:74-81``), the step right after bundle adjustment in both reference scripts
(``euclidiean_reconstruction.py:63``, ``affine_reconstruction.py:64``).

Same arguments and return value (a list with one ``(N, 2)`` array per camera); the Python loop
over cameras becomes one kernel launch behind ``ba_project_points`` of the C ABI.  No CPU path.
"""
from __future__ import annotations

import numpy as np

from . import _cabi
from .bundle_adjuster import _default_device


def project_all(X, K, R, t, device: int | None = None) -> np.ndarray:
    """All cameras at once: ``(M, N, 2)`` image points (perspective division included)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    K = np.ascontiguousarray(np.asarray(K), dtype=np.float64)
    R = np.ascontiguousarray(R, dtype=np.float64)
    t = np.ascontiguousarray(t, dtype=np.float64)
    if X.ndim != 2 or X.shape[1] != 3:
        raise ValueError("X must be (N, 3)")
    M = R.shape[0]
    if K.shape != (M, 3, 3) or R.shape != (M, 3, 3) or t.shape != (M, 3):
        raise ValueError("K, R must be (M, 3, 3) and t (M, 3)")
    out = np.empty((M, X.shape[0], 2), dtype=np.float64)
    lib = _cabi.load()
    dev = _default_device() if device is None else int(device)
    _cabi.check(lib.ba_project_points(dev, X.shape[0], M, X.ctypes.data, K.ctypes.data, R.ctypes.data,
                                      t.ctypes.data, out.ctypes.data, _cabi.BA_MEM_HOST, None))
    return out


def calc_projected_points(X, K, R, t):
    """Reference signature (``lib/camera.py:74``): list of per-camera ``(N, 2)`` arrays."""
    return list(project_all(X, K, R, t))
