"""B200-native Levenberg-Marquardt bundle adjustment (drop-in for the reference's
``lib/bundle_adjustment.py::BundleAdjuster``).

Only the hot path lives here: ``csrc/`` (sm_100a CUDA kernels + the C-ABI of
``include/ba_b200.h``), the ctypes binding, and the host-side mirror of the reference class.
Importing the package does not load the CUDA library; constructing an engine does, and
fails loudly if it is missing (there is no CPU fallback).
"""
from . import scenes  # noqa: F401


def __getattr__(name):
    if name in ("BundleAdjuster", "ObservationList"):
        from . import bundle_adjuster

        return getattr(bundle_adjuster, name)
    if name in ("calc_projected_points", "project_all"):
        from . import projection

        return getattr(projection, name)
    if name in ("projective_depth_primary", "compute_projective_depth_primary_method", "projective_depth_dual",
                "compute_projective_depth_dual_method", "factorize_rank4", "factorization_method"):
        from . import projective_depth

        return getattr(projective_depth, name)
    if name in ("orthographic_self_calibration", "symmetric_affine_self_calibration",
                "paraperspective_self_calibration"):
        from . import affine_calibration

        return getattr(affine_calibration, name)
    if name == "Engine":
        from . import engine

        return engine.Engine
    raise AttributeError(name)
