"""Affine self-calibrations: the initialisation that
``affine_reconstruction.py:42`` hands to the bundle adjustment.

Mirrors reference ``lib/affine_camera_calibration.py``: ``orthographic_self_calibration`` (``:7``),
``symmetric_affine_self_calibration`` (``:59``), ``paraperspective_self_calibration`` (``:137``) --
same names, arguments, return values ``(S.T (N, 3), R (M, 3, 3))`` and exceptions (``ValueError()``
for ragged input or a wrong number of focal lengths, ``numpy.linalg.LinAlgError`` from the Cholesky
factor of an indefinite metric).

Where the work is.  The reference centres the (2M x N) observation matrix and takes its FULL SVD
(``:21, :70, :154``: an N x N factor, O(N^2) memory, O(M N^2) time) to use three singular triplets.
Here ``ba_factorize_centred_rank4`` does both on the GPU: centroids by a fixed-order column sum, the
Gram matrix W W^T (2M x 2M) on the FP64 tensor cores with the Schur-product SYRK kernel, its leading
eigenspace, S = U^T W -- O(M N) memory, no CPU fallback.  Everything after that is O(M) algebra on
3-vectors (metric constraints, a 6 x 6 eigenproblem, per-image 3 x 3 polar decompositions) and stays
in NumPy on the host, written on the 6-vector form of symmetric 3 x 3 matrices instead of the
reference's 3^4-entry tensor loops.

Signs.  Singular vectors are defined up to sign, and the reference's result depends on the signs
LAPACK happens to return: with U -> U D, D = diag(+-1), the outputs change to (S D, D R) when
det D = +1 (a half-turn of the world frame) but to the *mirror* solution of the affine ambiguity
when det D = -1 (the cross product at ``:323`` changes sign against the other term).  The kernels
make each singular vector's largest-magnitude entry positive; ``signs=(d0, d1, d2)`` applies another
choice, so either of the reference's two possible answers can be reproduced
(``tests/test_affine_calibration.py`` compares against the unmodified reference for all eight).
"""
from __future__ import annotations

import numpy as np

from . import _cabi
from .projective_depth import MAX_IMAGES, _default_device

_SQ2 = np.sqrt(2.0)
_MODELS = ("orthographic", "symmetric_affine", "paraperspective")


# ---- GPU stage -------------------------------------------------------------------------------------
def factorize_observations(data_list, device: int | None = None):
    """``(U3 (2M, 3), S3 (3, N) = diag(sigma) V^T, sigma (3,), t (M, 2))`` of the centred observation
    matrix (``:224-240`` + the SVD).  ``data_list``: M arrays (N, 2)."""
    lengths = [len(x) for x in data_list]
    if lengths.count(lengths[0]) != len(lengths):
        raise ValueError()
    Wt = np.ascontiguousarray(np.hstack(data_list), dtype=np.float64)   # (N, 2M): point-major
    n_cols, n_rows = Wt.shape
    if not (2 <= len(data_list) <= MAX_IMAGES) or n_cols < 4:
        raise ValueError(f"the GPU factorisation takes 2..{MAX_IMAGES} images and >= 4 points")
    U4 = np.empty((n_rows, 4), dtype=np.float64)
    S4 = np.empty((4, n_cols), dtype=np.float64)
    sigma = np.empty(4, dtype=np.float64)
    mean = np.empty(n_rows, dtype=np.float64)
    lib = _cabi.load()
    dev = _default_device() if device is None else int(device)
    _cabi.check(lib.ba_factorize_centred_rank4(dev, n_cols, n_rows, Wt.ctypes.data, mean.ctypes.data,
                                               U4.ctypes.data, S4.ctypes.data, sigma.ctypes.data,
                                               _cabi.BA_MEM_HOST, None))
    U3, S3 = canonical_signs(U4[:, :3], S4[:3])
    return U3, S3, sigma[:3].copy(), mean.reshape(-1, 2)


def canonical_signs(U3, S3):
    """Largest-magnitude entry of every left singular vector positive."""
    U3 = np.array(U3, dtype=np.float64)
    S3 = np.array(S3, dtype=np.float64)
    pick = np.abs(U3).argmax(axis=0)
    d = np.where(U3[pick, np.arange(U3.shape[1])] < 0.0, -1.0, 1.0)
    return U3 * d[None, :], S3 * d[:, None]


# ---- host stage: O(M) algebra ------------------------------------------------------------------------
def _six(Q):
    """Symmetric (M, 3, 3) -> (M, 6): [Q00, Q11, Q22, sqrt2 Q12, sqrt2 Q20, sqrt2 Q01], the basis in
    which the reference's B (``:243-256``) is a plain sum of outer products."""
    return np.stack([Q[:, 0, 0], Q[:, 1, 1], Q[:, 2, 2], _SQ2 * Q[:, 1, 2], _SQ2 * Q[:, 2, 0], _SQ2 * Q[:, 0, 1]],
                    axis=1)


def metric_system(model: str, U3, t, f=None):
    """B (6 x 6) of the metric constraints: ``_create_B_cal`` + ``_get_B`` of the three models."""
    u0, u1 = U3[0::2], U3[1::2]
    p = _six(u0[:, :, None] * u0[:, None, :])
    q = _six(u1[:, :, None] * u1[:, None, :])
    s = _six(u0[:, :, None] * u1[:, None, :] + u1[:, :, None] * u0[:, None, :])
    if model == "orthographic":          # |u0|_T = |u1|_T = 1, u0.u1 = 0   (:27-38)
        return p.T @ p + q.T @ q + 0.25 * (s.T @ s)
    if model == "symmetric_affine":      # one quadratic form per image   (:74-116)
        a = t[:, 0] * t[:, 1]
        c = t[:, 0] ** 2 - t[:, 1] ** 2
        w = a[:, None] * (p - q) - 0.5 * c[:, None] * s
        return w.T @ w
    if model == "paraperspective":       # (:158-203)
        f = np.asarray(f, dtype=np.float64)
        al = 1.0 / (1.0 + t[:, 0] ** 2 / f ** 2)
        be = 1.0 / (1.0 + t[:, 1] ** 2 / f ** 2)
        ga = t[:, 0] * t[:, 1] / f ** 2
        g1 = ga ** 2 + 1.0
        pa, qb = al[:, None] * p, be[:, None] * q
        cross = pa.T @ (((ga ** 2 - 1.0) * be)[:, None] * q)
        mixed = (ga[:, None] * (pa + qb)).T @ s
        return (g1[:, None] * pa).T @ pa + (g1[:, None] * qb).T @ qb + s.T @ s + cross + cross.T - mixed - mixed.T
    raise ValueError(model)


def _metric(model: str, B):
    """tau -> T (``:43-49, :121-127, :208-214, :259-270``), made positive by the determinant rule."""
    if model == "orthographic":
        tau = np.linalg.solve(B, np.array([1.0, 1.0, 1.0, 0.0, 0.0, 0.0]))
    else:
        lam, vec = np.linalg.eigh(0.5 * (B + B.T))
        tau = vec[:, 0]
    T = np.array([[tau[0], tau[5] / _SQ2, tau[4] / _SQ2],
                  [tau[5] / _SQ2, tau[1], tau[3] / _SQ2],
                  [tau[4] / _SQ2, tau[3] / _SQ2, tau[2]]])
    return -T if np.linalg.det(T) < 0 else T


def camera_rotations(Mm, U3, T, t):
    """``_get_zeta_beta_g`` + ``_compute_rotation_mat`` (``:273-341``)."""
    u0, u1 = U3[0::2], U3[1::2]
    q00 = np.einsum("ni,ij,nj->n", u0, T, u0)
    q01 = np.einsum("ni,ij,nj->n", u0, T, u1)
    q11 = np.einsum("ni,ij,nj->n", u1, T, u1)
    # least squares for (1/zeta^2, b = beta^2) through the pseudo-inverse (minimum-norm answer when
    # tx = ty = 0).  The reference pairs its rows [1, tx^2], [1, ty^2], [0, tx ty] (:276-279) with the
    # right-hand sides in the order (q00, q01, q11) (:284-288), not (q00, q11, q01): kept as it is.
    P = np.zeros((len(t), 3, 2))
    P[:, 0, 0] = P[:, 1, 0] = 1.0
    P[:, 0, 1] = t[:, 0] ** 2
    P[:, 1, 1] = t[:, 1] ** 2
    P[:, 2, 1] = t[:, 0] * t[:, 1]
    sol = np.einsum("nij,nj->ni", np.linalg.pinv(P), np.stack([q00, q01, q11], axis=1))
    zi, b2 = sol[:, 0], np.maximum(sol[:, 1], 0.0)
    at_centre = (np.abs(t) < 1e-8).all(axis=1)
    b2 = np.where(at_centre, 0.0, b2)
    zi = np.where(at_centre, 0.5 * (q00 + q11), zi)
    zi = np.where(zi <= 0.0, 1e8, zi)
    zeta, beta = np.sqrt(1.0 / zi), np.sqrt(b2)
    g = zeta[:, None] * t
    m0, m1 = Mm[0::2], Mm[1::2]
    # :328 divides EVERY image's third axis by 1 + beta_n^2 |g_0|^2 -- image 0's |g|^2 -- kept as is
    r3 = (zeta[:, None] * np.cross(m0, m1) - beta[:, None] * (g[:, :1] * m0 + g[:, 1:] * m1)) \
        / (1.0 + beta[:, None] ** 2 * float(g[0] @ g[0]))
    r1 = zeta[:, None] * m0 + (beta * g[:, 0])[:, None] * r3
    r2 = zeta[:, None] * m1 + (beta * g[:, 1])[:, None] * r3
    Rraw = np.stack([r1, r2, r3], axis=2)
    Uu, _, Vt = np.linalg.svd(Rraw)          # nearest orthogonal matrix per image (3 x 3 each)
    return Uu @ Vt


def calibrate_from_factors(model: str, U3, S3, t, f=None, signs=None):
    """The part after the factorisation: metric upgrade, shape, rotations."""
    if model not in _MODELS:
        raise ValueError()
    U3 = np.asarray(U3, dtype=np.float64)
    S3 = np.asarray(S3, dtype=np.float64)
    if signs is not None:
        d = np.asarray(signs, dtype=np.float64).reshape(3)
        U3, S3 = U3 * d[None, :], S3 * d[:, None]
    T = _metric(model, metric_system(model, U3, t, f))
    A = np.linalg.cholesky(T)
    S = np.linalg.solve(A, S3)               # A^-1 diag(Sigma) V^T   (:52, :130, :217)
    R = camera_rotations(U3 @ A, U3, T, t)
    return S.T, R


def _run(model, data_list, f=None, signs=None, device=None):
    U3, S3, _, t = factorize_observations(data_list, device)
    return calibrate_from_factors(model, U3, S3, t, f, signs)


def orthographic_self_calibration(data_list, signs=None, device=None):
    """Reference signature (``:7-9``) plus ``signs`` / ``device``."""
    return _run("orthographic", data_list, None, signs, device)


def symmetric_affine_self_calibration(data_list, signs=None, device=None):
    """Reference signature (``:59-61``) plus ``signs`` / ``device``."""
    return _run("symmetric_affine", data_list, None, signs, device)


def paraperspective_self_calibration(data_list, f, signs=None, device=None):
    """Reference signature (``:137-139``) plus ``signs`` / ``device``."""
    if len(data_list) != len(f):
        raise ValueError()
    return _run("paraperspective", data_list, np.asarray(f, dtype=np.float64), signs, device)
