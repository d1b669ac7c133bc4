"""CPU oracle of the affine self-calibrations -- TEST INFRASTRUCTURE ONLY (tests/, the generator
scripts); the product (`3d-reconstruction-from-multi-view-exp_b200/affine_calibration.py`) never
imports it.

Restates reference lib/affine_camera_calibration.py in vectorised NumPy, keeping LAPACK's full
SVD (np.linalg.svd), i.e. LAPACK's sign convention for the singular vectors, so that it can be
pinned bit-for-bit-ish (1e-12) against outputs of the unmodified reference
(tests/golden/affine_calib.npz, written by oracle/gen_golden_affine.py):

  observation_matrix          :224-240   W = hstack(data_list).T, rows centred, t = centroids
  metric_tensor               :23-40 (orthographic), :72-118 (symmetric affine), :156-205
                              (paraperspective): the 3x3x3x3 tensor B_cal as sums of outer products
  reduce_metric_tensor        :243-256   B (6 x 6)
  metric_from_tau             :259-270   T (3 x 3)
  zeta_beta_g                 :273-313
  rotations                   :316-341   incl. the reference's `[0]` at :328 (the FIRST image's
                              |g|^2 enters every image's denominator) and the polar projection
  self_calibration            :7-56, :59-134, :137-221   the three entry points, `signs` = an
                              optional diagonal +-1 applied to the three singular vectors (what a
                              different SVD implementation may legitimately return)
"""
import numpy as np

SQ2 = np.sqrt(2.0)


def observation_matrix(data_list):
    """:224-240.  W (2M, N) centred per row, t (M, 2)."""
    n = [len(x) for x in data_list]
    if n.count(n[0]) != len(n):
        raise ValueError()
    W = np.hstack(data_list).T.astype(np.float64)
    t = W.mean(axis=1)[:, None]
    return W - t, t.reshape(-1, 2)


def _outer4(A, B):
    """sum_n A[n,i,j] B[n,k,l]"""
    return np.einsum("nij,nkl->ijkl", A, B)


def metric_tensor(model, U3, t, f=None):
    """B_cal (3,3,3,3) for model in {"orthographic", "symmetric_affine", "paraperspective"}."""
    u0, u1 = U3[0::2], U3[1::2]                       # (M, 3) each: the two rows of every image
    P00 = np.einsum("ni,nj->nij", u0, u0)
    P11 = np.einsum("ni,nj->nij", u1, u1)
    P01 = np.einsum("ni,nj->nij", u0, u1)
    P10 = np.einsum("ni,nj->nij", u1, u0)
    Ps = P01 + P10
    if model == "orthographic":                       # :27-38
        return _outer4(P00, P00) + _outer4(P11, P11) + 0.25 * _outer4(Ps, Ps)
    if model == "symmetric_affine":                   # :74-116
        a = t.prod(axis=1)
        c = t[:, 0] ** 2 - t[:, 1] ** 2
        a2 = (a ** 2)[:, None, None]
        c2 = (c ** 2)[:, None, None]
        ac = (a * c)[:, None, None]
        out = _outer4(a2 * P00, P00) + _outer4(a2 * P11, P11) - _outer4(a2 * P00, P11) - _outer4(a2 * P11, P00)
        # :90-97: (u0i u1j + u1i u0j)(u0k u1l + u1k u0l) written out term by term
        out += 0.25 * (_outer4(c2 * P01, P01) + _outer4(c2 * P10, P01) + _outer4(c2 * P01, P10) + _outer4(c2 * P10, P10))
        out -= 0.5 * (_outer4(ac * P00, P01) + _outer4(ac * P00, P10) + _outer4(ac * P01, P00) + _outer4(ac * P10, P00)
                      - _outer4(ac * P01, P11) - _outer4(ac * P10, P11) - _outer4(ac * P11, P01) - _outer4(ac * P11, P10))
        return out
    if model == "paraperspective":                    # :158-203
        alpha = 1 / (1 + t[:, 0] ** 2 / f ** 2)
        beta = 1 / (1 + t[:, 1] ** 2 / f ** 2)
        gamma = t.prod(axis=1) / f ** 2
        w = lambda v: v[:, None, None]
        out = _outer4(w((gamma ** 2 + 1) * alpha ** 2) * P00, P00) + _outer4(w((gamma ** 2 + 1) * beta ** 2) * P11, P11)
        out += _outer4(P01, P01) + _outer4(P01, P10) + _outer4(P10, P01) + _outer4(P10, P10)
        out -= _outer4(w(alpha * gamma) * P00, P01) + _outer4(w(alpha * gamma) * P00, P10) \
            + _outer4(w(alpha * gamma) * P01, P00) + _outer4(w(alpha * gamma) * P10, P00)
        out -= _outer4(w(beta * gamma) * P11, P01) + _outer4(w(beta * gamma) * P11, P10) \
            + _outer4(w(beta * gamma) * P01, P11) + _outer4(w(beta * gamma) * P10, P11)
        out += _outer4(w((gamma ** 2 - 1) * alpha * beta) * P00, P11) + _outer4(w((gamma ** 2 - 1) * alpha * beta) * P11, P00)
        return out
    raise ValueError(model)


def reduce_metric_tensor(Bc):
    """:243-256."""
    B = np.zeros((6, 6))
    for i in range(3):
        for j in range(3):
            i1, i2, j1, j2 = (i + 1) % 3, (i + 2) % 3, (j + 1) % 3, (j + 2) % 3
            B[i, j] = Bc[i, i, j, j]
            B[i, 3 + j] = SQ2 * Bc[i, i, j1, j2]
            B[3 + i, j] = SQ2 * Bc[i1, i2, j, j]
            B[3 + i, 3 + j] = 2 * Bc[i1, i2, j1, j2]
    return B


def metric_from_tau(tau):
    """:259-270."""
    return np.array([[tau[0], tau[5] / SQ2, tau[4] / SQ2],
                     [tau[5] / SQ2, tau[1], tau[3] / SQ2],
                     [tau[4] / SQ2, tau[3] / SQ2, tau[2]]])


def zeta_beta_g(U3, T, t):
    """:273-313."""
    M = t.shape[0]
    P = np.ones((M, 3, 2))
    P[:, :2, 1] = t ** 2
    P[:, 2, 0] = 0.0
    P[:, 2, 1] = t.prod(axis=1)
    U1, U2 = U3[::2], U3[1::2]
    Q = np.stack([np.einsum("ni,ij,nj->n", U1, T, U1), np.einsum("ni,ij,nj->n", U1, T, U2),
                  np.einsum("ni,ij,nj->n", U2, T, U2)], axis=1)
    sol = (np.linalg.pinv(P) @ Q[..., None])[..., 0]
    zeta2_inv, beta2 = sol[:, 0].copy(), sol[:, 1].copy()
    beta2[beta2 < 0.0] = 0.0
    centred = (np.abs(t) < 1e-8).all(axis=1)
    beta2[centred] = 0.0
    zeta2_inv[centred] = ((Q[:, 0] + Q[:, 2]) / 2)[centred]
    zeta2_inv[zeta2_inv <= 0.0] = 1e8
    zeta = np.sqrt(1 / zeta2_inv)
    beta = np.sqrt(beta2)
    return zeta, beta, zeta[:, None] * t


def rotations(Mm, U3, T, t):
    """:316-341."""
    zeta, beta, g = zeta_beta_g(U3, T, t)
    m0, m1 = Mm[::2], Mm[1::2]
    num = zeta[:, None] * np.cross(m0, m1) - beta[:, None] * (g[:, :1] * m0 + g[:, 1:] * m1)
    den = 1 + beta[:, None] ** 2 * (g[0] @ g[0])     # :328: the reference indexes image 0 here
    r3 = num / den
    r1 = zeta[:, None] * m0 + (beta * g[:, 0])[:, None] * r3
    r2 = zeta[:, None] * m1 + (beta * g[:, 1])[:, None] * r3
    R = np.stack([r1, r2, r3], axis=2)               # columns r1, r2, r3
    U, _, Vt = np.linalg.svd(R)
    return U @ Vt


def self_calibration(model, data_list, f=None, signs=None, return_parts=False):
    """The three entry points (:7, :59, :137): (S.T (N, 3), R (M, 3, 3))."""
    if model == "paraperspective" and len(data_list) != len(f):
        raise ValueError()
    W, t = observation_matrix(data_list)
    U, Sigma, Vt = np.linalg.svd(W, full_matrices=False)
    U3, V3 = U[:, :3].copy(), Vt[:3].copy()
    if signs is not None:
        d = np.asarray(signs, dtype=np.float64)
        U3 *= d[None, :]
        V3 *= d[:, None]
    B = reduce_metric_tensor(metric_tensor(model, U3, t, f))
    if model == "orthographic":
        tau = np.linalg.solve(B, np.array([1.0, 1, 1, 0, 0, 0]))     # :43
    else:
        L, Pm = np.linalg.eig(B)                                     # :121-122, :208-209
        tau = Pm[:, np.argmin(L)]
    T = metric_from_tau(tau)
    if np.linalg.det(T) < 0:
        T = -T
    A = np.linalg.cholesky(T)
    Mm = U3 @ A
    S = np.linalg.inv(A) @ np.diag(Sigma[:3]) @ V3
    R = rotations(Mm, U3, T, t)
    if return_parts:
        return S.T, R, {"U3": U3, "sigma": Sigma[:3], "t": t, "B": B, "T": T}
    return S.T, R
