"""CPU restatement of the reference's projective-depth iteration, primary method
(``lib/perspective_camera_calibration.py:61-144``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/`` as the checker of the CUDA path, never by the
product.  Pinned against outputs of the unmodified reference (``oracle/gen_golden_depth.py`` ->
``tests/golden/depth_primary.npz``).

The reference, per iteration (z starts at 1):
  :86-90    W = x * z, every point's 3M-vector scaled to unit length
  :92-96    SVD of the (3M x N) matrix, U_ = its four leading left singular vectors
  :98-113   per point j the M x M matrix A_j[i][i'] = sum_k (x_ij . u_ik)(x_i'j . u_i'k) / (|x_ij| |x_i'j|)
  :115-131  xi_j = unit eigenvector of the largest eigenvalue of A_j, sign chosen so that its
            entries sum to >= 0;  z_ij = xi_j[i] / |x_ij|
  :133-136  M = U_, S = diag(Sigma[:4]) Vt[:4];  E = f0 sqrt(mean |x - normalise(M S)|^2)  (:44-58)
  :138-142  stop when E < tolerance or after max_iter iterations

This restatement uses two identities instead of the N eigenproblems of size M and the SVD:
  * A_j = B_j B_j^T with B_j (M x 4), B_j[i][k] = (x_ij . u_ik) / |x_ij|, so its leading eigenvector
    is B_j v / |B_j v| with v the leading eigenvector of the 4 x 4 matrix B_j^T B_j;
  * the four leading left singular vectors of W span the leading eigenspace of the Gram matrix
    G = W W^T (3M x 3M), and M S = U_ U_^T W; everything above depends on U_ only through the
    projector U_ U_^T, so any orthonormal basis of that eigenspace gives the same z and E.
"""
from __future__ import annotations

import numpy as np


def create_data_matrix(x_list, f0: float) -> np.ndarray:
    """(n_points, n_images, 3) rows (x / f0, y / f0, 1)  (:35-41)."""
    x = np.stack([np.concatenate((np.asarray(p, dtype=np.float64) / f0, np.ones((len(p), 1))), axis=1)
                  for p in x_list])
    return np.ascontiguousarray(x.transpose(1, 0, 2))


def leading_subspace(Wn: np.ndarray) -> np.ndarray:
    """Orthonormal basis (3M x 4) of the four leading left singular vectors of Wn^T (N x 3M rows)."""
    G = Wn.T @ Wn
    vals, vecs = np.linalg.eigh(G)
    return vecs[:, np.argsort(vals)[::-1][:4]]


def depth_iteration(x: np.ndarray, z: np.ndarray, f0: float):
    """One pass of :86-136: returns (new z, reprojection error E of this pass)."""
    N, M = z.shape
    W = x * z[..., None]
    Wn = (W / np.sqrt((W * W).sum(axis=(1, 2)))[:, None, None]).reshape(N, 3 * M)
    U4 = leading_subspace(Wn)                        # (3M, 4)
    Uc = U4.reshape(M, 3, 4)
    xn = np.sqrt((x * x).sum(axis=2))                # (N, M)
    B = np.einsum("jic,ick->jik", x, Uc) / xn[..., None]   # (N, M, 4)
    C = np.einsum("jik,jil->jkl", B, B)              # (N, 4, 4)
    vals, vecs = np.linalg.eigh(C)
    v = vecs[np.arange(N), :, np.argmax(vals, axis=1)]     # (N, 4)
    xi = np.einsum("jik,jk->ji", B, v)
    xi /= np.sqrt((xi * xi).sum(axis=1))[:, None]
    xi[xi.sum(axis=1) < 0] *= -1.0
    z_new = xi / xn
    # reprojection error of this pass: M S = U4 U4^T Wn^T  (:133-136, :44-58)
    c = Wn @ U4                                      # (N, 4)
    PX = np.einsum("ick,jk->jic", Uc, c)             # (N, M, 3)
    PX = PX / PX[..., 2:3]
    d = x - PX
    E = f0 * np.sqrt((d * d).sum(axis=2).mean())
    return z_new, float(E)


def projective_depth_primary(x: np.ndarray, f0: float, tolerance: float, max_iter: int = 200):
    """z (n_points, n_images) and the list of per-iteration errors  (:61-144)."""
    x = np.asarray(x, dtype=np.float64)
    z = np.ones(x.shape[:2])
    errors = []
    while True:
        z, E = depth_iteration(x, z, f0)
        errors.append(E)
        if E < tolerance or len(errors) >= max_iter:
            break
    return z, errors


# ---- dual method (reference :147-235) ---------------------------------------------------------------
def dual_iteration(x: np.ndarray, z: np.ndarray, f0: float):
    """One pass of the dual method (:163-222): returns (new z, E of this pass).

    The reference normalises W per IMAGE (:171-176: every image's 3 x N block divided by its squared
    Frobenius norm), takes the four leading RIGHT singular vectors V_ (N x 4, :178-181) and solves,
    per image i, the N x N eigenproblem of B_i[j][l] = (v_j . v_l)(x_ij . x_il) / (|x_ij| |x_il|)
    (:183-204).  B_i = C_i C_i^T with the N x 12 matrix C_i[j][(a, b)] = v_j[a] x_ij[b] / |x_ij|, so
    its leading eigenvector is C_i w / |C_i w| with w from the 12 x 12 matrix C_i^T C_i.

    Sign: the reference takes LAPACK's eigenvector as is and then flips ROWS of the (N, M) array whose
    sum is negative (:212-215) -- a rule carried over from the primary method that does not fix the
    per-image sign.  The sign of column i of z is therefore LAPACK's choice; it cancels in every later
    pass (V_ does not depend on it) and in the Euclidean upgrade.  This restatement fixes the
    per-image sign by making each column's sum non-negative BEFORE applying the reference's row rule;
    callers compare with the reference up to a sign per image.
    """
    N, M = z.shape
    W = x * z[..., None]                                        # (N, M, 3)
    per_image = (W * W).sum(axis=(0, 2))                        # squared Frobenius norm of image i's block
    Wn = (W / per_image[None, :, None]).reshape(N, 3 * M)       # (:171-176)
    # leading right singular vectors of the (3M x N) matrix = leading eigenvectors of Wn Wn^T (N x N)
    # = Wn U4 Sigma^-1 with U4 from the small Gram matrix Wn^T Wn (3M x 3M)
    G = Wn.T @ Wn
    vals, vecs = np.linalg.eigh(G)
    order = np.argsort(vals)[::-1][:4]
    U4, sig = vecs[:, order], np.sqrt(vals[order])
    V4 = (Wn @ U4) / sig                                        # (N, 4), orthonormal columns
    xn = np.sqrt((x * x).sum(axis=2))                           # (N, M)
    xh = x / xn[..., None]
    xi = np.empty((N, M))
    for i in range(M):
        Ci = (V4[:, :, None] * xh[:, i, None, :]).reshape(N, 12)
        wv, wvec = np.linalg.eigh(Ci.T @ Ci)
        e = Ci @ wvec[:, np.argmax(wv)]
        e /= np.sqrt((e * e).sum())
        xi[:, i] = e if e.sum() >= 0 else -e
    xi[xi.sum(axis=1) < 0] *= -1.0                              # (:212-215), verbatim
    z_new = xi / xn
    # E (:219-221): M = U[:, :4], S = diag(Sigma[:4]) V_^T  ->  M S = U4 U4^T Wn^T
    c = Wn @ U4
    Uc = U4.reshape(M, 3, 4)
    PX = np.einsum("ick,jk->jic", Uc, c)
    PX = PX / PX[..., 2:3]
    d = x - PX
    return z_new, float(f0 * np.sqrt((d * d).sum(axis=2).mean()))


def projective_depth_dual(x: np.ndarray, f0: float, tolerance: float, max_iter: int = 50):
    """(:147-235) up to a sign per image (see dual_iteration)."""
    x = np.asarray(x, dtype=np.float64)
    z = np.ones(x.shape[:2])
    errors = []
    while True:
        z, E = dual_iteration(x, z, f0)
        errors.append(E)
        if E < tolerance or len(errors) >= max_iter:
            break
    return z, errors
