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
