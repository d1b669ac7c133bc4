"""Round-2 diagnostic: intermediates of one pass of the CUDA dual method against NumPy."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ba_b200  # noqa: E402

cabi = ba_b200.submodule("_cabi")
lib = cabi.load()
g = np.load(os.path.join(ROOT, "tests", "golden", "depth_dual.npz"))
x = np.ascontiguousarray(g["x"])
N, M = x.shape[:2]
V4, R60, W12, e, sums, U4 = (np.zeros((N, 4)), np.zeros((M, 60)), np.zeros((M, 12)), np.zeros((N, M)),
                             np.zeros((2, M)), np.zeros((3 * M, 4)))
lib.ba_depth_dual_probe(*(a.ctypes.data for a in (V4, R60, W12, e, sums, U4)))
z, errs = ba_b200.projective_depth_dual(x, 1.0, 1e-2, 1)
lib.ba_depth_dual_probe(None, None, None, None, None, None)
print("errs", errs)
np.set_printoptions(precision=4, linewidth=200, suppress=True)
print("W12[0]", W12[0])
print("W12[1]", W12[1])
W = x.copy()
norm2 = (W * W).sum(axis=(0, 2))
Wn = (W / norm2[None, :, None]).reshape(N, 3 * M)
G = Wn.T @ Wn
print("U4 orthonormal", np.abs(U4.T @ U4 - np.eye(4)).max())
vals, vecs = np.linalg.eigh(G)
Ut = vecs[:, np.argsort(vals)[::-1][:4]]
print("U4 spans the leading eigenspace", np.abs(U4 - Ut @ (Ut.T @ U4)).max())
print("V4 orthonormal", np.abs(V4.T @ V4 - np.eye(4)).max())
Vt = Wn @ Ut
Vt /= np.linalg.norm(Vt, axis=0)
print("V4 spans the right subspace", np.abs(V4 - Vt @ (Vt.T @ V4)).max())
xn = np.sqrt((x * x).sum(axis=2))
xh = x / xn[..., None]
idx4 = [(a, b) for a in range(4) for b in range(a, 4)]
idx3 = [(a, b) for a in range(3) for b in range(a, 3)]
for i in range(M):
    Ci = (V4[:, :, None] * xh[:, i, None, :]).reshape(N, 12)
    A = Ci.T @ Ci
    R = np.array([[A[3 * a0 + b0, 3 * a1 + b1] for (b0, b1) in idx3] for (a0, a1) in idx4]).ravel()
    wv, wvec = np.linalg.eigh(A)
    w = wvec[:, np.argmax(wv)]
    if i == 0:
        print("eigenvalues", wv[::-1][:6], "w", w, "diag A", np.diag(A))
    ee = Ci @ w
    print(i, "R60", np.abs(R60[i] - R).max(), "W12", min(np.abs(W12[i] - w).max(), np.abs(W12[i] + w).max()),
          "e", min(np.abs(e[:, i] - ee).max(), np.abs(e[:, i] + ee).max()),
          "sums", sums[0, i], (e[:, i] ** 2).sum(), sums[1, i], e[:, i].sum())
