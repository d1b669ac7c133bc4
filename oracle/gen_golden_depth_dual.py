"""Generate ``tests/golden/depth_dual.npz`` with the UNMODIFIED reference's dual projective-depth
method (``lib/perspective_camera_calibration.py:147-235``) on the scene of
``euclidiean_reconstruction.py`` (seed 123, tol 1e-2 as at ``:42``), and record what the reference's
own Euclidean upgrade makes of it (``perspective_self_calibration``, ``:513-540``).

Also asserts, at generation time, the fact the next round's parity definition rests on: flipping the
sign of any image's column of z (the one thing LAPACK's eigenvector convention decides, see
``oracle/depth_oracle.py``) does not change the upgraded (X, R, t, K).

    python oracle/gen_golden_depth_dual.py
"""
import contextlib
import io
import os
import sys

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import lib.perspective_camera_calibration as pcc  # noqa: E402
from lib.camera import Camera, calc_projected_points, get_camera_parames  # noqa: E402
from lib.utils import sample_hemisphere_points, set_points  # noqa: E402

np.random.seed(123)
pos = sample_hemisphere_points(10, 5)
targets = np.random.normal(0, 0.5, (10, 3))
cams = [Camera.create(p, t, 1.0, 1.0) for p, t in zip(pos, targets)]
K, R, t = get_camera_parames(cams)
X = set_points()
x_list = calc_projected_points(X, K, R, t)
for p in x_list:
    p += 0.005 * np.random.randn(*p.shape)

x = pcc._create_data_matrix(x_list, 1.0)
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    z = pcc._compute_projective_depth_dual_method(x.copy(), 1.0, 1e-2)
errs = np.array([float(l.split("=")[1]) for l in buf.getvalue().splitlines() if l.startswith("Iteration")])
print("dual: iterations", len(errs), "E", errs[:2], errs[-1], "column sums", np.sign(z.sum(axis=0)))


def upgrade(zz):
    """The reference's own tail of perspective_self_calibration (:528-540) for given depths."""
    W = x * zz[..., None]
    M, S = pcc.factorization_method(W.reshape(W.shape[0], -1).T)
    P = M.reshape(-1, 3, 4)
    with contextlib.redirect_stdout(io.StringIO()):
        H, Kk = pcc._euclidean_upgrading(P, 1.0)
        Xr, Rr, tr = pcc._reconstruct_3d(P, S, Kk, H)
        Xr, Rr, tr = pcc.correct_world_coordinates(Xr, Rr, tr, method="predict")
    return Xr, Rr, tr, Kk


base = upgrade(z)
flipped = z.copy()
flipped[:, [2, 5, 6]] *= -1.0
alt = upgrade(flipped)
worst = max(float(np.abs(a - b).max()) for a, b in zip(base, alt))
print("max change of (X, R, t, K) under per-image sign flips of z:", worst)
assert worst < 1e-8, "the Euclidean upgrade is expected to be invariant to per-image signs of z"
# a longer run (the script's tolerance is met after one pass): 15 passes
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    z15 = pcc._compute_projective_depth_dual_method(x.copy(), 1.0, 1e-9, 15)
errs15 = np.array([float(l.split("=")[1]) for l in buf.getvalue().splitlines() if l.startswith("Iteration")])
print("dual, 15 passes: E", errs15[0], errs15[-1])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "depth_dual.npz"), xy=np.stack(x_list), x=x, z=z, E=errs,
                    z15=z15, E15=errs15, X=base[0], R=base[1], t=base[2], K=base[3], f0=np.array(1.0),
                    tol=np.array(1e-2))
