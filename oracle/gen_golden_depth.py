"""Generate ``tests/golden/depth_primary.npz`` with the UNMODIFIED reference's
``_compute_projective_depth_primary_method`` (build container only).

    python oracle/gen_golden_depth.py

Scene: the recipe of ``euclidiean_reconstruction.py:14-40`` (seed 123, 10 cameras on the radius-5
hemisphere, ``set_points()``, 0.005 noise); the tolerance is lowered from the script's 1e-2 so that
the fixture holds several iterations.  A second, larger random scene exercises ragged sizes.
"""
import contextlib
import io
import os
import sys

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)

from lib.camera import Camera, calc_projected_points, get_camera_parames  # noqa: E402
from lib.perspective_camera_calibration import (_compute_projective_depth_primary_method,  # noqa: E402
                                                _create_data_matrix)
from lib.utils import sample_hemisphere_points, set_points  # noqa: E402


def run(x_list, f0, tol, max_iter):
    x = _create_data_matrix(x_list, f0)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        z = _compute_projective_depth_primary_method(x.copy(), f0, tol, max_iter)
    errs = [float(l.split("=")[1]) for l in buf.getvalue().splitlines() if l.startswith("Iteration")]
    return x, z, np.array(errs), buf.getvalue()


out = {}
# case A: the script's scene
np.random.seed(123)
pos = sample_hemisphere_points(10, 5)
targets = np.random.normal(0, 0.5, (10, 3))
cams = [Camera.create(p, t, 1.0, 1.0) for p, t in zip(pos, targets)]
K, R, t = get_camera_parames(cams)
X = set_points()
x_list = calc_projected_points(X, K, R, t)
for p in x_list:
    p += 0.005 * np.random.randn(*p.shape)
x, z, errs, text = run(x_list, 1.0, 6e-3, 200)
out.update(a_xy=np.stack(x_list), a_x=x, a_z=z, a_E=errs, a_f0=np.array(1.0), a_tol=np.array(6e-3),
           a_stdout=np.array(text))
print("case a:", x.shape, len(errs), errs[:3], errs[-1])

# case B: 7 cameras, 333 points in [-1, 1]^3, f0 = 2
rng = np.random.RandomState(9)
np.random.seed(9)
pos = sample_hemisphere_points(7, 5)
cams = [Camera.create(p, rng.normal(0, 0.5, 3), 1.2, 2.0) for p in pos]
K, R, t = get_camera_parames(cams)
X = rng.uniform(-1, 1, (333, 3))
x_list = calc_projected_points(X, K, R, t)
for p in x_list:
    p += 0.002 * rng.randn(*p.shape)
x, z, errs, text = run(x_list, 2.0, 1e-9, 12)   # never reaches the tolerance: stops at max_iter
out.update(b_xy=np.stack(x_list), b_x=x, b_z=z, b_E=errs, b_f0=np.array(2.0), b_tol=np.array(1e-9),
           b_max_iter=np.array(12), b_stdout=np.array(text))
print("case b:", x.shape, len(errs), errs[:3], errs[-1])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "depth_primary.npz"), **out)
