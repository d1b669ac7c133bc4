"""Generate ``tests/golden/projection.npz`` with the UNMODIFIED reference's
``lib.camera.calc_projected_points`` (build container only; ``/root/reference`` does not exist on
the GPU box).

    python oracle/gen_golden_projection.py
"""
import os
import sys

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)

from lib.camera import Camera, calc_projected_points, get_camera_parames  # noqa: E402
from lib.utils import sample_hemisphere_points  # noqa: E402

rng = np.random.RandomState(7)
np.random.seed(7)
M, N = 9, 157
cams = [Camera.create(p, rng.normal(0, 0.5, 3), f=1.0 + 0.1 * i, f0=1.0)
        for i, p in enumerate(sample_hemisphere_points(M, 5))]
K, R, t = get_camera_parames(cams)
K = K + rng.normal(0, 0.01, K.shape)  # general 3x3 intrinsics (skew, K[2, :] != (0, 0, 1)) like after BA
X = rng.uniform(-1, 1, (N, 3))
x = np.stack(calc_projected_points(X, K, R, t))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "projection.npz"), X=X, K=K, R=R, t=t, x=x)
print("projection.npz", x.shape)
