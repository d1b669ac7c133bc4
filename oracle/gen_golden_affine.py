"""Generate ``tests/golden/affine_calib.npz`` with the UNMODIFIED reference's three affine
self-calibrations (``lib/affine_camera_calibration.py:7, :59, :137``):

  * ``script``: the scene of ``affine_reconstruction.py`` (seed 123, 12 cameras, noise 0.005, f = 1),
    the paraperspective call the script makes (``:42``) and the other two models on the same data;
  * ``wide``: 7 cameras x 300 random points, image centroids well away from the principal point.

Stored per case and model: S (N, 3), R (M, 3, 3); plus the data.  At generation time the script
also (a) checks ``oracle/affine_oracle.py`` against these outputs, and (b) records what the
reference returns when the three singular vectors come back with other signs (what another SVD
may legitimately give): the unmodified functions are re-run with ``np.linalg.svd`` wrapped so
that U[:, k] and Vt[k] are multiplied by d[k] -- the sign classes the GPU path is compared
against (``tests/test_affine_calibration.py``).

    python oracle/gen_golden_affine.py
"""
import itertools
import os
import sys

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import lib.affine_camera_calibration as acc  # noqa: E402
from lib.camera import Camera, calc_projected_points, get_camera_parames  # noqa: E402
from lib.utils import sample_hemisphere_points, set_points  # noqa: E402
from oracle import affine_oracle as AO  # noqa: E402


def script_scene():
    np.random.seed(123)
    f, n_images = 1.0, 12
    pos = sample_hemisphere_points(n_images, 5)
    targets = np.random.normal(0, 0.5, (n_images, 3))
    cams = [Camera.create(p, t, f) for p, t in zip(pos, targets)]
    K, R, t = get_camera_parames(cams)
    X = set_points()
    x_list = calc_projected_points(X, K, R, t)
    for x in x_list:
        x += 0.005 * np.random.randn(*x.shape)
    return x_list, f * np.ones(n_images)


def wide_scene():
    rng = np.random.default_rng(5)
    n_images, n_points = 7, 300
    X = rng.normal(0, 1.0, (n_points, 3))
    x_list = []
    for i in range(n_images):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] *= -1
        c = q.T @ X.T
        depth = 9.0 + c[2]
        x = (c[:2] / depth).T + rng.normal(0, 0.3, 2) + 0.001 * rng.normal(size=(n_points, 2))
        x_list.append(x)
    return x_list, 1.0 + 0.2 * rng.random(n_images)


MODELS = {
    "orthographic": lambda xl, f: acc.orthographic_self_calibration(xl),
    "symmetric_affine": lambda xl, f: acc.symmetric_affine_self_calibration(xl),
    "paraperspective": lambda xl, f: acc.paraperspective_self_calibration(xl, f),
}


class _SignedSvd:
    """np.linalg.svd with chosen signs of the three leading singular vectors (any valid SVD)."""

    def __init__(self, d):
        self.d = np.asarray(d, dtype=np.float64)
        self.real = np.linalg.svd

    def __call__(self, a, *args, **kw):
        out = self.real(a, *args, **kw)
        if a.ndim == 2 and a.shape[1] > 3:      # the observation matrix, not the 3 x 3 polar step
            U, s, Vt = out
            U = U.copy()
            Vt = Vt.copy()
            U[:, :3] *= self.d[None, :]
            Vt[:3] *= self.d[:, None]
            return U, s, Vt
        return out


def main():
    store = {}
    for case, (x_list, f) in (("script", script_scene()), ("wide", wide_scene())):
        store[f"{case}_xy"] = np.stack(x_list)
        store[f"{case}_f"] = f
        for model, fn in MODELS.items():
            S, R = fn([x.copy() for x in x_list], f)
            store[f"{case}_{model}_S"] = S
            store[f"{case}_{model}_R"] = R
            So, Ro = AO.self_calibration(model, [x.copy() for x in x_list], f)
            dS, dR = float(np.abs(So - S).max()), float(np.abs(Ro - R).max())
            print(f"{case:7s} {model:17s} oracle vs reference: S {dS:.2e}  R {dR:.2e}")
            assert dS < 1e-9 and dR < 1e-9
            # the reference's answer for each of the 8 sign choices of the singular vectors
            for d in itertools.product((1.0, -1.0), repeat=3):
                wrapped = _SignedSvd(d)
                np.linalg.svd = wrapped
                try:
                    Sd, Rd = fn([x.copy() for x in x_list], f)
                finally:
                    np.linalg.svd = wrapped.real
                tag = "".join("p" if v > 0 else "m" for v in d)
                store[f"{case}_{model}_S_{tag}"] = Sd
                store[f"{case}_{model}_R_{tag}"] = Rd
                Sod, Rod = AO.self_calibration(model, [x.copy() for x in x_list], f, signs=d)
                assert np.abs(Sod - Sd).max() < 1e-9 and np.abs(Rod - Rd).max() < 1e-9, (case, model, d)
            # proper sign changes (det d = +1) are a 180-degree turn of the world frame: S -> S d, R -> d R
            d = np.array([-1.0, -1.0, 1.0])
            assert np.abs(store[f"{case}_{model}_S_mmp"] - S * d[None, :]).max() < 1e-9
            assert np.abs(store[f"{case}_{model}_R_mmp"] - d[None, :, None] * R).max() < 1e-9
    out = os.path.join(ROOT, "tests", "golden", "affine_calib.npz")
    np.savez_compressed(out, **store)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
