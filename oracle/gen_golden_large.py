"""Golden vectors for the LARGE configurations (run in the build container only).

    python oracle/gen_golden_large.py [c2_reference] [c4_shape] [c5_shape]

The scenes come from the seeded generator of the package (``scenes.make_scene``), so the fixtures
hold only what a run produced, not its inputs:

``c2_reference.npz``
    The UNMODIFIED reference (``/root/reference/lib/bundle_adjustment.py``, imported as-is) on
    the full config 2 (50 cameras x 10 000 points, dense; ~16 GB, ~20 s per iteration): two LM
    iterations.  Pins the oracle at C2, the largest size the reference can hold (BASELINE.md
    section 3.3) -- ``tests/test_oracle_golden.py::test_oracle_matches_reference_at_c2``.
``c4_shape.npz``
    Config 4's shape (1000 cameras, 10 % visibility, n = 8993) at 3000 points: ten LM iterations of
    the oracle (sparsity-aware Schur restatement, Cholesky solve).  The reference itself cannot run
    a 1000-camera scene of any useful size ((N, n, n) = 647 MB per point), so this trajectory is the
    oracle's; the oracle's sparse path is checked against its dense path, which the reference pins.
``c5_shape.npz``
    Config 5's shape (as above plus 1 % outliers, plain L2 cost) at 5000 points: the oracle run to
    convergence (``optimize(2.0, 1e-8, max_iter=50)``) -- cost trajectory, damping, inner solves and
    the converged state, i.e. the "same converged reprojection RMS" check of BASELINE.json.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import time

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

import ba_b200  # noqa: E402
from oracle import ba_oracle as O  # noqa: E402

C4_SHAPE = dict(n_cams=1000, n_points=3000, seed=41, visibility=0.1)
C5_SHAPE = dict(n_cams=1000, n_points=5000, seed=51, visibility=0.1, outlier_frac=0.01)
SUB = 10  # every SUB-th point of the final state is stored


def obs_of(sc) -> O.ObsList:
    return O.ObsList(sc.n_points, sc.n_cams, np.repeat(np.arange(sc.n_points), np.diff(sc.obs_ptr)),
                     sc.obs_cam.astype(np.int64), sc.obs_xy, sc.obs_ptr)


def gen_c2_reference():
    sys.path.insert(0, REF)
    from lib.bundle_adjustment import BundleAdjuster  # the real reference class

    sc = ba_b200.scenes.make_scene(**ba_b200.scenes.CONFIGS["c2"])
    x, _ = sc.dense_x()
    t0 = time.time()
    ba = BundleAdjuster(x, sc.X0, sc.K0, sc.R0, sc.t0, f0=sc.f0, axis=sc.axis)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        X, K, R, t = ba.optimize(2.0, 1e-8, max_iter=2, is_debug=True)
    E = np.array([d["reprojection_error"] for d in ba.get_log()])
    print(f"c2_reference: {time.time() - t0:.1f} s, E = {E}")
    np.savez_compressed(os.path.join(OUT, "c2_reference.npz"), E=E, X_sub=X[::SUB], K=K, R=R, t=t,
                        X_sum=X.sum(axis=0), stdout=np.array(buf.getvalue()), sub=SUB)


def _oracle_run(cfg, max_iter, tol):
    sc = ba_b200.scenes.make_scene(**cfg)
    ora = O.OracleBundleAdjuster(None, sc.X0, sc.K0, sc.R0, sc.t0, f0=sc.f0, axis=sc.axis, obs=obs_of(sc))
    t0 = time.time()
    X, K, R, t = ora.optimize(2.0, tol, max_iter=max_iter, verbose=False, schur="sparse", solver="cholesky")
    E = np.array([r["E"] for r in ora.trace])
    c = np.array([np.nan if r["c"] is None else r["c"] for r in ora.trace])
    solves = np.array([r["solves"] for r in ora.trace])
    print(f"{cfg}: {time.time() - t0:.1f} s, {len(E) - 1} iterations, {solves.sum()} solves, "
          f"rms {np.sqrt(E[-1] / sc.nobs):.6g}")
    return dict(E=E, c=c, solves=solves, X_sub=X[::SUB], K=K, R=R, t=t, X_sum=X.sum(axis=0),
                nobs=sc.nobs, sub=SUB, rms=np.sqrt(E[-1] / sc.nobs))


def gen_c4_shape():
    np.savez_compressed(os.path.join(OUT, "c4_shape.npz"), **_oracle_run(C4_SHAPE, 10, -1.0))


def gen_c5_shape():
    np.savez_compressed(os.path.join(OUT, "c5_shape.npz"), **_oracle_run(C5_SHAPE, 50, 1e-8))


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2_reference", "c4_shape", "c5_shape"]
    for w in which:
        {"c2_reference": gen_c2_reference, "c4_shape": gen_c4_shape, "c5_shape": gen_c5_shape}[w]()
