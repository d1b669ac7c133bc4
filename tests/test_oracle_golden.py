"""The CPU oracle (oracle/ba_oracle.py) against outputs of the unmodified reference
(tests/golden/*.npz, produced by oracle/gen_golden.py).  CPU only."""
import io
import contextlib

import numpy as np
import pytest

from conftest import RUN_CASES, SMALL_CASES, case_inputs, load_golden
from oracle import ba_oracle as O


def _obs_dense_index(obs, M):
    return obs.pt * M + obs.cam


@pytest.mark.parametrize("name", SMALL_CASES)
def test_first_linearisation_matches_reference(name):
    g = load_golden(name)
    x, vis, X0, K0, R0, t0, axis, f0 = case_inputs(g)
    N, M = x.shape[:2]
    obs = O.ObsList.from_dense(x, vis)
    X, R, t = O.normalize_gauge(X0, R0, t0, axis)
    f, u = K0[:, 0, 0].copy(), K0[:, :2, 2].copy()
    np.testing.assert_allclose(X, g["lin_nX"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(R, g["lin_nR"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(t, g["lin_nt"], rtol=0, atol=1e-14)

    assert O.cost(obs, X, f, u, R, t, f0) == pytest.approx(float(g["lin_E0"]), rel=1e-13)
    lin = O.linearize(obs, X, f, u, R, t, f0)
    flat = _obs_dense_index(obs, M)
    np.testing.assert_allclose(lin.Jx, g["lin_Jx"].reshape(N * M, 2, 3)[flat], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(lin.Jc, g["lin_Jc"].reshape(N * M, 2, 9)[flat], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(lin.g_pt.ravel(), g["lin_d_P"], rtol=1e-11, atol=1e-13)
    removed, kept = O.gauge_indices(M, axis)
    np.testing.assert_allclose(lin.g_cam.ravel()[kept], g["lin_d_F"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(lin.V, g["lin_matE"], rtol=1e-11, atol=1e-12)
    # matF (N,3,9M-7): scatter the per-observation W blocks
    F = np.zeros((N, 3, 9 * M))
    for o in range(obs.nobs):
        F[obs.pt[o], :, 9 * obs.cam[o]: 9 * obs.cam[o] + 9] = lin.W[o]
    np.testing.assert_allclose(F[:, :, kept], g["lin_matF"], rtol=1e-11, atol=1e-12)
    G = np.zeros((9 * M, 9 * M))
    for i in range(M):
        G[9 * i: 9 * i + 9, 9 * i: 9 * i + 9] = lin.U[i]
    np.testing.assert_allclose(G[np.ix_(kept, kept)], g["lin_matG"], rtol=1e-11, atol=1e-11)

    c = float(g["lin_c"])
    dxi, dX, A, b = O.solve_damped(obs, lin, c, axis, chunk_points=7)
    scale = np.abs(g["lin_A"]).max()
    np.testing.assert_allclose(A, g["lin_A"], rtol=0, atol=1e-12 * scale)
    np.testing.assert_allclose(b, g["lin_b"], rtol=0, atol=1e-11 * np.abs(g["lin_b"]).max())
    np.testing.assert_allclose(dxi.ravel()[kept], g["lin_dxi"], rtol=1e-8, atol=1e-11)
    assert np.all(dxi.ravel()[removed] == 0)
    np.testing.assert_allclose(dX, g["lin_dX"], rtol=1e-8, atol=1e-11)
    tX, tf, tu, tR, tt = O.apply_update(X, f, u, R, t, dxi, dX)
    np.testing.assert_allclose(tR, g["lin_tR"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(tt, g["lin_tt"], rtol=0, atol=1e-11)
    assert O.cost(obs, tX, tf, tu, tR, tt, f0) == pytest.approx(float(g["lin_E_trial"]), rel=1e-9)


@pytest.mark.parametrize("name", RUN_CASES)
def test_full_run_matches_reference(name):
    g = load_golden(name)
    x, vis, X0, K0, R0, t0, axis, f0 = case_inputs(g)
    ba = O.OracleBundleAdjuster(x, X0, K0, R0, t0, f0=f0, visibility_index=vis, axis=axis)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        X, K, R, t = ba.optimize(2.0, 1e-8, max_iter=100, is_debug=True)
    E = np.array([d["reprojection_error"] for d in ba.get_log()])
    # identical accept/reject sequence => identical number of accepted iterations
    assert E.shape == g["E"].shape
    np.testing.assert_allclose(E, g["E"], rtol=1e-9, atol=0)
    np.testing.assert_allclose(X, g["X"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(K, g["K"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(R, g["R"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(t, g["t"], rtol=0, atol=1e-6)
    # same number of printed lines in the reference's format (:188)
    ref_lines = str(g["stdout"]).strip().splitlines()
    got_lines = buf.getvalue().strip().splitlines()
    assert len(ref_lines) == len(got_lines)
    assert got_lines[0].startswith("Iteration 1: reprojection_error_delta = ")
    log_pts = np.stack([d["points"] for d in ba.get_log()])
    np.testing.assert_allclose(log_pts, g["log_points"], rtol=0, atol=1e-6)


def test_c1_literals_from_survey():
    """Literals recorded by the survey as an independent check of the
    fixture itself."""
    g = load_golden("c1_euclid")
    assert len(g["E"]) == 38
    assert g["E"][0] == pytest.approx(66.31926634440296, rel=1e-12)
    assert g["E"][1] == pytest.approx(10.101413537088082, rel=1e-12)
    assert g["E"][-1] == pytest.approx(0.08011501624122017, rel=1e-10)
    assert g["K"][0, 0, 0] == pytest.approx(0.8848476409241687, rel=1e-9)


def test_bad_axis_raises_value_error():
    g = load_golden("small_dense_xup")
    x, vis, X0, K0, R0, t0, axis, f0 = case_inputs(g)
    with pytest.raises(ValueError):
        O.OracleBundleAdjuster(x, X0, K0, R0, t0, axis="z-up")


def test_rodrigues_exact_identity_for_zero():
    assert np.array_equal(O.rodrigues(np.zeros(3)), np.eye(3))
    w = np.array([0.3, -0.2, 0.5])
    Rm = O.rodrigues(w)
    np.testing.assert_allclose(Rm @ Rm.T, np.eye(3), atol=1e-15)
    np.testing.assert_allclose(np.linalg.det(Rm), 1.0, atol=1e-15)
