"""Synthetic bundle-adjustment scenes (benchmark / test input, not part of the hot path).

The geometry follows the reference's own recipe so the scenes are the ones its code was
written for:
  * cameras on a radius-5 hemisphere, ``theta ~ U[0, pi/2]``, ``phi ~ U[0, 2 pi]``
    (reference ``lib/utils.py:40-52``), looking at targets ``N(0, 0.5^2)^3``
    (``euclidiean_reconstruction.py:21``), ``f = f0 = 1`` (``:24``);
  * camera frame: columns of ``R`` are the camera's up / right / forward axes in world
    coordinates, ``t`` is the camera centre, ``P = K [R^T | -R^T t]``
    (``lib/camera.py:13-14,43-55``);
  * points ``U[-1, 1]^3``; observations = exact projection + ``noise * N(0, 1)``
    (``euclidiean_reconstruction.py:40``);
  * initial value = ground truth perturbed (X, t: ``+ perturb * N(0,1)``; R: ``exp(perturb *
    N(0,1)) R``; f scaled by ``f_scale``) because the factorization initialiser is
    O(M N^2) memory (``lib/perspective_camera_calibration.py:188``).

Scenes are produced directly as observation lists (CSR by point) so the 1000 x 1M sparse
configs never need the 16 GB dense ``(N, M, 2)`` array.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Scene:
    n_points: int
    n_cams: int
    f0: float
    axis: str
    # observation list, sorted by point then camera
    obs_ptr: np.ndarray  # (N+1,) int64
    obs_cam: np.ndarray  # (nobs,) int32
    obs_xy: np.ndarray  # (nobs, 2) float64
    dense: bool  # True when every point sees every camera
    # initial values in the caller's (un-normalised) frame
    X0: np.ndarray
    K0: np.ndarray
    R0: np.ndarray
    t0: np.ndarray
    # ground truth
    X_gt: np.ndarray
    K_gt: np.ndarray
    R_gt: np.ndarray
    t_gt: np.ndarray

    @property
    def nobs(self) -> int:
        return int(self.obs_cam.shape[0])

    def dense_x(self):
        """Dense ``x (N, M, 2)`` and bool mask ``(N, M)`` in the reference layout.

        Invisible entries are filled with 0.0: the reference evaluates them too and needs
        finite values there (``lib/bundle_adjustment.py:674``).
        """
        x = np.zeros((self.n_points, self.n_cams, 2))
        vis = np.zeros((self.n_points, self.n_cams), dtype=bool)
        pt = np.repeat(np.arange(self.n_points), np.diff(self.obs_ptr))
        x[pt, self.obs_cam] = self.obs_xy
        vis[pt, self.obs_cam] = True
        return x, vis


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def hemisphere_positions(rs: np.random.RandomState, n: int, radius: float) -> np.ndarray:
    """Same distribution (and, for the legacy RandomState stream, the same draw order:
    theta_0, phi_0, theta_1, ...) as reference ``lib/utils.py:40-52``."""
    uu = rs.random_sample((n, 2))
    theta = uu[:, 0] * (np.pi / 2)
    phi = uu[:, 1] * (2 * np.pi)
    return np.stack(
        (radius * np.cos(theta), radius * np.sin(theta) * np.cos(phi), radius * np.sin(theta) * np.sin(phi)),
        axis=1,
    )


def look_at(pos: np.ndarray, target: np.ndarray) -> np.ndarray:
    """Camera-to-world rotations: columns = camera up, right, forward (``lib/camera.py:43-55``)."""
    top = np.array([1.0, 0.0, 0.0])
    cz = _unit(target - pos)
    cy = _unit(np.cross(cz, top))
    cx = _unit(np.cross(cy, cz))
    return np.stack((cx, cy, cz), axis=2)


def rodrigues_batch(w: np.ndarray) -> np.ndarray:
    """Axis-angle vectors (n, 3) -> rotation matrices (n, 3, 3)."""
    th = np.linalg.norm(w, axis=1)
    safe = np.where(th > 0, th, 1.0)
    l = w / safe[:, None]
    c, s = np.cos(th)[:, None, None], np.sin(th)[:, None, None]
    skew = np.zeros((w.shape[0], 3, 3))
    skew[:, 0, 1], skew[:, 0, 2] = -l[:, 2], l[:, 1]
    skew[:, 1, 0], skew[:, 1, 2] = l[:, 2], -l[:, 0]
    skew[:, 2, 0], skew[:, 2, 1] = -l[:, 1], l[:, 0]
    return (1 - c) * (l[:, :, None] * l[:, None, :]) + c * np.eye(3) + s * skew


def project_obs(X, f, u, R, t, f0, pt, cam):
    """Image coordinates (in f0 units times f0) of the listed (point, camera) pairs."""
    d = X[pt] - t[cam]
    xc = np.einsum("oki,ok->oi", R[cam], d)  # R^T (X - t)
    p = f[cam] * xc[:, 0] + u[cam, 0] * xc[:, 2]
    q = f[cam] * xc[:, 1] + u[cam, 1] * xc[:, 2]
    r = f0 * xc[:, 2]
    return np.stack((p / r, q / r), axis=1) * f0


def make_scene(
    n_cams: int,
    n_points: int,
    seed: int = 0,
    visibility: float = 1.0,
    noise: float = 0.005,
    outlier_frac: float = 0.0,
    perturb: float = 0.02,
    f_scale: float = 1.05,
    min_views: int = 3,
    axis: str = "x-up_z-forward",
    chunk: int = 20000,
    point_stream: int = 0,
    chunk_seeded: bool = False,
    point_range: tuple[int, int] | None = None,
) -> Scene:
    """`seed` fixes the cameras (ground truth and initial perturbation); the points, their
    visibility and the image noise come from a second stream keyed by (`seed`,
    `point_stream`), so that the ranks of a sharded run share the cameras and draw disjoint
    point shards (`point_stream = rank`).

    `chunk_seeded=True` keys that second stream per chunk of `chunk` points instead, so that any
    contiguous `point_range=(lo, hi)` of the scene can be generated on its own and is identical
    to rows lo:hi of the whole scene: the ranks of a strong-scaled run each build their shard of
    the *same* global scene without anybody building all of it."""
    if chunk_seeded:
        return _make_scene_chunk_seeded(n_cams, n_points, seed, visibility, noise, outlier_frac, perturb,
                                        f_scale, min_views, axis, chunk, point_stream, point_range)
    if point_range is not None:
        raise ValueError("point_range needs chunk_seeded=True")
    rs_cam = np.random.RandomState(seed)
    f0 = 1.0
    pos = hemisphere_positions(rs_cam, n_cams, 5.0)
    targets = rs_cam.normal(0, 0.5, (n_cams, 3))
    R_gt = look_at(pos, targets)
    t_gt = pos
    f_gt = np.ones(n_cams)
    u_gt = np.zeros((n_cams, 2))
    t0 = t_gt + perturb * rs_cam.standard_normal(t_gt.shape)
    R0 = rodrigues_batch(perturb * rs_cam.standard_normal((n_cams, 3))) @ R_gt

    rs = np.random.RandomState((seed * 7919 + 104729 * (point_stream + 1)) % (2**31 - 1))
    X_gt = rs.uniform(-1, 1, (n_points, 3))

    dense = visibility >= 1.0
    counts = np.empty(n_points, dtype=np.int64)
    cam_chunks, xy_chunks = [], []
    for lo in range(0, n_points, chunk):
        hi = min(n_points, lo + chunk)
        if dense:
            vis = np.ones((hi - lo, n_cams), dtype=bool)
        else:
            vis = rs.random_sample((hi - lo, n_cams)) < visibility
            # cameras 0 and 1 carry the gauge: keep them well covered, and give every
            # point at least `min_views` views (the reference raises on a 0-view point and
            # silently mis-solves a 1-view point)
            short = np.nonzero(vis.sum(axis=1) < min_views)[0]
            for j in short:
                need = min_views - int(vis[j].sum())
                free = np.nonzero(~vis[j])[0]
                vis[j, rs.choice(free, size=need, replace=False)] = True
        pt, cam = np.nonzero(vis)
        counts[lo:hi] = vis.sum(axis=1)
        xy = project_obs(X_gt, f_gt, u_gt, R_gt, t_gt, f0, pt + lo, cam)
        xy += noise * rs.standard_normal(xy.shape)
        if outlier_frac > 0:
            bad = rs.random_sample(xy.shape[0]) < outlier_frac
            xy[bad] = rs.uniform(-0.5, 0.5, (int(bad.sum()), 2))
        cam_chunks.append(cam.astype(np.int32))
        xy_chunks.append(xy)
    obs_cam = np.concatenate(cam_chunks)
    obs_xy = np.ascontiguousarray(np.concatenate(xy_chunks))
    obs_ptr = np.concatenate(([0], np.cumsum(counts))).astype(np.int64)

    # perturbed-ground-truth initial value
    X0 = X_gt + perturb * rs.standard_normal(X_gt.shape)
    K_gt = np.zeros((n_cams, 3, 3))
    K_gt[:, 0, 0] = K_gt[:, 1, 1] = f_gt
    K_gt[:, 2, 2] = f0
    K0 = K_gt.copy()
    K0[:, 0, 0] *= f_scale
    K0[:, 1, 1] *= f_scale
    return Scene(n_points, n_cams, f0, axis, obs_ptr, obs_cam, obs_xy, dense,
                 X0, K0, R0, t0, X_gt, K_gt, R_gt, t_gt)


def _cameras(seed: int, n_cams: int, perturb: float, f_scale: float, f0: float):
    """Ground-truth cameras and their perturbed initial values (the `seed` stream of make_scene)."""
    rs_cam = np.random.RandomState(seed)
    pos = hemisphere_positions(rs_cam, n_cams, 5.0)
    targets = rs_cam.normal(0, 0.5, (n_cams, 3))
    R_gt = look_at(pos, targets)
    t_gt = pos
    t0 = t_gt + perturb * rs_cam.standard_normal(t_gt.shape)
    R0 = rodrigues_batch(perturb * rs_cam.standard_normal((n_cams, 3))) @ R_gt
    K_gt = np.zeros((n_cams, 3, 3))
    K_gt[:, 0, 0] = K_gt[:, 1, 1] = 1.0
    K_gt[:, 2, 2] = f0
    K0 = K_gt.copy()
    K0[:, 0, 0] *= f_scale
    K0[:, 1, 1] *= f_scale
    return R_gt, t_gt, K_gt, R0, t0, K0


def _make_scene_chunk_seeded(n_cams, n_points, seed, visibility, noise, outlier_frac, perturb, f_scale,
                             min_views, axis, chunk, point_stream, point_range) -> Scene:
    f0 = 1.0
    R_gt, t_gt, K_gt, R0, t0, K0 = _cameras(seed, n_cams, perturb, f_scale, f0)
    f_gt, u_gt = np.ones(n_cams), np.zeros((n_cams, 2))
    lo_pt, hi_pt = (0, n_points) if point_range is None else (int(point_range[0]), int(point_range[1]))
    if not (0 <= lo_pt <= hi_pt <= n_points):
        raise ValueError("point_range out of bounds")
    dense = visibility >= 1.0
    Xg, X0s, cnts, cams, xys = [], [], [], [], []
    for ci in range(lo_pt // chunk, (max(hi_pt, lo_pt + 1) - 1) // chunk + 1):
        c0, c1 = ci * chunk, min(n_points, (ci + 1) * chunk)
        rs = np.random.RandomState([seed % (2**31 - 1), point_stream, ci, 0x5CE9E])
        X_c = rs.uniform(-1, 1, (c1 - c0, 3))
        if dense:
            vis = np.ones((c1 - c0, n_cams), dtype=bool)
        else:
            vis = rs.random_sample((c1 - c0, n_cams)) < visibility
            for j in np.nonzero(vis.sum(axis=1) < min_views)[0]:  # see make_scene
                free = np.nonzero(~vis[j])[0]
                vis[j, rs.choice(free, size=min_views - int(vis[j].sum()), replace=False)] = True
        pt, cam = np.nonzero(vis)
        xy = project_obs(X_c, f_gt, u_gt, R_gt, t_gt, f0, pt, cam)
        xy += noise * rs.standard_normal(xy.shape)
        if outlier_frac > 0:
            bad = rs.random_sample(xy.shape[0]) < outlier_frac
            xy[bad] = rs.uniform(-0.5, 0.5, (int(bad.sum()), 2))
        X0_c = X_c + perturb * rs.standard_normal(X_c.shape)
        # rows of this chunk inside the requested range
        a, b = max(lo_pt, c0) - c0, min(hi_pt, c1) - c0
        cnt = vis.sum(axis=1)
        ptr_c = np.concatenate(([0], np.cumsum(cnt)))
        Xg.append(X_c[a:b]); X0s.append(X0_c[a:b]); cnts.append(cnt[a:b])
        cams.append(cam[ptr_c[a]:ptr_c[b]].astype(np.int32)); xys.append(xy[ptr_c[a]:ptr_c[b]])
    counts = np.concatenate(cnts) if cnts else np.zeros(0, dtype=np.int64)
    obs_ptr = np.concatenate(([0], np.cumsum(counts))).astype(np.int64)
    return Scene(hi_pt - lo_pt, n_cams, f0, axis, obs_ptr, np.concatenate(cams),
                 np.ascontiguousarray(np.concatenate(xys)), dense, np.concatenate(X0s), K0, R0, t0,
                 np.concatenate(Xg), K_gt, R_gt, t_gt)


# The named benchmark configurations of BASELINE.json (`configs[1..4]`).  c3-c5 are chunk-seeded so
# that the ranks of a strong-scaled run can each generate their shard of the same global scene.
CONFIGS = {
    "c2": dict(n_cams=50, n_points=10_000, visibility=1.0, seed=2),
    "c3": dict(n_cams=200, n_points=100_000, visibility=1.0, seed=3, chunk_seeded=True),
    "c4": dict(n_cams=1000, n_points=1_000_000, visibility=0.1, seed=4, chunk_seeded=True),
    "c5": dict(n_cams=1000, n_points=1_000_000, visibility=0.1, outlier_frac=0.01, seed=5, chunk_seeded=True),
}
