"""The N > 1 path on CPU: the package's sharded LM loop (``sharded.lm_loop``: two all-reduces per
inner solve) driven over gloo with world_size 2 and 3.  The per-rank compute is a NumPy stand-in
built from the oracle (test infrastructure) that implements the same phase interface as the CUDA
``Engine``; the loop, the sharding helper and the collectives are the product code.  The sharded
cost trajectory must reproduce the unmodified reference's single-process run."""
import os
import socket
import sys
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class OracleShardEngine:
    """NumPy implementation of the engine's phase interface for one shard of points."""

    def __init__(self, obs, X, f, u, R, t, f0, axis):
        from oracle import ba_oracle as O

        self.O, self.obs, self.f0, self.axis = O, obs, f0, axis
        self.X, self.f, self.u, self.R, self.t = X.copy(), f.copy(), u.copy(), R.copy(), t.copy()
        M = obs.n_cams
        self.M, self.nfull = M, 9 * M
        self._red = torch.zeros(self.nfull * self.nfull + self.nfull + 90 * M, dtype=torch.float64)
        self._cost = torch.zeros(4, dtype=torch.float64)  # as ba_cost_buffer: E0, E_trial, singular flag, pad
        self.st = types.SimpleNamespace()
        self.records = []

    def reduce_tensor(self):
        return self._red

    def cost_tensor(self):
        return self._cost

    def lm_begin(self, scale, tol, max_iter, max_retries=200):
        s = self.st
        s.E = s.E_trial = s.delta = 0.0
        s.c, s.scale, s.tol, s.max_iter, s.max_retries = 1e-4, scale, tol, max_iter, max_retries
        s.count = s.solves = s.iter_solves = 0
        s.need_linearize, s.accepted, s.done, s.status = 1, 0, 0, 0
        self._cost[0] = self.O.cost(self.obs, self.X, self.f, self.u, self.R, self.t, self.f0)

    def lm_phase_reduce(self):
        O, s, M, n = self.O, self.st, self.M, self.nfull
        if s.done:
            return
        if s.solves == 0:
            s.E = float(self._cost[0])
        if s.need_linearize:
            self.lin = O.linearize(self.obs, self.X, self.f, self.u, self.R, self.t, self.f0)
        # local sum_j F^T E^-1 F and sum_j F^T E^-1 d_P (damped point blocks, undamped U)
        try:
            A0, b0, self.Vinv = O.reduced_system(self.obs, self.lin, s.c)
        except np.linalg.LinAlgError:
            # like the device: flag it, keep going with whatever the buffers hold; the flag is
            # summed over the ranks with the trial cost and stops everybody in lm_phase_decide
            s.status = 3  # BA_ERR_SINGULAR
            self._red.zero_()
            return
        Ud = np.zeros((n, n))
        for i in range(M):
            blk = self.lin.U[i].copy()
            blk[np.arange(9), np.arange(9)] *= 1 + s.c
            Ud[9 * i: 9 * i + 9, 9 * i: 9 * i + 9] = blk
        P = Ud - A0
        rhsP = b0 + self.lin.g_cam.reshape(-1)
        buf = np.concatenate((P.ravel(), rhsP, self.lin.U.ravel(), self.lin.g_cam.ravel()))
        self._red.copy_(torch.from_numpy(buf))

    def lm_phase_solve(self):
        O, s, M, n = self.O, self.st, self.M, self.nfull
        if s.done:
            return
        self._cost[2] = 1.0 if s.status == 3 else 0.0
        if s.status == 3:
            self._cost[1] = 0.0
            return
        buf = self._red.numpy()
        P = buf[: n * n].reshape(n, n)
        rhsP = buf[n * n: n * n + n]
        U = buf[n * n + n: n * n + n + 81 * M].reshape(M, 9, 9)
        gcam = buf[n * n + n + 81 * M:]
        A = -P.copy()
        for i in range(M):
            blk = U[i].copy()
            blk[np.arange(9), np.arange(9)] *= 1 + s.c
            A[9 * i: 9 * i + 9, 9 * i: 9 * i + 9] += blk
        b = rhsP - gcam
        _, kept = O.gauge_indices(M, self.axis)
        dxi = np.zeros(n)
        try:
            dxi[kept] = np.linalg.solve(A[np.ix_(kept, kept)], b[kept])
        except np.linalg.LinAlgError:  # a peer's singular block left a garbage system
            dxi[:] = 0.0
        dxi = dxi.reshape(M, 9)
        obs, lin = self.obs, self.lin
        Fd = O._segment_sum(obs.pt, np.einsum("oab,ob->oa", lin.W, dxi[obs.cam]), obs.n_points)
        dX = -np.einsum("jab,jb->ja", self.Vinv, Fd + lin.g_pt)
        self.trial = O.apply_update(self.X, self.f, self.u, self.R, self.t, dxi, dX)
        tX, tf, tu, tR, tt = self.trial
        self._cost[1] = O.cost(obs, tX, tf, tu, tR, tt, self.f0)

    def lm_phase_decide(self):
        s = self.st
        if s.done:
            return
        E_ = float(self._cost[1])
        s.E_trial = E_
        s.solves += 1
        s.iter_solves += 1
        if float(self._cost[2]) > 0.0:
            s.status = 3
        if s.status != 0:
            s.done, s.accepted = 1, 0
            return
        if E_ > s.E:
            s.c *= s.scale
            s.accepted, s.need_linearize = 0, 0
            return
        s.accepted, s.need_linearize = 1, 1
        s.count += 1
        s.delta = abs(E_ - s.E)
        self.records.append(E_)
        self.X, self.f, self.u, self.R, self.t = self.trial
        s.iter_solves = 0
        if s.delta <= s.tol or s.count >= s.max_iter:
            s.done = 1
            s.E = E_
        else:
            s.E = E_
            s.c /= s.scale

    def lm_state(self):
        return types.SimpleNamespace(**vars(self.st))


def _worker(rank, world, port, case, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import ba_b200
        from conftest import case_inputs, load_golden
        from oracle import ba_oracle as O

        sharded = ba_b200.submodule("sharded")
        g = load_golden(case)
        x, vis, X0, K0, R0, t0, axis, f0 = case_inputs(g)
        obs = O.ObsList.from_dense(x, vis)
        X, R, t = O.normalize_gauge(X0, R0, t0, axis)
        f, u = K0[:, 0, 0].copy(), K0[:, :2, 2].copy()
        lo, hi = sharded.shard_bounds(obs.n_points, world, obs.ptr)[rank]
        eng = OracleShardEngine(obs.subset_points(lo, hi), X[lo:hi], f, u, R, t, f0, axis)
        st = sharded.lm_loop(eng, dist, None, 2.0, 1e-8, 100)
        E = np.array([float(g["E"][0])] + eng.records)
        np.save(os.path.join(out_dir, f"E_{rank}.npy"), E)
        np.save(os.path.join(out_dir, f"f_{rank}.npy"), eng.f)
        np.save(os.path.join(out_dir, f"X_{rank}.npy"), np.concatenate(([lo, hi], eng.X.ravel())))
        assert st.done
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,case", [(2, "small_sparse_xup"), (3, "small_dense_xright")])
def test_sharded_loop_reproduces_reference(tmp_path, world, case):
    from conftest import load_golden

    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    g = load_golden(case)
    Es = [np.load(tmp_path / f"E_{r}.npy") for r in range(world)]
    for E in Es:  # every rank saw the same (all-reduced) costs and took the same decisions
        assert E.shape == g["E"].shape
        np.testing.assert_allclose(E, g["E"], rtol=1e-9)
        np.testing.assert_array_equal(E, Es[0])
    fs = [np.load(tmp_path / f"f_{r}.npy") for r in range(world)]
    for f in fs:  # replicated camera state stays identical
        np.testing.assert_array_equal(f, fs[0])
    np.testing.assert_allclose(fs[0], g["K"][:, 0, 0], atol=1e-6)
    # the point shards tile the scene
    spans = sorted(tuple(np.load(tmp_path / f"X_{r}.npy")[:2].astype(int)) for r in range(world))
    assert spans[0][0] == 0 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _worker_singular(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import ba_b200
        from conftest import case_inputs, load_golden
        from oracle import ba_oracle as O

        sharded = ba_b200.submodule("sharded")
        g = load_golden("small_sparse_xup")
        x, vis, X0, K0, R0, t0, axis, f0 = case_inputs(g)
        vis = vis.copy()
        vis[1, :] = False  # a point without any view, owned by rank 0: inv() raises there (:128)
        obs = O.ObsList.from_dense(x, vis)
        X, R, t = O.normalize_gauge(X0, R0, t0, axis)
        f, u = K0[:, 0, 0].copy(), K0[:, :2, 2].copy()
        lo, hi = sharded.shard_bounds(obs.n_points, world, obs.ptr)[rank]
        assert (lo <= 1 < hi) == (rank == 0)
        eng = OracleShardEngine(obs.subset_points(lo, hi), X[lo:hi], f, u, R, t, f0, axis)
        st = sharded.lm_loop(eng, dist, None, 2.0, 1e-8, 100)
        np.save(os.path.join(out_dir, f"st_{rank}.npy"), np.array([st.status, st.done, st.solves, st.count]))
    finally:
        dist.destroy_process_group()


def test_singular_block_on_one_rank_stops_every_rank(tmp_path):
    """with the collective exchange only the trial cost was summed, so the rank owning
    a 0-view point stopped (LinAlgError) while its peers waited in the next all-reduce forever.
    The singular flag now travels with the cost: every rank stops in the first solve."""
    world = 2
    mp.spawn(_worker_singular, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        status, done, solves, count = np.load(tmp_path / f"st_{r}.npy")
        assert (status, done, solves, count) == (3, 1, 1, 0)
