"""Sharded CUDA path: two ranks (processes), each owns half of the points.

* exchange="collective": the two sums per inner solve are `torch.distributed` all-reduces (gloo
  here, both ranks share cuda:0; NCCL needs one GPU per rank and is exercised by
  `bench.py --gpus N --exchange collective`).  Checks the phase API (`ba_lm_phase_*`, reduce /
  cost buffers).
* exchange="peer": the sums are formed by the library's own kernels over peer memory (CUDA IPC
  windows, `csrc/comm_peer.cu`) and the loop is the same CUDA-graph loop as on one GPU.  Needs one
  GPU per rank (NVLink); skipped on a single-GPU box (see `_need_gpu_per_rank`).

Both are checked against the reference's golden trajectory."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _need_gpu_per_rank(world: int):
    """The peer exchange spins on flags that the other rank's kernels write.  Two such ranks
    time-sliced on ONE GPU can never be guaranteed to run at the same time (B200_PROFILING.md: Xid 109
    context-switch timeouts with 2 and 4 ranks on one GPU), so these tests need one GPU per rank;
    on a single-GPU box the protocol is covered by the collective variant, by the CPU gloo tests and
    -- over NVLink, under an assertion -- by `bench.py --gpus N`'s parity_vs_n1 gate."""
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"peer-memory exchange needs {world} GPUs (one per rank); this box has "
                    f"{torch.cuda.device_count()}")


def _worker(rank, world, port, case, out_dir, exchange="collective", per_rank_gpu=False):
    import contextlib
    import io

    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = rank if per_rank_gpu else 0
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import ba_b200
        from conftest import case_inputs, load_golden

        sharded = ba_b200.submodule("sharded")
        adjuster = ba_b200.submodule("bundle_adjuster")
        g = load_golden(case)
        x, vis, X0, K0, R0, t0, axis, f0 = case_inputs(g)
        full = adjuster.ObservationList.from_dense(x, vis)
        lo, hi = sharded.shard_bounds(full.n_points, world, full.obs_ptr)[rank]
        a, b = int(full.obs_ptr[lo]), int(full.obs_ptr[hi])
        xy = full.obs_xy if full.obs_xy is not None else np.ascontiguousarray(full.dense_x).reshape(-1, 2)
        adj = ba_b200.BundleAdjuster.from_observations(
            full.obs_ptr[lo:hi + 1] - full.obs_ptr[lo],
            None if full.dense else full.obs_cam[a:b], xy[a:b],
            X0[lo:hi], K0, R0, t0, f0=f0, axis=axis, dense=full.dense, device=dev,
            process_group=dist.group.WORLD, exchange=exchange)
        assert adj._peer_exchange == (exchange == "peer")
        assert adj.engine.comm_size() == (world if exchange == "peer" else 1)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            X, K, R, t = adj.optimize(2.0, 1e-8, max_iter=100, is_debug=(rank == 0))
        E = np.array([adj.records[0]["E_prev"]] + [r["E"] for r in adj.records])
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), E=E, X=X, K=K, R=R, t=t, lo=lo, hi=hi,
                 lines=len(buf.getvalue().strip().splitlines()),
                 nlog=len(adj.get_log()))
        adj.engine.close()
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("exchange", ["collective", "peer"])
@pytest.mark.parametrize("case", ["small_sparse_xup", "mid_dense_xup"])
def test_two_ranks_match_reference(tmp_path, case, exchange):
    import torch
    import torch.multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_golden

    world = 2
    if exchange == "peer":
        _need_gpu_per_rank(world)
    per_rank_gpu = exchange == "peer"
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path), exchange, per_rank_gpu),
             nprocs=world, join=True)
    g = load_golden(case)
    out = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    X = np.zeros_like(g["X"])
    for o in out:
        assert o["E"].shape == g["E"].shape
        np.testing.assert_allclose(o["E"], g["E"], rtol=1e-9)
        np.testing.assert_allclose(o["K"], g["K"], atol=1e-6)
        np.testing.assert_allclose(o["R"], g["R"], atol=1e-6)
        np.testing.assert_allclose(o["t"], g["t"], atol=1e-6)
        X[int(o["lo"]):int(o["hi"])] = o["X"]
    np.testing.assert_allclose(X, g["X"], atol=1e-6)
    # only rank 0 prints the reference's iteration lines; its debug log has one entry per state
    assert int(out[0]["lines"]) == len(g["E"]) - 1 and int(out[1]["lines"]) == 0
    assert int(out[0]["nlog"]) == len(g["E"])


def _worker_large(rank, world, port, out_dir, per_rank_gpu):
    import contextlib
    import io

    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = rank if per_rank_gpu else 0
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        import ba_b200

        sharded = ba_b200.submodule("sharded")
        sc = ba_b200.scenes.make_scene(320, 450, seed=7, visibility=0.3)
        lo, hi = sharded.shard_bounds(sc.n_points, world, sc.obs_ptr)[rank]
        a, b = int(sc.obs_ptr[lo]), int(sc.obs_ptr[hi])
        adj = ba_b200.BundleAdjuster.from_observations(
            sc.obs_ptr[lo:hi + 1] - sc.obs_ptr[lo], sc.obs_cam[a:b], sc.obs_xy[a:b], sc.X0[lo:hi], sc.K0,
            sc.R0, sc.t0, f0=sc.f0, axis=sc.axis, device=dev, process_group=dist.group.WORLD, exchange="peer")
        with contextlib.redirect_stdout(io.StringIO()):
            X, K, R, t = adj.optimize(2.0, 1e-8, max_iter=3)
        E = np.array([adj.records[0]["E_prev"]] + [r["E"] for r in adj.records])
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), E=E, K=K, R=R, t=t)
        adj.engine.close()
    finally:
        dist.destroy_process_group()


def test_divided_cholesky_of_a_large_reduced_system(tmp_path):
    """320 cameras (n = 2880): the reduced system is factored by the two-level Cholesky whose first
    rank-256 trailing update is divided over the two ranks (each updates its own 128-tile rows and
    stores the next block column into the other rank's matrix over peer memory).  Three iterations
    against the CPU oracle on the whole scene; both ranks must hold identical cameras."""
    import torch
    import torch.multiprocessing as mp

    import ba_b200
    from oracle import ba_oracle as O

    world = 2
    _need_gpu_per_rank(world)
    per_rank_gpu = True
    mp.spawn(_worker_large, args=(world, _free_port(), str(tmp_path), per_rank_gpu), nprocs=world, join=True)
    out = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    sc = ba_b200.scenes.make_scene(320, 450, seed=7, visibility=0.3)
    obs = O.ObsList(sc.n_points, sc.n_cams, np.repeat(np.arange(sc.n_points), np.diff(sc.obs_ptr)),
                    sc.obs_cam.astype(np.int64), sc.obs_xy, sc.obs_ptr)
    ora = O.OracleBundleAdjuster(None, sc.X0, sc.K0, sc.R0, sc.t0, f0=sc.f0, axis=sc.axis, obs=obs)
    Xo, Ko, Ro, to = ora.optimize(2.0, 1e-8, max_iter=3, verbose=False)
    Eo = np.array([r["E"] for r in ora.trace])
    for o in out:
        assert o["E"].shape == Eo.shape
        np.testing.assert_allclose(o["E"], Eo, rtol=1e-9)
        np.testing.assert_allclose(o["K"], Ko, atol=1e-6)
        np.testing.assert_allclose(o["R"], Ro, atol=1e-6)
        np.testing.assert_allclose(o["t"], to, atol=1e-6)
    assert np.array_equal(out[0]["R"], out[1]["R"]) and np.array_equal(out[0]["t"], out[1]["t"])


def _worker_singular(rank, world, port, out_dir, exchange, per_rank_gpu):
    import contextlib
    import io

    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = rank if per_rank_gpu else 0
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import ba_b200
        from conftest import case_inputs, load_golden

        sharded = ba_b200.submodule("sharded")
        adjuster = ba_b200.submodule("bundle_adjuster")
        g = load_golden("small_sparse_xup")
        x, vis, X0, K0, R0, t0, axis, f0 = case_inputs(g)
        vis = vis.copy()
        vis[1, :] = False  # a 0-view point in rank 0's shard (reference: LinAlgError at :128)
        full = adjuster.ObservationList.from_dense(x, vis)
        lo, hi = sharded.shard_bounds(full.n_points, world, full.obs_ptr)[rank]
        a, b = int(full.obs_ptr[lo]), int(full.obs_ptr[hi])
        adj = ba_b200.BundleAdjuster.from_observations(
            full.obs_ptr[lo:hi + 1] - full.obs_ptr[lo], full.obs_cam[a:b], full.obs_xy[a:b],
            X0[lo:hi], K0, R0, t0, f0=f0, axis=axis, device=dev,
            process_group=dist.group.WORLD, exchange=exchange)
        raised = False
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                adj.optimize(2.0, 1e-8, max_iter=20)
        except np.linalg.LinAlgError:
            raised = True
        st = adj.engine.lm_state()
        np.save(os.path.join(out_dir, f"s{rank}.npy"), np.array([int(raised), st.solves, st.count, int(lo <= 1 < hi)]))
        adj.engine.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("exchange", ["collective", "peer"])
def test_singular_block_on_one_rank_raises_on_every_rank(tmp_path, exchange):
    """the rank that owns a point without views must not stop alone -- with the
    collective exchange its peers used to wait in the next all-reduce forever.  Both exchanges carry
    the flag with the trial cost: every rank raises LinAlgError after the same (first) solve."""
    import torch
    import torch.multiprocessing as mp

    world = 2
    if exchange == "peer":
        _need_gpu_per_rank(world)
    per_rank_gpu = exchange == "peer"
    mp.spawn(_worker_singular, args=(world, _free_port(), str(tmp_path), exchange, per_rank_gpu),
             nprocs=world, join=True)
    out = [np.load(tmp_path / f"s{r}.npy") for r in range(world)]
    assert sorted(int(o[3]) for o in out) == [0, 1]  # exactly one rank owns the bad point
    for o in out:
        assert tuple(o[:3]) == (1, 1, 0)
