"""Point-sharded Levenberg-Marquardt loop: one engine (GPU) per rank, cameras replicated.

Every rank owns a contiguous shard of the points with all their observations and evaluates
residuals, Jacobians, point blocks and its share of the Schur products locally.  Per inner
solve there are exactly two exchanges.

This module is the ``exchange="collective"`` form: both are ``all_reduce(SUM)`` over
``torch.distributed`` (NCCL over NVLink on a GPU box, gloo in the CPU tests).  The default on
NCCL groups is ``exchange="peer"``: the sums are kernels of the library itself over NVLink peer
memory (``csrc/comm_peer.cu``), the loop is the single-engine CUDA-graph loop (``ba_lm_run``), and
nothing of this module but ``shard_bounds`` is used.  The two exchanges:

  1. the partial reduced system  [sum_j Y_j Y_j^T | rhs row | U_i | dF_i]  (one flat buffer),
  2. the trial cost and the singular-block flag (two doubles).

After (1) every rank holds the same reduced camera system and factors it redundantly
(replicated Cholesky), so the camera step, the accept/reject decision and the damping schedule
are bit-identical on all ranks; each rank back-substitutes only its own points.

The loop is written against a small *phase interface* (``lm_begin``, ``lm_phase_reduce``,
``lm_phase_solve``, ``lm_phase_decide``, ``lm_state``, ``reduce_tensor``, ``cost_tensor``) that
the CUDA ``Engine`` implements; the CPU tests drive the very same loop with a NumPy stand-in
under gloo (``tests/test_sharded_gloo.py``).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_points: int, world: int, obs_ptr: np.ndarray | None = None):
    """Contiguous point ranges, balanced by observation count when ``obs_ptr`` is given."""
    if obs_ptr is None:
        edges = [(n_points * r) // world for r in range(world + 1)]
    else:
        total = int(obs_ptr[-1])
        targets = [(total * r) // world for r in range(world + 1)]
        edges = [int(np.searchsorted(obs_ptr, t, side="left")) for t in targets]
        edges[0], edges[-1] = 0, n_points
        for r in range(1, world + 1):  # keep the ranges non-decreasing
            edges[r] = max(edges[r], edges[r - 1])
    return [(edges[r], edges[r + 1]) for r in range(world)]


def lm_loop(engine, dist, group, scale_factor, delta_tol, max_iter, max_retries=200,
            on_state=None):
    """Run the sharded LM loop; ``on_state(st)`` is called after every decide on every rank.

    Returns the final control block.  ``dist`` is ``torch.distributed`` (or any object with the
    same ``all_reduce``); ``group`` the process group (None = world).
    """
    red = engine.reduce_tensor()
    cost = engine.cost_tensor()
    engine.lm_begin(scale_factor, delta_tol, max_iter, max_retries)
    dist.all_reduce(cost[0:1], group=group)  # cost of the initial state (:85-87)

    def enqueue_solve():
        engine.lm_phase_reduce()
        dist.all_reduce(red, group=group)
        engine.lm_phase_solve()
        # trial cost and the "singular point block" flag together: the rank that owns a singular
        # block (reference :128 raises LinAlgError) must stop every rank in this very solve, or
        # the others would wait for it in the next collective forever
        dist.all_reduce(cost[1:3], group=group)
        engine.lm_phase_decide()

    if on_state is None and hasattr(engine, "lm_state_post"):
        # One solve ahead of the host: solve n+1 is queued (kernels and both all-reduces) before
        # the control block of solve n is read, so the device never idles while the host
        # launches.  After termination every kernel of the extra solve is a no-op and the extra
        # all-reduces sum buffers that nobody reads again; `done` is identical on all ranks (it
        # derives from all-reduced data only: the cost and the summed singular flag), so all ranks
        # issue the same collectives.
        enqueue_solve()
        engine.lm_state_post(0)
        n = 1
        while True:
            enqueue_solve()
            engine.lm_state_post(n & 1)
            st = engine.lm_state_wait((n & 1) ^ 1)
            n += 1
            if st.done:
                engine.lm_state_wait((n & 1) ^ 1)  # drain the speculative solve
                return st
    while True:
        enqueue_solve()
        st = engine.lm_state()
        if on_state is not None:
            on_state(st)
        if st.done:
            if hasattr(engine, "lm_state_post"):
                # ranks may mix this loop with the one above (e.g. is_debug on rank 0 only): issue
                # the same trailing pair of (no-op) collectives so the sequences match
                enqueue_solve()
                engine.lm_state()
            return st


def run_sharded(adjuster, scale_factor, delta_tol, max_iter, is_debug):
    """``BundleAdjuster.optimize`` body when a process group was given."""
    import torch.distributed as dist

    from . import _cabi

    eng = adjuster._engine
    rank = dist.get_rank(adjuster._group)
    seen = {"count": 0, "init_logged": False}
    if is_debug:
        adjuster._log.clear()

    def on_state(st):
        if is_debug and not seen["init_logged"]:
            # entry 0 is the initial state; it is still intact until the first accept
            seen["init_logged"] = True
        if st.count > seen["count"]:
            seen["count"] = st.count
            rec = eng.lm_records()[-1]
            if rank == 0:
                adjuster._record(rec.E_prev, rec.E, rec.delta, rec.c, rec.solves, rec.count)
            else:
                adjuster.records.append({"E_prev": rec.E_prev, "E": rec.E, "delta": rec.delta,
                                         "c": rec.c, "solves": rec.solves, "count": rec.count})
            if is_debug:
                adjuster._log_state(rec.E)

    if is_debug:
        # the initial cost needs the all-reduce that lm_loop performs; log the state now and
        # patch the cost in afterwards
        adjuster._log_state(float("nan"))
        st = lm_loop(eng, dist, adjuster._group, scale_factor, delta_tol, max_iter,
                     adjuster._max_retries, on_state)
    else:
        # no per-iteration host work: the loop runs one solve ahead of the host and the
        # iteration lines are printed afterwards (same text, as in the single-GPU path)
        st = lm_loop(eng, dist, adjuster._group, scale_factor, delta_tol, max_iter,
                     adjuster._max_retries, None)
        for rec in eng.lm_records():
            if rank == 0:
                adjuster._record(rec.E_prev, rec.E, rec.delta, rec.c, rec.solves, rec.count)
            else:
                adjuster.records.append({"E_prev": rec.E_prev, "E": rec.E, "delta": rec.delta,
                                         "c": rec.c, "solves": rec.solves, "count": rec.count})
    if is_debug and adjuster.records:
        adjuster._log[0]["reprojection_error"] = np.float64(adjuster.records[0]["E_prev"])
    if st.status == _cabi.BA_ERR_SINGULAR:
        raise np.linalg.LinAlgError("Singular matrix")
    if st.status != _cabi.BA_OK:
        raise RuntimeError(f"sharded LM loop stopped with device status {st.status}")
    return st
