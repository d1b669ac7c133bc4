#!/usr/bin/env python
"""Benchmark of the bundle-adjustment hot path (BASELINE.json metric: observations/s of
Levenberg-Marquardt iterations, i.e. visible observations x accepted LM iterations / second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl reference]

One "step" = one accepted LM iteration (linearise, damped Schur solve(s), trial cost, accept)
of the named synthetic scene, started from the same perturbed-ground-truth state every time.
N > 1 is launched by torchrun (one rank per GPU); each rank owns its own shard of points
(weak scaling: the per-GPU shard is the named config, cameras are shared) and the ranks
all-reduce the partial reduced system and the trial cost once per inner solve.

`value`   device-resident: inputs already in HBM, CUDA events around exactly K iterations.
`e2e`     through the public class (BundleAdjuster.from_observations(...).optimize(...)) with
          pinned HOST buffers: construction, H2D of observations and state, K iterations, D2H
          of the result, wall clock between device synchronisations.
`roofline` the dominant kernel (K3, the FP64-tensor-core SYRK): algorithmic flops / launch
          over its CUDA-event duration, against the DMMA peak measured live on this GPU.
`cpu_baseline` the CPU oracle port (oracle/ba_oracle.py, NumPy/BLAS) on a bounded sample.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "BA LM observations/s (visible observations x accepted LM iterations per second)"
UNIT = "observations/s"
SCALE, TOL_NEVER = 2.0, -1.0  # optimize(2.0, ...) as in the reference script; tol<0: never stop early

# bounded CPU samples (points of the named scene, cameras unchanged, LM iterations)
CPU_SAMPLE = {"c2": (10_000, 2), "c3": (1_500, 1), "c4": (250, 1), "c5": (250, 1)}
E2E_REPS = 3
CHUNK = 10  # LM iterations per run from the perturbed start (see run_iters); c5 is ONE run (convergence)
REF_STEP_SAMPLE = {"c2": 2_000, "c3": 800, "c4": 150, "c5": 150}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "collective"],
                    help="N > 1: how the ranks sum the partial reduced system and the cost: the library's "
                         "own kernels over NVLink peer memory (auto/peer) or torch.distributed all-reduces")
    ap.add_argument("--profile-ranks", action="store_true",
                    help="N > 1: re-run the K iterations with per-phase CUDA-event timing on every rank "
                         "(eager launches) and add rank 0's phase table; `comm` includes waiting for peers")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: the named scene's points are split over the ranks "
                         "(default: weak, every rank owns a full named scene)")
    return ap.parse_args()


def workload_config(name: str) -> dict:
    import ba_b200

    cfg = dict(ba_b200.scenes.CONFIGS[name])
    return cfg


def describe(name: str, cfg: dict, world: int, nobs_total: int, exchange: str = "") -> dict:
    vis = cfg.get("visibility", 1.0)
    return {
        "workload": f"{name}: synthetic {cfg['n_cams']} cameras x {cfg['n_points']} points per GPU, "
                    f"{'full' if vis >= 1.0 else f'{vis:.0%} random'} visibility"
                    + (f", {cfg['outlier_frac']:.0%} outliers" if cfg.get("outlier_frac") else ""),
        "n_cams": cfg["n_cams"], "n_points_per_gpu": cfg["n_points"], "observations_total": nobs_total,
        "unknowns_reduced": 9 * cfg["n_cams"] - 7,
        "lm": "optimize(scale_factor=2.0), one step = one accepted LM iteration; K steps = LM runs of <= 10 "
              "iterations from the perturbed start (the state is reset on the device between runs)",
        "parallelism": f"points sharded over {world} GPU(s), cameras replicated; {exchange}" if world > 1 else "single GPU",
        "l2": "no flush: iterations are data-dependent; per-iteration working set "
              "(Jacobian rows + Y) exceeds the 126 MB L2 for c2 and larger",
    }


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is under load."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for ts, r in self.rows if len(r) >= 8 and (t0 is None or t0 <= ts <= t1)]
        if len(rows) < 3:
            rows = [r for _, r in self.rows if len(r) >= 8]
        sm, reasons, smax = [], set(), None
        for r in rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def oracle_run(sc, n_points: int, iters: int):
    """Time `iters` LM iterations of the CPU oracle on the first `n_points` points of `sc`."""
    from oracle import ba_oracle as O

    n_points = min(n_points, sc.n_points)
    hi = int(sc.obs_ptr[n_points])
    counts = np.diff(sc.obs_ptr[: n_points + 1])
    obs = O.ObsList(n_points, sc.n_cams, np.repeat(np.arange(n_points), counts),
                    sc.obs_cam[:hi].astype(np.int64), sc.obs_xy[:hi], sc.obs_ptr[: n_points + 1].copy())
    ora = O.OracleBundleAdjuster(None, sc.X0[:n_points], sc.K0, sc.R0, sc.t0, f0=sc.f0, axis=sc.axis, obs=obs)
    chunk = max(64, min(4096, int(2.5e8 // (27 * sc.n_cams * 8))))
    t0 = time.perf_counter()
    ora.optimize(SCALE, TOL_NEVER, max_iter=iters, verbose=False, chunk_points=chunk)
    dt = time.perf_counter() - t0
    done = len(ora.trace) - 1
    return hi * done / dt, dt, hi, done, float(np.sqrt(ora.trace[-1]["E"] / hi))


def measured_traffic(workload: str):
    """DRAM bytes per launch of the workload's dominant kernel from the committed `ncu --set full`
    capture (profiles/roofline_traffic.json names the summary file each figure comes from)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info

        return max([d.get("num_threads", 1) for d in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank: int):
    """The reference arm: the CPU implementation of the path on the box's host cores.  The
    reference itself is Python and does not travel to the GPU box, so this is its validated
    observation-list port (oracle/ba_oracle.py; kind "port")."""
    if rank != 0:
        return
    import ba_b200

    cfg = workload_config(args.workload)
    n_sample = REF_STEP_SAMPLE[args.workload]
    sc = ba_b200.scenes.make_scene(**{**cfg, "n_points": n_sample})
    if args.warmup > 0:
        oracle_run(sc, n_sample, min(args.warmup, 1))
    value, dt, nobs, done, rms = oracle_run(sc, n_sample, args.steps)
    cores = blas_threads()
    sample = (f"{done} LM iterations on the first {n_sample} points x {cfg['n_cams']} cameras "
              f"({nobs} observations) of {args.workload}; NumPy/BLAS with {cores} threads")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / max(done, 1) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": describe(args.workload, cfg, 1, nobs),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "final_rms": rms, "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
def run_cuda(args, rank: int, world: int, local_rank: int):
    import torch

    import ba_b200

    torch.cuda.set_device(local_rank)
    dist = None
    group = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        group = dist.group.WORLD
    engine_mod = ba_b200.submodule("engine")
    sharded = ba_b200.submodule("sharded")

    cfg = workload_config(args.workload)
    if args.strong and world > 1:
        cfg["n_points"] = cfg["n_points"] // world
    sc = ba_b200.scenes.make_scene(**cfg, point_stream=rank)
    K, W = args.steps, max(args.warmup, 0)

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy()

    h_ptr, h_xy = pinned(sc.obs_ptr), pinned(sc.obs_xy)
    h_cam = None if sc.dense else pinned(sc.obs_cam)
    h_X0, h_K0, h_R0, h_t0 = pinned(sc.X0), pinned(sc.K0), pinned(sc.R0), pinned(sc.t0)

    def make_adjuster(gauge_on_device=False):
        return ba_b200.BundleAdjuster.from_observations(
            h_ptr, h_cam, h_xy, h_X0, h_K0, h_R0, h_t0, f0=sc.f0, axis=sc.axis, dense=sc.dense,
            device=local_rank, process_group=group, exchange=args.exchange,
            gauge_on_device=gauge_on_device)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run_iters(adj, n):
        """Exactly n accepted LM iterations, as LM runs of at most CHUNK iterations, each started
        from the adjuster's initial state (kept on the device).  A run from the perturbed start
        needs ~14 iterations to reach the noise floor; there the accept test E_ > E is decided
        by the last bits of the cost sum, so the number of retries would depend on the summation
        order (hence on the GPU count) instead of on the work.  Chunks stay clear of the floor."""
        eng = adj.engine
        if getattr(adj, "_dev_init", None) is None:
            adj._dev_init = [torch.from_numpy(np.ascontiguousarray(a)).cuda(local_rank)
                             for a in (adj._X, adj._R, adj._t, adj._f, adj._u)]
        chunk = n if args.workload == "c5" else CHUNK  # c5: the 50-iteration convergence run
        done, solves, st = 0, 0, None
        while done < n:
            m = min(chunk, n - done)
            eng.set_state(*adj._dev_init)
            if world > 1 and not adj._peer_exchange:
                st = sharded.lm_loop(eng, dist, group, SCALE, TOL_NEVER, m)
            else:
                _, st = eng.lm_run(SCALE, TOL_NEVER, m)
            assert st.count == m, f"ran {st.count} iterations instead of {m}"
            done += m
            solves += st.solves
        st.solves = solves
        st.count = done
        return st

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- device-resident arm ---------------------------------------------------------------
    adj = make_adjuster()
    if W > 0:
        run_iters(adj, W)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = engine_mod.launch_count()
    t_wall0 = time.perf_counter()
    ev0.record()
    st = run_iters(adj, K)
    ev1.record()
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    launches = engine_mod.launch_count() - launches0
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    nobs_t = torch.tensor([float(sc.nobs)], dtype=torch.float64, device="cuda")
    solves_t = torch.tensor([float(st.solves)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(nobs_t, op=dist.ReduceOp.SUM)
    ms, nobs_total = float(ms.item()), int(nobs_t.item())
    assert st.count == K, f"ran {st.count} iterations instead of {K}"
    value = nobs_total * K / (ms * 1e-3)
    final_rms = float(np.sqrt(st.E / nobs_total))

    out = None
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "strong" if (args.strong and world > 1) else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": describe(args.workload, cfg, world, nobs_total,
                               "sums over NVLink peer memory by the library's kernels (CUDA-graph loop)"
                               if adj._peer_exchange else "two NCCL all-reduces per solve"),
            "lm_iterations_per_s": K / (ms * 1e-3), "inner_solves": int(st.solves), "final_rms": final_rms,
            "gpu_launches": int(launches),
        }

    # ---- roofline of the dominant kernel (rank 0, profiled re-run of the same K iterations) --
    if rank == 0 and world == 1:
        eng = adj.engine
        eng.profile_enable(True)
        eng.profile_reset()
        st_prof = run_iters(adj, K)
        prof = eng.profile()
        eng.profile_enable(False)
        n_red = 9 * sc.n_cams - 7
        counts = np.diff(sc.obs_ptr)
        flops = float(np.sum(3.0 * (9.0 * counts) * (9.0 * counts + 1.0))) if not sc.dense else \
            3.0 * sc.n_points * n_red * (n_red + 1.0)
        peak = engine_mod.fp64_peak(local_rank, True)
        tot = sum(v["ms"] for k, v in prof.items() if k in ("k1", "k2", "k3", "k4", "cost", "other"))
        if sc.dense and prof["syrk"]["launches"] > 0:
            avg_ms = prof["syrk"]["ms"] / prof["syrk"]["launches"]
            achieved = flops / (avg_ms * 1e-3) / 1e12
            out["roofline"] = {
                "kernel": "syrk_dmma_kernel (K3, Schur SYRK on FP64 tensor cores)",
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": None,
                "algorithmic_flops_per_launch": flops, "avg_launch_ms": avg_ms,
                "share_of_step": prof["syrk"]["ms"] / tot if tot > 0 else None,
                "peak_source": "measured live: register-resident DMMA.8x8x4 loop (ba_fp64_peak); "
                               "MEASURED_PEAKS.json has no FP64 figure",
            }
        else:
            k3 = prof["k3"]
            avg_ms = k3["ms"] / max(st_prof.solves, 1)
            achieved = flops / (avg_ms * 1e-3) / 1e12
            out["roofline"] = {
                "kernel": "schur_pairs_kernel + schur_diag_kernel (K3, sparse Schur products, matrix-free pair kernel)",
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": None, "algorithmic_flops_per_launch": flops,
                "avg_launch_ms": avg_ms, "share_of_step": k3["ms"] / tot if tot > 0 else None,
                "peak_source": "measured live: register-resident DMMA.8x8x4 loop (ba_fp64_peak)",
            }
        tr = measured_traffic(args.workload)
        if tr:
            out["roofline"]["traffic"] = tr["dram_bytes_per_launch"]
            out["roofline"]["traffic_source"] = tr["source"]
        out["phase_ms_per_step"] = {k: v["ms"] / K for k, v in prof.items()}
        hbm = None
        try:
            hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
        except Exception:
            pass
        # memory-bound phases against the measured copy bandwidth (K1: 16 B read + 224 B written
        # per observation and linearisation)
        if prof["k1"]["ms"] > 0:
            gbs = sc.nobs * 240.0 * K / (prof["k1"]["ms"] * 1e-3) / 1e9
            out["k1_hbm"] = {"achieved_gbs": gbs, "peak_gbs": hbm or 6650.0,
                             "peak_source": "MEASURED_PEAKS.json" if hbm else "fallback",
                             "frac": gbs / (hbm or 6650.0), "bytes_per_obs": 240}

    if world > 1 and args.profile_ranks and adj._peer_exchange:
        eng = adj.engine
        eng.profile_enable(True)
        eng.profile_reset()
        barrier()
        run_iters(adj, K)
        prof = eng.profile()
        eng.profile_enable(False)
        barrier()
        if rank == 0:
            out["phase_ms_per_step"] = {k: v["ms"] / K for k, v in prof.items()}

    # the device-resident engine is released first: the end-to-end arm re-creates one, as a
    # caller that adjusts scene after scene would
    adj.engine.close()
    # ---- end-to-end arm (host buffers in, host results out) ---------------------------------
    # Each repetition is the complete user call for K steps; the median of E2E_REPS wall times is
    # reported (all samples are listed), since a single cold call is dominated by allocator noise.
    state_bytes = (3 * sc.n_points + 15 * sc.n_cams) * 8
    n_calls = len([1 for _ in range(0, K, K if args.workload == "c5" else CHUNK)])
    h2d = n_calls * (h_xy.nbytes + h_ptr.nbytes + (0 if h_cam is None else h_cam.nbytes) + state_bytes)
    d2h = n_calls * state_bytes + K * 40 + (st.solves + n_calls) * 88
    e2e_samples = []
    e2e_chunk = K if args.workload == "c5" else CHUNK
    calls = [min(e2e_chunk, K - k0) for k0 in range(0, K, e2e_chunk)]
    for _ in range(E2E_REPS):
        barrier()
        t0 = time.perf_counter()
        for m in calls:
            adj2 = make_adjuster(gauge_on_device=True)
            with contextlib.redirect_stdout(io.StringIO()):
                Xr, Kr, Rr, tr = adj2.optimize(SCALE, TOL_NEVER, max_iter=m)
            assert len(adj2.records) == m
            adj2.engine.close()
        torch.cuda.synchronize()
        e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e_samples.append(float(e2e_t.item()))
    e2e_s = float(np.median(e2e_samples))

    if rank == 0:
        out["e2e"] = {"value": nobs_total * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d / K,
                      "d2h_bytes_per_step": d2h / K, "ms_per_step": e2e_s / K * 1e3,
                      "samples_ms_per_step": [t / K * 1e3 for t in e2e_samples],
                      "what": f"{n_calls} complete user call(s) of <= {K if args.workload == 'c5' else CHUNK} iterations each: "
                              "BundleAdjuster.from_observations(pinned host arrays, gauge_on_device=True)"
                              ".optimize(max_iter=...): engine creation (device memory from the library's "
                              "retained pool), H2D of observations and state, gauge normalisation, LM "
                              f"iterations, de-normalisation, D2H of X/K/R/t; median of {E2E_REPS} repetitions"}

    if rank == 0:
        out["clocks"] = sampler.stop(t_wall0, t_wall1)
        if world == 1 and not args.no_cpu_baseline:
            n_pts, iters = CPU_SAMPLE[args.workload]
            cval, cdt, cobs, cdone, _ = oracle_run(sc, n_pts, iters)
            cores = blas_threads()
            out["cpu_baseline"] = {
                "value": cval, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{cdone} LM iteration(s) of the CPU oracle port on the first {min(n_pts, sc.n_points)} "
                          f"points x {sc.n_cams} cameras ({cobs} observations) in {cdt:.1f} s; NumPy/BLAS, "
                          f"{cores} threads of {os.cpu_count()} cores"}
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # not under torchrun: re-launch one rank per GPU
        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_cuda(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
