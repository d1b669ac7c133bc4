"""Per-phase device times (ms) of a few LM iterations on one scene (profiling mode of the engine)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ba_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cams", type=int, default=200)
ap.add_argument("--points", type=int, default=20000)
ap.add_argument("--vis", type=float, default=1.0)
ap.add_argument("--iters", type=int, default=3)
args = ap.parse_args()
t0 = time.perf_counter()
sc = ba_b200.scenes.make_scene(args.cams, args.points, seed=1, visibility=args.vis)
t1 = time.perf_counter()
adj = ba_b200.BundleAdjuster.from_observations(sc.obs_ptr, sc.obs_cam, sc.obs_xy, sc.X0, sc.K0, sc.R0, sc.t0,
                                               f0=sc.f0, axis=sc.axis, dense=sc.dense)
t2 = time.perf_counter()
eng = adj.engine
eng.lm_run(2.0, -1.0, 1)
eng.set_state(adj._X, adj._R, adj._t, adj._f, adj._u)
eng.profile_enable(True)
eng.profile_reset()
t3 = time.perf_counter()
recs, st = eng.lm_run(2.0, -1.0, args.iters)
t4 = time.perf_counter()
prof = eng.profile()
print(f"scene {args.cams}x{args.points} vis={args.vis} nobs={sc.nobs}: gen {t1 - t0:.1f}s construct {t2 - t1:.2f}s "
      f"run {1e3 * (t4 - t3) / args.iters:.2f} ms/iter solves={st.solves} E={[round(r.E, 4) for r in recs]}")
print({k: round(v["ms"] / args.iters, 3) for k, v in prof.items()})
