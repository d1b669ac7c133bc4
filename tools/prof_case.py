"""Small driver for ncu / compute-sanitizer runs: one scene, a few phases.

    python tools/prof_case.py --cams 200 --points 20000 [--vis 1.0] [--solves 2] [--lm 0]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import ba_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cams", type=int, default=200)
ap.add_argument("--points", type=int, default=20000)
ap.add_argument("--vis", type=float, default=1.0)
ap.add_argument("--solves", type=int, default=2)
ap.add_argument("--lm", type=int, default=0, help="run this many LM iterations instead of single phases")
args = ap.parse_args()

sc = ba_b200.scenes.make_scene(args.cams, args.points, seed=1, visibility=args.vis)
adj = ba_b200.BundleAdjuster.from_observations(sc.obs_ptr, sc.obs_cam, sc.obs_xy, sc.X0, sc.K0, sc.R0, sc.t0,
                                               f0=sc.f0, axis=sc.axis, dense=sc.dense)
eng = adj.engine
if args.lm > 0:
    recs, st = eng.lm_run(2.0, -1.0, args.lm)
    print("lm", st.count, st.solves, st.E)
else:
    eng.linearize()
    for k in range(args.solves):
        c = 1e-4 * (2 ** k)
        eng.build_reduced(c)
        eng.solve_trial(c)
    print("cost", eng.cost_values())
