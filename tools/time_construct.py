"""Wall-clock breakdown of engine construction for one workload (host side and C ABI calls)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ba_b200  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
sc = ba_b200.scenes.make_scene(**ba_b200.scenes.CONFIGS[name])
gauge = ba_b200.submodule("gauge")
Engine = ba_b200.submodule("engine").Engine
torch.cuda.init()
torch.cuda.synchronize()
for rep in range(4):
    t = [time.perf_counter()]
    X, R, tt = gauge.normalize(sc.X0, sc.R0, sc.t0, sc.axis)
    t.append(time.perf_counter())
    eng = Engine(sc.n_points, sc.n_cams, sc.nobs, sc.f0, sc.axis, sc.dense, 0)
    t.append(time.perf_counter())
    eng.set_observations(sc.obs_ptr, sc.obs_cam, sc.obs_xy)
    t.append(time.perf_counter())
    eng.set_state(X, R, tt, sc.K0[:, 0, 0].copy(), sc.K0[:, :2, 2].copy())
    t.append(time.perf_counter())
    torch.cuda.synchronize()
    t.append(time.perf_counter())
    eng.lm_run(2.0, -1.0, 1)
    t.append(time.perf_counter())
    eng.lm_run(2.0, -1.0, 1)
    t.append(time.perf_counter())
    eng.close()
    t.append(time.perf_counter())
    names = ["normalize", "create", "set_obs", "set_state", "sync", "lm_run#1(1 it)", "lm_run#2(1 it)", "close"]
    print(f"rep {rep}: " + ", ".join(f"{n} {1e3 * (b - a):.2f}" for n, a, b in zip(names, t, t[1:])))
