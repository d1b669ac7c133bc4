"""time_e2e with the C ABI calls of the constructor timed individually."""
import contextlib
import io
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ba_b200  # noqa: E402

engine_mod = ba_b200.submodule("engine")
Engine = engine_mod.Engine
log = []


def timed(cls, name):
    orig = getattr(cls, name)

    def wrapper(*a, **k):
        t0 = time.perf_counter()
        r = orig(*a, **k)
        log.append((name, 1e3 * (time.perf_counter() - t0)))
        return r

    setattr(cls, name, wrapper)


for n in ("__init__", "set_observations", "set_state", "lm_run", "get_state", "close"):
    timed(Engine, n)

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
sc = ba_b200.scenes.make_scene(**ba_b200.scenes.CONFIGS[name])
torch.cuda.init()
for rep in range(4):
    log.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    adj = ba_b200.BundleAdjuster.from_observations(sc.obs_ptr, sc.obs_cam, sc.obs_xy, sc.X0, sc.K0, sc.R0, sc.t0,
                                                   f0=sc.f0, axis=sc.axis, dense=sc.dense)
    t1 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        adj.optimize(2.0, -1.0, max_iter=K)
    t2 = time.perf_counter()
    adj.engine.close()
    t3 = time.perf_counter()
    print(f"rep {rep}: construct {1e3 * (t1 - t0):.2f} optimize {1e3 * (t2 - t1):.2f} close {1e3 * (t3 - t2):.2f} | "
          + ", ".join(f"{n} {v:.2f}" for n, v in log))
