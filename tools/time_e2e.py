"""Wall-clock breakdown of the public-API path (construction vs optimize) for one workload."""
import contextlib
import io
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ba_b200  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
sc = ba_b200.scenes.make_scene(**ba_b200.scenes.CONFIGS[name])
torch.cuda.init()
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    adj = ba_b200.BundleAdjuster.from_observations(sc.obs_ptr, sc.obs_cam, sc.obs_xy, sc.X0, sc.K0, sc.R0, sc.t0,
                                                   f0=sc.f0, axis=sc.axis, dense=sc.dense)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        adj.optimize(2.0, -1.0, max_iter=K)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    adj.engine.close()
    t3 = time.perf_counter()
    print(f"rep {rep}: construct {1e3 * (t1 - t0):.1f} ms, optimize({K}) {1e3 * (t2 - t1):.1f} ms, close {1e3 * (t3 - t2):.1f} ms")
