"""Which part of a script-style process crashes at interpreter exit?  (round-2 diagnostic)

    python -X faulthandler tools/gpu/exit_probe.py MODE
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
mode = sys.argv[1]
print("mode", mode, flush=True)


def stub_matplotlib():
    class _Anything:
        def __getattr__(self, name):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

        def __iter__(self):
            return iter(())

    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    plt.__getattr__ = lambda name: _Anything()
    plt.fignum_exists = lambda *_: False
    mpl.pyplot = plt
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt


if mode == "ref_only":
    # the unmodified script with the reference's own class: no CUDA anywhere
    import runpy

    stub_matplotlib()
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    runpy.run_path(os.path.join(ROOT, "oracle", "_ref", "euclidiean_reconstruction.py"), run_name="__main__")
elif mode in ("ba_small", "ba_small_debug", "ba_small_torch_first"):
    if mode == "ba_small_torch_first":
        import torch  # noqa: F401
    import contextlib
    import io

    import numpy as np

    import ba_b200

    sc = ba_b200.scenes.make_scene(8, 60, seed=13, visibility=1.0)
    x, vis = sc.dense_x()
    adj = ba_b200.BundleAdjuster(x, sc.X0, sc.K0, sc.R0, sc.t0, f0=sc.f0, axis=sc.axis)
    with contextlib.redirect_stdout(io.StringIO()):
        adj.optimize(2.0, 1e-8, max_iter=5, is_debug=(mode == "ba_small_debug"))
    print("records", len(adj.records), "torch loaded:", "torch" in sys.modules, flush=True)
elif mode == "project_only":
    import numpy as np

    import ba_b200

    sc = ba_b200.scenes.make_scene(8, 60, seed=13, visibility=1.0)
    out = ba_b200.calc_projected_points(sc.X0, sc.K0, sc.R0, sc.t0)
    print("projected", len(out), "torch loaded:", "torch" in sys.modules, flush=True)
elif mode == "ba_small_no_torch":
    # keep torch out of the process entirely (the engine then uses the default stream)
    sys.modules["torch"] = None
    import contextlib
    import io

    import ba_b200

    sc = ba_b200.scenes.make_scene(8, 60, seed=13, visibility=1.0)
    x, vis = sc.dense_x()
    adj = ba_b200.BundleAdjuster(x, sc.X0, sc.K0, sc.R0, sc.t0, f0=sc.f0, axis=sc.axis)
    with contextlib.redirect_stdout(io.StringIO()):
        adj.optimize(2.0, 1e-8, max_iter=5)
    print("records", len(adj.records), flush=True)
print("end of main", flush=True)
