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

    def _plt_getattr(name):
        # a stub must not answer dunder look-ups (__file__, __path__, __spec__ ...): code that walks
        # sys.modules -- inspect.getmodule during `import torch`, for one -- would take the answers
        # for real ones
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    plt.__getattr__ = _plt_getattr
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
elif mode == "script_trace":
    # the unmodified script on the engine, with every call into the script / the reference modules /
    # this package printed as it happens: the last line before a crash localises it
    import runpy

    pkg = os.path.join(ROOT, "3d-reconstruction-from-multi-view-exp_b200")
    ref = os.path.join(ROOT, "oracle", "_ref")
    depth = [0]

    def tracer(frame, event, arg):
        fn = frame.f_code.co_filename
        if event in ("call", "return") and (fn.startswith(ref) or fn.startswith(pkg)):
            if event == "call":
                print("  " * depth[0] + f"> {os.path.relpath(fn, ROOT)}:{frame.f_code.co_name}", flush=True)
                depth[0] += 1
            else:
                depth[0] -= 1
        elif event == "c_call" and getattr(arg, "__name__", "").startswith("ba_"):
            print("  " * depth[0] + f"c> {arg.__name__}", flush=True)

    stub_matplotlib()
    sys.path[:0] = [pkg, ref]
    sys.setprofile(tracer)
    try:
        runpy.run_path(os.path.join(ref, "euclidiean_reconstruction.py"), run_name="__main__")
    finally:
        sys.setprofile(None)
    print("script returned", flush=True)
    import gc

    gc.collect()
    print("gc done", flush=True)
print("end of main", flush=True)
