"""Run one of the reference's own scripts UNCHANGED on the B200 engine.

    python tools/run_reference_script.py /path/to/reference/euclidiean_reconstruction.py

Puts this repo's package directory (which contains lib/bundle_adjustment.py) before the
reference checkout on sys.path, so `from lib.bundle_adjustment import BundleAdjuster` resolves
to the CUDA engine while every other `lib.*` module is the reference's.  If matplotlib is not
installed a no-op stub is injected (the scripts only plot with it).
"""
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
script = os.path.abspath(sys.argv[1])
sys.path[:0] = [os.path.join(ROOT, "3d-reconstruction-from-multi-view-exp_b200"), os.path.dirname(script), ROOT]

try:
    import matplotlib  # noqa: F401
except ImportError:
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

runpy.run_path(script, run_name="__main__")

# evidence that the script's BundleAdjuster was this package's (written only when asked for)
if os.environ.get("BA_SCRIPT_REPORT"):
    import json

    mod = sys.modules["lib.bundle_adjustment"]
    import ba_b200

    with open(os.environ["BA_SCRIPT_REPORT"], "w") as f:
        json.dump({"script": script,
                   "adjuster_class": f"{mod.BundleAdjuster.__module__}.{mod.BundleAdjuster.__qualname__}",
                   "kernel_launches": ba_b200.submodule("engine").launch_count()}, f)
