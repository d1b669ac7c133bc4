import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


SMALL_CASES = ["small_dense_xup", "small_dense_xright", "small_sparse_xup",
               "small_sparse_xright", "small_flip_xup"]
RUN_CASES = ["c1_euclid", "affine_script"] + SMALL_CASES + ["mid_dense_xup", "mid_sparse_xup"]


def case_inputs(g):
    """(x, vis-or-None, X0, K0, R0, t0, axis, f0) of a golden case."""
    vis = g.get("vis")
    if vis is not None and vis.all():
        vis = None
    return g["x"], vis, g["X0"], g["K0"], g["R0"], g["t0"], str(g["axis"]), float(g["f0"])
