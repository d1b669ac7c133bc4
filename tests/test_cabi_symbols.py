"""The C-ABI library loads and exports every symbol include/ba_b200.h declares; the ctypes
binding covers exactly that set.  No compute calls (CPU only)."""
import ctypes
import os
import re

import pytest

import ba_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ba_b200.h")


@pytest.fixture(scope="module")
def cabi():
    import __graft_entry__

    __graft_entry__.build()
    return ba_b200.submodule("_cabi")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ba_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for must in ("ba_create", "ba_destroy", "ba_set_observations", "ba_set_state", "ba_get_state",
                 "ba_cost", "ba_linearize", "ba_build_reduced", "ba_solve_trial", "ba_lm_run",
                 "ba_lm_phase_reduce", "ba_lm_phase_solve", "ba_lm_phase_decide"):
        assert must in names


def test_library_exports_every_declared_symbol(cabi):
    lib = ctypes.CDLL(cabi.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in ba_b200.h but not exported"


def test_binding_matches_header(cabi):
    assert sorted(cabi.SIGNATURES) == declared_functions()
    lib = cabi.load()
    assert lib.ba_version() == 1


def test_struct_layouts_match_header(cabi):
    # sizes implied by the header's field lists (natural alignment)
    assert ctypes.sizeof(cabi.Problem) == 40
    assert ctypes.sizeof(cabi.LMState) == 6 * 8 + 10 * 4
    assert ctypes.sizeof(cabi.IterRecord) == 4 * 8 + 2 * 4


def test_fails_loudly_without_a_device(cabi):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        ba_b200.Engine(10, 3, 30, 1.0, "x-up_z-forward", True)
    # argument errors surface as ValueError even before a device is needed
    with pytest.raises(ValueError):
        ba_b200.Engine(10, 3, 30, 1.0, "z-up", True)


def test_product_code_does_not_import_the_oracle():
    pkg = ba_b200.PACKAGE_DIR
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "ba_oracle" not in text and "from oracle" not in text and "import oracle" not in text, fn
