"""Compile the plain-C parts of the CPU oracle (TEST / BASELINE INFRASTRUCTURE ONLY).

    python oracle/build_c.py          ->  oracle/_ref/libba_oracle.so

gcc -O3 -fopenmp on ``oracle/ba_schur_sparse.c`` (the sparsity-aware Schur reduction that is the
CPU denominator of the 1000-camera configurations).  The output directory is git-ignored and
travels to the GPU box with the snapshot; the same image (with gcc) runs there, so ``load()``
rebuilds the library if it is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ba_schur_sparse.c")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libba_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        raise RuntimeError("no C compiler for the oracle's C restatement")
    cmd = [cc, "-O3", "-fopenmp", "-mavx2", "-mfma", "-fPIC", "-shared", "-std=c11", SRC, "-o", OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed on ba_schur_sparse.c:\n" + res.stderr)
    return OUT


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.ba_oracle_threads.restype = C.c_int
        lib.ba_oracle_schur_sparse.restype = C.c_int64
        lib.ba_oracle_schur_sparse.argtypes = [C.c_int64, C.c_int32] + [C.c_void_p] * 7
        _lib = lib
    return _lib


if __name__ == "__main__":
    print(build(force=True))
