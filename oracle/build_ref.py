"""Recipe for ``oracle/_ref/``: a byte-for-byte copy of the UNMODIFIED reference, made at build time.

TEST / BASELINE INFRASTRUCTURE ONLY (see ``oracle/ba_oracle.py``).  The reference is pure Python
(nothing to compile), so "building" it means copying its ``lib/`` modules and its two scripts from
``/root/reference`` into ``oracle/_ref/``.  That directory is git-ignored (no reference source ever
enters the history) but not gpurun-ignored, so it travels to the GPU box, where ``/root/reference``
does not exist: ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` leg time the
reference class itself there (``cpu_baseline.kind == "reference"``), and
``tools/run_reference_script.py`` runs ``euclidiean_reconstruction.py`` unchanged on the B200 engine.

    python oracle/build_ref.py            # no-op when /root/reference is absent

A manifest with the SHA-256 of every copied file is written next to the copies; ``verify()``
re-checks it so a hand-edited copy is never timed as "the reference".
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("BA_REFERENCE_DIR", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
SCRIPTS = ("euclidiean_reconstruction.py", "affine_reconstruction.py")


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(verbose: bool = False) -> str | None:
    """Copy the reference's path modules; returns the destination (None if there is no source)."""
    lib_src = os.path.join(REF_SRC, "lib")
    if not os.path.isdir(lib_src):
        return REF_DST if available() else None
    os.makedirs(os.path.join(REF_DST, "lib"), exist_ok=True)
    manifest = {}
    for name in sorted(os.listdir(lib_src)):
        if name.endswith(".py"):
            shutil.copyfile(os.path.join(lib_src, name), os.path.join(REF_DST, "lib", name))
            manifest["lib/" + name] = _sha(os.path.join(lib_src, name))
    for name in SCRIPTS:
        if os.path.isfile(os.path.join(REF_SRC, name)):
            shutil.copyfile(os.path.join(REF_SRC, name), os.path.join(REF_DST, name))
            manifest[name] = _sha(os.path.join(REF_SRC, name))
    with open(os.path.join(REF_DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_SRC, "files": manifest}, f, indent=1)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} files copied from {REF_SRC}")
    return REF_DST


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DST, "lib", "bundle_adjustment.py")) and \
        os.path.isfile(os.path.join(REF_DST, "MANIFEST.json"))


def verify() -> bool:
    """True when every copied file still has the hash recorded at build time."""
    if not available():
        return False
    with open(os.path.join(REF_DST, "MANIFEST.json")) as f:
        files = json.load(f)["files"]
    return all(os.path.isfile(os.path.join(REF_DST, k)) and _sha(os.path.join(REF_DST, k)) == v
               for k, v in files.items())


def load_reference_class():
    """The reference's own ``BundleAdjuster`` (``lib/bundle_adjustment.py:10``), imported from the
    copy under a private module name so that it can never be confused with the shadow modules of
    the product package."""
    import importlib.util
    import sys
    import types

    if not verify():
        raise RuntimeError("oracle/_ref is missing or was modified: run `python oracle/build_ref.py` "
                           "in a container that has /root/reference")
    # the reference imports `from lib.utils import ...` (lib/bundle_adjustment.py:7): give it a
    # private namespace package that resolves to the copy only
    pkg = types.ModuleType("_ba_ref_lib")
    pkg.__path__ = [os.path.join(REF_DST, "lib")]
    saved = {k: sys.modules.get(k) for k in ("lib", "lib.utils", "lib.bundle_adjustment")}
    try:
        lib = types.ModuleType("lib")
        lib.__path__ = [os.path.join(REF_DST, "lib")]
        sys.modules["lib"] = lib
        for k in ("lib.utils", "lib.bundle_adjustment"):
            sys.modules.pop(k, None)
        spec = importlib.util.spec_from_file_location("lib.bundle_adjustment",
                                                      os.path.join(REF_DST, "lib", "bundle_adjustment.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["lib.bundle_adjustment"] = mod
        spec.loader.exec_module(mod)
        return mod.BundleAdjuster
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


if __name__ == "__main__":
    dst = build(verbose=True)
    print(dst if dst else f"{REF_SRC} not present: nothing to do")
