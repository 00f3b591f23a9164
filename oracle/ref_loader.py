"""TEST INFRASTRUCTURE ONLY — loads the *real* reference modules, in the authoring container only.

`/root/reference` does not exist on the GPU box.  There the loader falls back to `oracle/_ref/` — the two hot-path
modules staged, unmodified and git-ignored, by `stage_reference()` during `__graft_entry__.build()` — which is what
bench.py's CPU baseline (kind "reference") runs.  Also used by `oracle/make_golden*.py` (which write
`tests/golden/*.npz`).

The reference's hot-path modules load by file path with two shims:
  * `np.int = int`        — `ssrs/movmodel.py:134,137` use the removed alias at import time;
  * empty `richdem` module — `ssrs/layers.py:4` imports it, only the unused `*_richdem_*` functions need it.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# `oracle/_ref/` (git-ignored, travels to the GPU box with the snapshot): the two hot-path modules of the reference,
# byte for byte, placed there by `stage_reference()` from __graft_entry__.build() while /root/reference is present.
STAGED_ROOT = os.path.join(HERE, "_ref")
HOT_PATH_MODULES = ("layers.py", "movmodel.py")


def _root() -> str:
    env = os.environ.get("SSRS_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile(os.path.join("/root/reference", "ssrs", "movmodel.py")):
        return "/root/reference"
    return STAGED_ROOT


REFERENCE_ROOT = _root()


def available() -> bool:
    return all(os.path.isfile(os.path.join(REFERENCE_ROOT, "ssrs", m)) for m in HOT_PATH_MODULES)


def stage_reference(src_root: str = "/root/reference") -> bool:
    """Copies the reference's two hot-path modules, unmodified, to oracle/_ref/ssrs/ so that the CPU baseline of
    bench.py can run the reference's OWN code on the GPU box (where /root/reference does not exist).  The directory is
    git-ignored: reference sources never enter the repository's history.  Returns False when the tree is absent."""
    import shutil
    if not os.path.isfile(os.path.join(src_root, "ssrs", "movmodel.py")):
        return False
    dst = os.path.join(STAGED_ROOT, "ssrs")
    os.makedirs(dst, exist_ok=True)
    for m in HOT_PATH_MODULES:
        shutil.copyfile(os.path.join(src_root, "ssrs", m), os.path.join(dst, m))
    with open(os.path.join(STAGED_ROOT, "README"), "w") as f:
        f.write("Unmodified copies of /root/reference/ssrs/{layers,movmodel}.py staged by oracle/ref_loader.py for the\n"
                "CPU baseline (bench.py, kind \"reference\").  Not part of the repository (see .gitignore).\n")
    return True


def _load(name: str):
    path = os.path.join(REFERENCE_ROOT, "ssrs", f"{name}.py")
    spec = importlib.util.spec_from_file_location(f"_ssrs_reference_{name}", path)
    mod = importlib.util.module_from_spec(spec)
    # registered under its name: multiprocess/dill pickles a closure's modules by name, and forked pool workers inherit
    # sys.modules (oracle/ref_cpu.py runs the reference's `pool.map(lambda ...)` pattern)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def load_reference():
    """Returns (layers, movmodel) modules of the unmodified reference."""
    if "mods" not in _cache:
        if not available():
            raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
        if not hasattr(np, "int"):
            np.int = int  # noqa: NPY001  shim for movmodel.py:134
        sys.modules.setdefault("richdem", types.ModuleType("richdem"))
        _cache["mods"] = (_load("layers"), _load("movmodel"))
    return _cache["mods"]
