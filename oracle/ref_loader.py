"""TEST INFRASTRUCTURE ONLY — loads the *real* reference modules, in the authoring container only.

`/root/reference` does not exist on the GPU box, so nothing that runs there may import this file.
It is used by `oracle/make_golden.py` (which writes `tests/golden/*.npz`) and by CPU tests that are
skipped when the reference tree is absent.

The reference's hot-path modules load by file path with two shims:
  * `np.int = int`        — `ssrs/movmodel.py:134,137` use the removed alias at import time;
  * empty `richdem` module — `ssrs/layers.py:4` imports it, only the unused `*_richdem_*` functions need it.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("SSRS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "ssrs", "movmodel.py"))


def _load(name: str):
    path = os.path.join(REFERENCE_ROOT, "ssrs", f"{name}.py")
    spec = importlib.util.spec_from_file_location(f"_ssrs_reference_{name}", path)
    mod = importlib.util.module_from_spec(spec)
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
