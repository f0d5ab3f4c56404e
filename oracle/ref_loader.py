"""Loader of the REFERENCE's own model files -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's hot path is three torch-only Python files (``src/models/Hang2020.py``, ``src/models/year.py``,
``src/models/metadata.py``).  They can be imported as they lie under ``/root/reference`` in the build container, but that
tree does not travel to the GPU box.  ``build_ref()`` (run by ``__graft_entry__.build()`` in the build container, recipe
below) copies those three files -- unmodified -- into ``oracle/_ref/src/models/``: a git-ignored directory (the copies never
enter the history) that is NOT gpurun-ignored, so it ships to the GPU box next to the built ``.so`` files.  There
``bench.py --impl reference`` times the reference's own module (``cpu_baseline.kind = "reference"``) and the TreeModel-shim
test compares against it.  When neither tree exists every caller falls back to the oracle port
(``oracle/hang2020_oracle.py``), which is pinned to the reference's outputs by ``tests/golden``.

``src/models/metadata.py`` imports ``src.main`` (Lightning, geopandas, deepforest ... none installed) only to subclass
``TreeModel`` in ``MetadataModel`` (metadata.py:47-89, out of scope); ``src/models/year.py`` imports ``torchmetrics`` without
using it (year.py:6).  Both are satisfied with empty stand-in modules in ``sys.modules`` while the files load
(SURVEY.md 8c); the arithmetic of ``metadata`` / ``metadata_sensor_fusion`` / ``learned_ensemble`` is untouched.
"""
from __future__ import annotations

import importlib
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_TREE = "/root/reference"
REF_COPY = os.path.join(HERE, "_ref")
FILES = ("src/models/Hang2020.py", "src/models/year.py", "src/models/metadata.py")


def build_ref(verbose: bool = False) -> bool:
    """Recipe for ``oracle/_ref``: copy the three model files from /root/reference (build container only)."""
    if not os.path.isdir(REFERENCE_TREE):
        return os.path.exists(os.path.join(REF_COPY, FILES[0]))
    for rel in FILES:
        dst = os.path.join(REF_COPY, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REFERENCE_TREE, rel), dst)
    with open(os.path.join(REF_COPY, "README"), "w") as f:
        f.write("Unmodified copies of weecology/DeepTreeAttention model files made by oracle/ref_loader.build_ref(); git-ignored.\n")
    if verbose:
        print("oracle/_ref: copied", ", ".join(FILES))
    return True


def reference_root():
    """Directory that holds ``src/models/Hang2020.py`` of the reference, or None."""
    for root in (REFERENCE_TREE, REF_COPY):
        if os.path.exists(os.path.join(root, FILES[0])):
            return root
    return None


_cache = {}


def load(which: str = "Hang2020"):
    """The reference module ``src.models.<which>`` (``Hang2020`` | ``year`` | ``metadata``), or None when no reference tree
    is available.  Loaded under a private package name so it cannot collide with anything called ``src``."""
    if which in _cache:
        return _cache[which]
    root = reference_root()
    if root is None:
        return None
    pkg_root, pkg_models = "_dta_ref_src", "_dta_ref_src.models"
    if pkg_root not in sys.modules:
        p = types.ModuleType(pkg_root); p.__path__ = [os.path.join(root, "src")]
        m = types.ModuleType(pkg_models); m.__path__ = [os.path.join(root, "src", "models")]
        sys.modules[pkg_root], sys.modules[pkg_models] = p, m
        p.models = m
    base = load("Hang2020") if which != "Hang2020" else None
    # the files say "from src.models ... / from src import main": alias the private package as `src` only while loading
    saved = {k: sys.modules.get(k) for k in ("src", "src.models", "src.main", "torchmetrics", "src.models.Hang2020")}
    try:
        sys.modules["src"], sys.modules["src.models"] = sys.modules[pkg_root], sys.modules[pkg_models]
        stub_main = types.ModuleType("src.main")
        stub_main.TreeModel = type("TreeModel", (), {})
        sys.modules["src.main"] = stub_main
        sys.modules[pkg_root].main = stub_main
        if saved["torchmetrics"] is None:
            sys.modules["torchmetrics"] = types.ModuleType("torchmetrics")
        if base is not None:        # the siblings import Hang2020 as src.models.Hang2020: keep ONE module object
            sys.modules["src.models.Hang2020"] = base
        mod = importlib.import_module(f"{pkg_models}.{which}")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cache[which] = mod
    return mod


if __name__ == "__main__":
    ok = build_ref(verbose=True)
    print("reference root:", reference_root(), "built" if ok else "unavailable")
