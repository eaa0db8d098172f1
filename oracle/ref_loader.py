"""Import the UNMODIFIED reference (baegwangbin/MaGNet) for the checker / baseline legs.  Test infrastructure only:
nothing under magnet_b200/ imports this.

Source of the modules, in order: /root/reference (build container), else the vendored copy in baseline/_ref/ (made by
scripts/vendor_ref.sh, git-ignored, shipped to the GPU box with the snapshot).  Returns None when neither exists.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = ("/root/reference", os.path.join(ROOT, "baseline", "_ref"))
_loaded = None


def _stub_matplotlib():
    """utils/utils.py imports matplotlib (for colour maps of its visualisations) — absent in this image; the hot path
    never touches it, so an empty stand-in is enough to import the module."""
    try:
        import matplotlib  # noqa: F401
        return
    except Exception:
        pass
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


def reference_root():
    for c in CANDIDATES:
        if os.path.isfile(os.path.join(c, "models", "submodules", "homography.py")):
            return c
    return None


def load_reference():
    """-> namespace with .root, .homography, .MAGNET (module), .losses, .utils  — or None if the reference is absent."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = reference_root()
    if root is None:
        return None
    _stub_matplotlib()
    if root not in sys.path:
        sys.path.insert(0, root)
    for name in ("models", "models.submodules", "utils", "data"):    # namespace packages of the reference tree
        if name in sys.modules and not str(getattr(sys.modules[name], "__path__", [""])).count(root):
            del sys.modules[name]
    ns = types.SimpleNamespace(root=root)
    ns.homography = importlib.import_module("models.submodules.homography")
    ns.MAGNET = importlib.import_module("models.MAGNET")
    ns.losses = importlib.import_module("utils.losses")
    ns.utils = importlib.import_module("utils.utils")
    _loaded = ns
    return ns
