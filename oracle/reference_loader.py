"""Import the UNMODIFIED reference (``/root/reference/src``) for golden-vector
generation and differential checks.  Oracle / test infrastructure only; works only
where /root/reference exists (the build container, never the GPU box).

The reference imports four third-party packages at module top that are absent here
(robust_laplacian, meshio, pyvista, matplotlib) but never touches them on the hot
path, so they are replaced by empty stub modules.
"""
import os
import sys
import types

REFERENCE_SRC = "/root/reference/src"
REFERENCE_ROOT = "/root/reference"
_STUBS = ("robust_laplacian", "meshio", "pyvista", "matplotlib", "matplotlib.pyplot")
_FLAT = ("utils", "mesh_helpers", "samplers", "multigrid_model", "corrector_model",
         "config", "Mesh", "diagnostics")


def available():
    return os.path.isdir(REFERENCE_SRC)


def load():
    """Return a namespace with the reference's flat modules as attributes."""
    if not available():
        raise RuntimeError("reference sources not present at %s" % REFERENCE_SRC)
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    clash = [m for m in _FLAT if m in sys.modules
             and not str(getattr(sys.modules[m], "__file__", "")).startswith(REFERENCE_SRC)]
    if clash:
        raise RuntimeError("flat module names already imported from elsewhere: %s "
                           "(load the reference in a fresh process)" % clash)
    sys.path.insert(0, REFERENCE_SRC)
    try:
        import importlib
        ns = types.SimpleNamespace()
        for m in _FLAT[:-1]:
            setattr(ns, m, importlib.import_module(m))
    finally:
        sys.path.remove(REFERENCE_SRC)
    return ns
