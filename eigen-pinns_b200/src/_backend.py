"""Locate and import the kernel package (its directory name carries a hyphen, so it is
imported by string) from the flat, reference-shaped modules in this directory."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

PKG = "eigen-pinns_b200"


def module(name=None):
    return importlib.import_module(PKG if name is None else PKG + "." + name)
