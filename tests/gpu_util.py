import importlib
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "eigen-pinns_b200", "src")


def pkg(name=None):
    return importlib.import_module("eigen-pinns_b200" + ("." + name if name else ""))


def dropin(name):
    """Import one of the flat, reference-shaped modules (src/ must be on sys.path like the reference)."""
    if SRC not in sys.path:
        sys.path.insert(0, SRC)
    return importlib.import_module(name)


def dev():
    return torch.device("cuda", 0)


def bunny_levels():
    from conftest import load_golden, csr_from_golden
    fem = load_golden("bunny_fem.npz")
    n = fem["verts"].shape[0]
    K, M = csr_from_golden(fem, "K", n), csr_from_golden(fem, "M", n)
    Kc, Mc = pkg("fem").assemble_stiffness_mass(fem["coarse_verts"], fem["coarse_tris"])
    return fem, (K, M), (Kc, Mc)


def random_csr(n_rows, n_cols, max_nnz_row, seed, empty_every=0):
    rng = np.random.default_rng(seed)
    rows, cols = [], []
    for r in range(n_rows):
        if empty_every and r % empty_every == 0:
            continue
        c = rng.choice(n_cols, size=rng.integers(1, max_nnz_row + 1), replace=False)
        rows += [r] * len(c)
        cols += list(c)
    vals = rng.standard_normal(len(rows))
    return sp.csr_matrix((vals, (rows, cols)), shape=(n_rows, n_cols))
