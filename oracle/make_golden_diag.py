"""Third fixture script: the alignment metrics of the UNMODIFIED reference (src/diagnostics.py:12-115:
align_eigenvectors, get_subspace_error_and_alignment, compute_rayleigh_quotients) on the bunny FEM operators with a
perturbed, permuted and sign-flipped copy of the exact eigenvectors.  ORACLE - test infrastructure only; run where
/root/reference exists.

    python oracle/make_golden_diag.py      ->  tests/golden/diagnostics.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import reference_loader  # noqa: E402


def main():
    ref = reference_loader.load()
    import importlib
    sys.path.insert(0, reference_loader.REFERENCE_SRC)
    diag = importlib.import_module("diagnostics")
    sys.path.remove(reference_loader.REFERENCE_SRC)
    g = np.load(os.path.join(ROOT, "tests", "golden", "bunny_fem.npz"))
    mesh = ref.Mesh.Mesh(verts=g["verts"], connectivity=g["tris"])
    K, M = mesh.computeLaplacian()                       # dense, as the reference's diagnostics use them
    U_exact = g["evec10"]
    rng = np.random.default_rng(11)
    perm = rng.permutation(10)
    signs = rng.choice([-1.0, 1.0], size=10)
    U_pred = (U_exact + 0.05 * rng.standard_normal(U_exact.shape))[:, perm] * signs
    U_al, permutation, sg = diag.align_eigenvectors(U_pred, U_exact, M, verbose=False)
    U_pa, sub_err = diag.get_subspace_error_and_alignment(U_pred, U_exact, M)
    lam_pred = diag.compute_rayleigh_quotients(U_pred, K, M)
    lam_exact = diag.compute_rayleigh_quotients(U_exact, K, M)
    U_al_e, permutation_e, sg_e = diag.align_eigenvectors(U_pred, U_exact, None, verbose=False)
    out = os.path.join(ROOT, "tests", "golden", "diagnostics.npz")
    np.savez_compressed(out, U_pred=U_pred, U_aligned=U_al, permutation=permutation, signs=sg, U_procrustes=U_pa,
                        subspace_error=sub_err, lam_pred=lam_pred, lam_exact=lam_exact,
                        U_aligned_euclid=U_al_e, permutation_euclid=permutation_e, signs_euclid=sg_e)
    print("wrote", out, "subspace error", sub_err)


if __name__ == "__main__":
    main()
