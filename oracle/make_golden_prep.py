"""Second fixture script: pre-processing helpers and the coarse-grid correction of the UNMODIFIED reference
(utils.build_prolongation / build_knn_graph / jacobi_smooth / orthonormalize, MultigridGNN.apply_coarse_grid_correction)
on the bunny / coarse_3 pair.  ORACLE - test infrastructure only; run where /root/reference exists.

    python oracle/make_golden_prep.py      ->  tests/golden/prep_cgc.npz
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import reference_loader  # noqa: E402
from oracle.make_golden import make_config  # noqa: E402


def main():
    ref = reference_loader.load()
    from scipy.linalg import eigh
    from scipy.sparse import coo_matrix
    g = np.load(os.path.join(ROOT, "tests", "golden", "bunny_fem.npz"))
    fine = ref.Mesh.Mesh(verts=g["verts"], connectivity=g["tris"])
    coarse = ref.Mesh.Mesh(verts=g["coarse_verts"], connectivity=g["coarse_tris"])
    Kf_d, Mf_d = fine.computeLaplacian()
    Kc_d, Mc_d = coarse.computeLaplacian()
    K_f, M_f, K_c, M_c = (coo_matrix(a) for a in (Kf_d, Mf_d, Kc_d, Mc_d))
    k = 12
    _, U0 = eigh(Kc_d, Mc_d, subset_by_index=[0, k - 1])
    P = ref.utils.build_prolongation(coarse.verts, fine.verts, k=8).tocoo()
    U1 = ref.utils.jacobi_smooth(M_f, K_f, P @ U0, alpha=0.1, n_iters=10)
    knn = ref.utils.build_knn_graph(coarse.verts, k=5).numpy()
    rng = np.random.default_rng(3)
    A = rng.standard_normal((60, 5))
    Ms = np.diag(rng.uniform(0.5, 2.0, 60))
    ortho = ref.utils.orthonormalize(A, Ms)
    ncol, nrm = ref.utils.normalize_columns_np(A)
    # coarse-grid correction with a REGULAR coarse operator (K_c + 0.1 M_c): the plain K_c of these meshes is singular
    cfg = make_config(ref, n_modes=k)
    gnn = ref.multigrid_model.MultigridGNN(cfg)
    gnn.device = torch.device("cpu")
    K_reg = coo_matrix(Kc_d + 0.1 * Mc_d)
    with contextlib.redirect_stdout(io.StringIO()):
        U_cgc, lam_f = gnn.apply_coarse_grid_correction(torch.FloatTensor(U1), K_f, M_f, K_reg, P)
    out = dict(k=k, U0=U0, P_row=P.row, P_col=P.col, P_data=P.data, U1=U1, knn_coarse=knn, ortho_in=A,
               ortho_M=np.diag(Ms), ortho_out=ortho, ncol=ncol, nrm=nrm, U_cgc=U_cgc.numpy(), lam_f=lam_f.numpy())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "prep_cgc.npz"), **out)
    print({k_: np.asarray(v).shape for k_, v in out.items()})


if __name__ == "__main__":
    main()
