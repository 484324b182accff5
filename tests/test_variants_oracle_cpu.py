"""The restated notebook functionals (oracle/variants_port.py) behave as the notebooks intend on exact eigenpairs:
the checker used by tests/test_gpu_variants.py is itself sane.  CPU only."""
import os
import sys

import numpy as np
import scipy.linalg
import torch

from conftest import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import variants_port as vp          # noqa: E402
import importlib                                  # noqa: E402

fem_host = importlib.import_module("eigen-pinns_b200.fem")


def _coarse():
    g = load_golden("bunny_fem.npz")
    K, M = fem_host.assemble_stiffness_mass(g["coarse_verts"], g["coarse_tris"])
    return g["coarse_verts"], K.tocsr(), M.tocsr()


def test_functionals_vanish_on_exact_eigenpairs():
    verts, K, M = _coarse()
    k = 8
    vals, vecs = scipy.linalg.eigh(K.toarray(), M.toarray(), subset_by_index=[0, k - 1])
    U = torch.from_numpy(vecs)                                    # M-orthonormal, K U = M U diag(vals)
    Kt, Mt = vp.to_torch_sparse(K), vp.to_torch_sparse(M)
    loss, loss_1, diag_loss, off, lam = vp.dense_rayleigh_loss(U, Kt, Mt)
    assert loss_1.item() < 1e-8 and diag_loss.item() < 1e-12 and off.item() < 1e-12
    np.testing.assert_allclose(lam.numpy(), vals, rtol=1e-5, atol=1e-9)     # (the mesh has several zero modes)

    Kn, Mn, ks, ms = vp.frobenius_normalised(K, M)
    Q = torch.from_numpy(np.random.default_rng(0).standard_normal((k, k)))
    _, terms, eigs = vp.whitened_subspace_loss(U @ Q, vp.to_torch_sparse(Kn), vp.to_torch_sparse(Mn))
    assert terms["orth"].item() < 1e-16 and terms["ordering"].item() == 0.0
    # the whitening removes the mixing Q up to a rotation: the TRACE of the whitened Rayleigh matrix is the sum of the
    # exact eigenvalues of the normalised, shifted pencil ((K + 1e-4 I) / ks, M / ms) - up to the effect of the shift
    # 1e-4 I (not 1e-4 M) on the invariant subspace, 1e-7 relative here
    exact = scipy.linalg.eigh(Kn.toarray(), Mn.toarray(), eigvals_only=True, subset_by_index=[0, k - 1])
    np.testing.assert_allclose(eigs.sum().item(), exact.sum(), rtol=1e-6)
    assert eigs.sum().item() >= exact.sum() * (1 - 1e-12)          # Rayleigh-Ritz bounds from above

    u = U[:, 3:4].clone()
    total, eig, norm, ortho = vp.single_mode_loss(u, torch.tensor(vals[3]), Kt, Mt, previous=[U[:, 0], U[:, 1]])
    assert eig.item() < 1e-16 and norm.item() < 1e-16 and ortho.item() < 1e-16


def test_networks_follow_the_notebook_shapes():
    torch.manual_seed(0)
    net = vp.CoordinateMLP(3, 12, (32, 16), "silu").double()
    assert sorted(net.state_dict()) == ["net.0.bias", "net.0.weight", "net.2.bias", "net.2.weight", "net.4.bias", "net.4.weight"]
    x = torch.randn(7, 3, dtype=torch.float64)
    assert net(x).shape == (7, 12)
    one = vp.EigenfunctionNN(16, 3, initial_eigenvalue=-2.5).double()
    u, lam = one(x)
    assert u.shape == (7, 1) and lam.item() == 2.5                # lambda = |w|
    assert one.fc2.weight.shape == (16, 17)                        # the eigenvalue is appended to every layer's input


def test_operator_preparation_matches_the_restatement():
    """variants.frobenius_normalised (product, sparse) == the notebook's dense preparation restated in the oracle."""
    variants = importlib.import_module("eigen-pinns_b200.variants")
    _, K, M = _coarse()
    Kn, Mn, ks, ms = variants.frobenius_normalised(K, M, epsilon=1e-4)
    rKn, rMn, rks, rms = vp.frobenius_normalised(K, M, epsilon=1e-4)
    dense = K.toarray() + 1e-4 * np.eye(K.shape[0])
    assert abs(ks - np.linalg.norm(dense, "fro")) < 1e-9 * ks and abs(ks - rks) < 1e-12 * ks and abs(ms - rms) < 1e-12 * ms
    assert abs(Kn - rKn).max() < 1e-15 and abs(Mn - rMn).max() < 1e-15
    assert abs(np.sqrt((Kn.toarray() ** 2).sum()) - 1.0) < 1e-12
