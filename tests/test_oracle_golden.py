"""Pin the CPU oracle (oracle/*.py) against fixtures written by the UNMODIFIED reference
(oracle/make_golden.py) and against the reference's printed known answers.  CPU only."""
import importlib

import numpy as np
import pytest
import torch

from conftest import load_golden, csr_from_golden
from oracle import samplers_port, step_port

fem_mod = importlib.import_module("eigen-pinns_b200.fem")


@pytest.fixture(scope="module")
def fem():
    return load_golden("bunny_fem.npz")


def _levels(fem):
    n, nc = fem["verts"].shape[0], fem["coarse_verts"].shape[0]
    K, M = csr_from_golden(fem, "K", n), csr_from_golden(fem, "M", n)
    Kc, Mc = fem_mod.assemble_stiffness_mass(fem["coarse_verts"], fem["coarse_tris"])
    return (K, M), (Kc, Mc)


def test_known_answer_bunny_eigenvalues(fem):
    # delta_pinns_validation/downsampling_toy_example.ipynb cell 14 (printed by the reference authors)
    printed = np.array([1.60038744e-01, 4.25258130e-01, 4.38250633e-01, 5.38463815e-01])
    np.testing.assert_allclose(fem["eig10"][1:5], printed, rtol=0, atol=5.1e-10)   # printed to 9 s.f.
    assert abs(fem["eig10"][0]) < 1e-10
    # iterative_eigenvalues_on_cloud.ipynb cell 17: first 10 to 3 d.p.
    np.testing.assert_allclose(np.round(fem["eig10"], 3),
                               [0.0, 0.160, 0.425, 0.438, 0.538, 0.612, 0.896, 1.274, 1.496, 1.643], atol=1e-12)


def test_sparse_fem_equals_reference_dense(fem):
    n = fem["verts"].shape[0]
    K, M = fem_mod.assemble_stiffness_mass(fem["verts"], fem["tris"])
    Kr, Mr = csr_from_golden(fem, "K", n), csr_from_golden(fem, "M", n)
    assert K.nnz == Kr.nnz == 17449
    assert np.array_equal(K.indptr, Kr.indptr) and np.array_equal(K.indices, Kr.indices)
    assert np.array_equal(M.indices, K.indices)
    np.testing.assert_allclose(K.data, Kr.data, rtol=0, atol=5e-13 * np.abs(Kr.data).max())
    np.testing.assert_allclose(M.data, Mr.data, rtol=0, atol=5e-15)
    Kc, Mc = fem_mod.assemble_stiffness_mass(fem["coarse_verts"], fem["coarse_tris"])
    np.testing.assert_allclose(np.asarray(abs(Kc).sum(1)).ravel(), fem["Kc_rowsum_abs"], rtol=1e-11)
    np.testing.assert_allclose(np.asarray(Mc.sum(1)).ravel(), fem["Mc_rowsum"], rtol=1e-12)


def test_sparse_fem_eigenpairs_match_reference(fem):
    from scipy.sparse.linalg import eigsh
    K, M = fem_mod.assemble_stiffness_mass(fem["verts"], fem["tris"])
    w, V = eigsh(K, k=10, M=M, sigma=-0.01, which="LM")
    np.testing.assert_allclose(w[1:], fem["eig10"][1:], rtol=1e-8)
    ref = fem["evec10"]
    for j in range(1, 10):
        s = np.sign(V[:, j] @ (M @ ref[:, j]))
        assert np.abs(s * V[:, j] - ref[:, j]).max() < 1e-6


@pytest.mark.parametrize("tag,k", [("k16_1lvl", 16), ("k64_1lvl", 64), ("k16_2lvl", 16)])
def test_eigen_loss_port(fem, tag, k):
    g = load_golden("eigen_loss.npz")
    (K, M), (Kc, Mc) = _levels(fem)
    levels = [(Kc, Mc), (K, M)] if tag.endswith("2lvl") else [(K, M)]
    U = torch.from_numpy(g[f"{tag}_U"]).requires_grad_(True)
    Ks, Ms = [a for a, _ in levels], [b for _, b in levels]
    l_res, l_orth, lams = step_port.residual_ortho_loss(U, Ks, Ms, g[f"{tag}_offsets"], 1000.0, 10.0, k)
    extra = step_port.eigenvalue_losses(lams[0], torch.from_numpy(g[f"{tag}_lam_target"]), 0.0, 0.5, 2.0, 3.0)
    total = l_res + l_orth + sum(extra)
    total.backward()
    # the coarse level here is re-assembled by the sparse path (1e-13 away from the dense reference)
    rtol = 1e-6
    assert l_res.item() == pytest.approx(float(g[f"{tag}_loss_res"]), rel=rtol)
    assert l_orth.item() == pytest.approx(float(g[f"{tag}_loss_orth"]), rel=rtol)
    assert total.item() == pytest.approx(float(g[f"{tag}_total"]), rel=rtol)
    np.testing.assert_allclose([e.item() for e in extra], g[f"{tag}_extra"], rtol=1e-5, atol=1e-7)
    for i, l in enumerate(lams):
        np.testing.assert_allclose(l.detach().numpy(), g[f"{tag}_lam{i}"], rtol=2e-5, atol=1e-6)
    gref = g[f"{tag}_grad"]
    assert np.abs(U.grad.numpy() - gref).max() <= 2e-5 * np.abs(gref).max()


@pytest.mark.parametrize("model_type", ["simple", "spectral"])
def test_corrector_and_training_port(fem, model_type):
    g = load_golden("corrector_train.npz")
    (K, M), (Kc, Mc) = _levels(fem)
    k, hidden = 16, list(g["hidden"])
    t = model_type
    # trainer set-up helpers
    U0 = [torch.from_numpy(g["U0_0"]), torch.from_numpy(g["U0_1"])]
    U_norm = [step_port.m_normalize(U0[0], Mc), step_port.m_normalize(U0[1], M)]
    np.testing.assert_allclose(U_norm[0].numpy(), g["U_norm_0"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(U_norm[1].numpy(), g["U_norm_1"], rtol=1e-5, atol=1e-7)
    vals, U_rr = step_port.rayleigh_ritz(g["U0_1"], K, M)
    np.testing.assert_allclose(vals, g["rr_vals"], rtol=1e-4, atol=1e-5)
    ei = [torch.from_numpy(g["edge_index_0"]), torch.from_numpy(g["edge_index_1"])]
    lam = [torch.from_numpy(g["lam_0"]), torch.from_numpy(g["lam_1"])]
    X = [fem["coarse_verts"], fem["verts"]]
    feats = torch.cat([step_port.level_features(X[i], torch.from_numpy(g[f"U_norm_{i}"]), lam[i], ei[i],
                                                (Kc, K)[i], (Mc, M)[i], i, 2) for i in range(2)], dim=0)
    np.testing.assert_allclose(feats.numpy(), g[f"{t}_x_feats"], rtol=2e-4, atol=2e-5)
    # forward + six reference epochs
    x = torch.from_numpy(g[f"{t}_x_feats"])
    ei_all = torch.from_numpy(g["edge_index_all"])
    A_norm = step_port.gcn_norm_adjacency(ei_all, x.shape[0]) if t == "spectral" else None
    tr = step_port.CorrectorTrainer(x, ei_all, torch.cat([torch.from_numpy(g["U_norm_0"]),
                                                          torch.from_numpy(g["U_norm_1"])]),
                                    [Kc, K], [Mc, M], lam[0], hidden, k, model_type=t, A_norm=A_norm)
    n_lin = len(hidden) + 1
    tr.load_parameters([g[f"{t}_init_net.{2 * i}.weight"] for i in range(n_lin)],
                       [g[f"{t}_init_net.{2 * i}.bias"] for i in range(n_lin)])
    out0 = tr.forward().detach().numpy()
    np.testing.assert_allclose(out0, g[f"{t}_out0"], rtol=1e-4, atol=1e-6)
    tr.epoch = 2500
    hist = np.array([tr.step()[:3] for _ in range(6)])
    np.testing.assert_allclose(hist, g[f"{t}_losses"], rtol=2e-4)
    for i in range(n_lin):
        np.testing.assert_allclose(tr.weights[i].detach().numpy(), g[f"{t}_after_net.{2 * i}.weight"],
                                   rtol=0, atol=2e-5)


FPS_CASES = ["bunny", "ico8", "cloud", "grid"]
VOX_CASES = ["bunny", "cloud", "grid", "ico8"]


@pytest.mark.parametrize("tag", FPS_CASES)
def test_fps_port_bit_exact(tag):
    g = load_golden("samplers.npz")
    hier = [int(h) for h in g[f"fps_{tag}_hier"]]
    out = samplers_port.fps_levels(g[f"fps_{tag}_pts"], hier, int(g[f"fps_{tag}_start"]))
    assert sorted(out.keys()) == list(range(len(hier) + 1))
    for lv, idx in out.items():
        assert np.array_equal(idx, g[f"fps_{tag}_lv{lv}"]), (tag, lv)


@pytest.mark.parametrize("tag", VOX_CASES)
def test_voxel_port_bit_exact(tag):
    g = load_golden("samplers.npz")
    hier = [int(h) for h in g[f"vox_{tag}_hier"]]
    out = samplers_port.voxel_levels(g[f"vox_{tag}_pts"], hier)
    for lv, idx in out.items():
        assert np.array_equal(idx, g[f"vox_{tag}_lv{lv}"]), (tag, lv)


def test_voxel_known_answer_level_sizes():
    # multigrid_gnn_multires_voxel_downsampling.ipynb cell 1 prints 256 / 512 / 1003 / 2503
    g = load_golden("samplers.npz")
    assert [g[f"vox_bunny_lv{i}"].size for i in range(4)] == [256, 512, 1003, 2503]


def test_fps_q2_bare_array():
    pts = np.random.default_rng(0).standard_normal((10, 3))
    out = samplers_port.fps_levels(pts, [4, 20], 0)
    assert isinstance(out, np.ndarray) and np.array_equal(out, np.arange(10))
