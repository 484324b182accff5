"""GPU parity: corrector MLP (fp32 path), drop-in modules, optimiser and a short training run,
against the golden fixtures written by the reference and the CPU oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import pkg, dev, dropin, bunny_levels
from oracle import step_port

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,d_in,d_out,relu", [(1, 7, 5, True), (300, 82, 256, True), (1000, 256, 32, False),
                                               (4097, 146, 64, True), (129, 50, 16, False)])
def test_linear_forward_backward_vs_torch(n, d_in, d_out, relu):
    ops = pkg("ops")
    g = torch.Generator().manual_seed(n)
    X = torch.randn(n, d_in, generator=g).abs() - 0.3          # mixed signs (ReLU mask on the input)
    W = torch.randn(d_out, d_in, generator=g) / np.sqrt(d_in)
    b = torch.randn(d_out, generator=g)
    dY = torch.randn(n, d_out, generator=g)
    Y = ops.linear_fwd(X.to(dev()), W.to(dev()), b.to(dev()), relu)
    ref = torch.nn.functional.linear(X.double(), W.double(), b.double())
    ref = torch.relu(ref) if relu else ref
    np.testing.assert_allclose(Y.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-5)
    dX, dW, db = ops.linear_bwd(X.to(dev()), W.to(dev()), dY.to(dev()), True, relu_mask=True)
    dX_ref = (dY.double() @ W.double()) * (X > 0).double()
    dW_ref = dY.double().t() @ X.double()
    np.testing.assert_allclose(dX.cpu().numpy(), dX_ref.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dW.cpu().numpy(), dW_ref.numpy(), rtol=1e-5, atol=1e-5 * np.sqrt(n))
    np.testing.assert_allclose(db.cpu().numpy(), dY.double().sum(0).numpy(), rtol=1e-5, atol=1e-5 * np.sqrt(n))


@pytest.mark.parametrize("model_type", ["simple", "spectral"])
def test_corrector_forward_matches_reference(model_type):
    cm = dropin("corrector_model")
    utils = dropin("utils")
    g = load_golden("corrector_train.npz")
    t = model_type
    x = torch.from_numpy(g[f"{t}_x_feats"]).to(dev())
    hidden = [int(h) for h in g["hidden"]]
    cls = cm.SimpleCorrector if t == "simple" else cm.SpectralCorrector
    model = cls(x.shape[1], 16, hidden, 0.0).to(dev())
    state = {k_[len(t) + 6:]: torch.from_numpy(g[k_]) for k_ in g.files if k_.startswith(f"{t}_init_")}
    model.load_state_dict(state)
    ei = torch.from_numpy(g["edge_index_all"])
    if t == "simple":
        out = model(x, ei.to(dev()))
    else:
        A_norm = utils.build_A_norm(ei, x.shape[0], dev())
        A = A_norm.coalesce()
        np.testing.assert_array_equal(A.indices().cpu().numpy(), g["A_norm_indices"])
        np.testing.assert_allclose(A.values().cpu().numpy(), g["A_norm_values"], rtol=1e-6)
        out = model(x, A_norm)
    np.testing.assert_allclose(out.detach().cpu().numpy(), g[f"{t}_out0"], rtol=1e-4, atol=2e-6)
    # autograd through the layer kernels
    out.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


def _golden_trainer(model_type, mlp_mode="fp32"):
    mg = dropin("multigrid_model")
    cfgm = dropin("config")
    import os
    from gpu_util import SRC
    g = load_golden("corrector_train.npz")
    fem, (K, M), (Kc, Mc) = bunny_levels()
    cfg = cfgm.PINNConfig.from_yaml(os.path.join(SRC, "parameters.yml"))
    cfg.n_modes, cfg.hidden_layers, cfg.model_type, cfg.mlp_mode = 16, [int(h) for h in g["hidden"]], model_type, mlp_mode
    gnn = mg.MultigridGNN(cfg)
    return gnn, g, fem, (K, M), (Kc, Mc)


def test_trainer_setup_methods_match_reference():
    gnn, g, fem, (K, M), (Kc, Mc) = _golden_trainer("simple")
    U_norm = gnn._normalize_eigenvectors([g["U0_0"], g["U0_1"]], [Mc, M])
    np.testing.assert_allclose(U_norm[0].cpu().numpy(), g["U_norm_0"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(U_norm[1].cpu().numpy(), g["U_norm_1"], rtol=1e-5, atol=1e-7)
    vals, U_rr = gnn.refine_eigenvectors(g["U0_1"], K, M)
    np.testing.assert_allclose(vals, g["rr_vals"], rtol=1e-4, atol=2e-5)
    assert np.abs(np.abs(U_rr) - np.abs(g["rr_U"])).max() < 5e-3
    ei = [torch.from_numpy(g["edge_index_0"]), torch.from_numpy(g["edge_index_1"])]
    lam = [torch.from_numpy(g["lam_0"]), torch.from_numpy(g["lam_1"])]
    X = [fem["coarse_verts"], fem["verts"]]
    x_feats, edge_all, A_norm = gnn._build_features(X, [torch.from_numpy(g["U_norm_0"]), torch.from_numpy(g["U_norm_1"])],
                                                    lam, ei, [Kc, K], [Mc, M])
    np.testing.assert_allclose(x_feats.cpu().numpy(), g["simple_x_feats"], rtol=2e-4, atol=2e-5)
    np.testing.assert_array_equal(edge_all.cpu().numpy(), g["edge_index_all"])
    assert gnn._compute_node_offsets(X) == [0, X[0].shape[0]]


@pytest.mark.parametrize("model_type", ["simple", "spectral"])
def test_six_training_epochs_follow_the_reference(model_type):
    """Loss trajectory of six reference epochs (epochs 2500..2505 of the scale ramp) and the weights
    after them: fp32 engine vs the reference's own CPU run (fixture), tolerance 2e-4 on losses."""
    gnn, g, fem, (K, M), (Kc, Mc) = _golden_trainer(model_type)
    t = model_type
    x = torch.from_numpy(g[f"{t}_x_feats"]).to(dev())
    ei = torch.from_numpy(g["edge_index_all"])
    utils = dropin("utils")
    A_norm = utils.build_A_norm(ei, x.shape[0], dev()) if t == "spectral" else None
    gnn._initialize_model(x.shape[1], 16, gnn.hidden_layers, 0.0)
    state = {k_[len(t) + 6:]: torch.from_numpy(g[k_]) for k_ in g.files if k_.startswith(f"{t}_init_")}
    gnn.model.load_state_dict(state)
    opt, sched = gnn._create_optimizer(gnn.lr, gnn.weight_decay)
    U_all = torch.cat([torch.from_numpy(g["U_norm_0"]), torch.from_numpy(g["U_norm_1"])])
    offs = [0, g["U_norm_0"].shape[0]]
    eng = gnn._make_engine(x, ei, A_norm, U_all, [Kc, K], [Mc, M], torch.from_numpy(g["lam_0"]), offs, opt)
    hist = []
    for epoch in range(2500, 2506):
        acc = eng.step(epoch, lr=opt.param_groups[0]["lr"]).cpu().numpy()
        hist.append([acc[5], acc[0], acc[1]])
    np.testing.assert_allclose(np.array(hist), g[f"{t}_losses"], rtol=2e-4)
    sd = gnn.model.state_dict()                      # module parameters are views of the flat buffer
    for name in sd:
        if name.endswith("weight"):
            np.testing.assert_allclose(sd[name].cpu().numpy(), g[f"{t}_after_{name}"], rtol=0, atol=3e-5)
    U_final = gnn._generate_final_predictions(x, ei, A_norm, U_all, [torch.from_numpy(g["U_norm_0"]),
                                                                      torch.from_numpy(g["U_norm_1"])], [Mc, M])
    np.testing.assert_allclose(U_final, g[f"{t}_U_final"], rtol=0, atol=5e-4)


def test_dropin_loss_methods_autograd():
    gnn, g, fem, (K, M), (Kc, Mc) = _golden_trainer("simple")
    ge = load_golden("eigen_loss.npz")
    U = torch.from_numpy(ge["k16_2lvl_U"]).to(dev()).requires_grad_(True)
    l_res, l_orth, lams = gnn._compute_residual_ortho_loss(U, [Kc, K], [Mc, M], [0, Kc.shape[0]], 1000.0, 10.0, 16)
    extra = gnn._compute_eigenvalue_losses([Mc, M], torch.from_numpy(ge["k16_2lvl_lam_target"]).to(dev()), lams,
                                           0.0, 0.5, 2.0, 3.0)
    total = l_res + l_orth + sum(extra)
    total.backward()
    assert total.item() == pytest.approx(float(ge["k16_2lvl_total"]), rel=1e-5)
    np.testing.assert_allclose([e.item() for e in extra], ge["k16_2lvl_extra"], rtol=1e-5, atol=1e-7)
    gref = ge["k16_2lvl_grad"]
    assert np.abs(U.grad.cpu().numpy() - gref).max() <= 2e-5 * np.abs(gref).max()


def test_adam_clip_kernel_vs_torch():
    ops = pkg("ops")
    torch.manual_seed(3)
    p0 = torch.randn(10007)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3, weight_decay=1e-5)
    p = p0.clone().to(dev())
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    sq = torch.zeros(1, dtype=torch.float64, device=dev())
    for step in range(1, 6):
        gr = torch.randn(10007) * (30.0 if step % 2 else 0.01)     # clipped and unclipped steps
        ref.grad = gr.clone()
        torch.nn.utils.clip_grad_norm_([ref], 10.0)
        opt.step()
        gd = gr.to(dev())
        ops.grad_sqnorm(gd, sq)
        ops.adam_clip_step(p, gd, m, v, 1e-3, 0.9, 0.999, 1e-8, 1e-5, step, 10.0, sq)
        np.testing.assert_allclose(p.cpu().numpy(), ref.detach().numpy(), rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("mlp_mode", ["fp32", "bf16"])
def test_cuda_graph_replay_equals_eager(mlp_mode):
    """The captured step must reproduce the eager step BIT FOR BIT: every per-step scalar (scale ramp, learning rate,
    Adam step count) reaches the kernels with the same value either way, and the Adam bias corrections are derived
    from the integer step count inside the kernel in both cases."""
    runs = []
    for graphed in (False, True):
        gnn, g, fem, (K, M), (Kc, Mc) = _golden_trainer("simple", mlp_mode)
        x = torch.from_numpy(g["simple_x_feats"]).to(dev())
        ei = torch.from_numpy(g["edge_index_all"])
        gnn._initialize_model(x.shape[1], 16, gnn.hidden_layers, 0.0)
        gnn.model.load_state_dict({k_[12:]: torch.from_numpy(g[k_]) for k_ in g.files if k_.startswith("simple_init_")})
        opt, _ = gnn._create_optimizer(gnn.lr, gnn.weight_decay)
        U_all = torch.cat([torch.from_numpy(g["U_norm_0"]), torch.from_numpy(g["U_norm_1"])])
        eng = gnn._make_engine(x, ei, None, U_all, [Kc, K], [Mc, M], torch.from_numpy(g["lam_0"]),
                               [0, g["U_norm_0"].shape[0]], opt)
        hist = [eng.step(2500).cpu().numpy().copy()]
        if graphed:
            eng.enable_graph()
        for e in range(2501, 2512):
            lr = 1e-3 if e < 2506 else 5e-4                       # the schedule may change the rate between replays
            hist.append(eng.step(e, lr=lr).cpu().numpy().copy())
        runs.append((np.array(hist), eng.params.flat.clone(), eng.params.m.clone(), eng.params.v.clone()))
    assert np.array_equal(runs[0][0], runs[1][0]), np.abs(runs[0][0] - runs[1][0]).max()
    for a, b in zip(runs[0][1:], runs[1][1:]):
        assert torch.equal(a, b), (a - b).abs().max().item()


def test_coarse_grid_correction_matches_reference():
    """apply_coarse_grid_correction (reference multigrid_model.py:410-450) on the bunny / coarse pair with a regular
    coarse operator K_c + 0.1 M_c (cond 1.6e3; the plain K_c of these meshes is singular)."""
    import scipy.sparse as sp
    gnn, _, fem, (K, M), (Kc, Mc) = _golden_trainer("simple")
    g = load_golden("prep_cgc.npz")
    P = sp.coo_matrix((g["P_data"], (g["P_row"], g["P_col"])), shape=(K.shape[0], Kc.shape[0]))
    K_reg = sp.coo_matrix(Kc + 0.1 * Mc)
    U_cgc, lam = gnn.apply_coarse_grid_correction(torch.from_numpy(g["U1"]).float(), K, M, K_reg, P)
    np.testing.assert_allclose(lam.cpu().numpy(), g["lam_f"], rtol=2e-4, atol=2e-5)
    ref = g["U_cgc"]
    assert np.abs(U_cgc.cpu().numpy() - ref).max() <= 2e-3 * np.abs(ref).max()


def test_regularized_cg_coarse_solve_matches_reference_dense_solve():
    """cgc_mode 'regularized' (K_c + shift M_c solved by block CG on the device, nothing densified) against the
    reference's dense solve of the same regular operator (fixture: K_c + 0.1 M_c passed to the reference CGC)."""
    import scipy.sparse as sp
    gnn, _, fem, (K, M), (Kc, Mc) = _golden_trainer("simple")
    gnn.cgc_mode, gnn.cgc_shift = "regularized", 0.1
    g = load_golden("prep_cgc.npz")
    P = sp.coo_matrix((g["P_data"], (g["P_row"], g["P_col"])), shape=(K.shape[0], Kc.shape[0]))
    U_cgc, lam = gnn.apply_coarse_grid_correction(torch.from_numpy(g["U1"]).float(), K, M, sp.coo_matrix(Kc), P,
                                                  M_coarse=sp.coo_matrix(Mc))
    np.testing.assert_allclose(lam.cpu().numpy(), g["lam_f"], rtol=2e-4, atol=2e-5)
    ref = g["U_cgc"]
    assert np.abs(U_cgc.cpu().numpy() - ref).max() <= 2e-3 * np.abs(ref).max()
    assert 8 <= gnn.cgc_iterations < 2000
    # the singular plain K_c is handled too: the shift keeps the operator positive definite
    gnn.cgc_shift = 1e-3
    U2, _ = gnn.apply_coarse_grid_correction(torch.from_numpy(g["U1"]).float(), K, M, sp.coo_matrix(Kc), P,
                                             M_coarse=sp.coo_matrix(Mc))
    assert torch.isfinite(U2).all()


def test_sum_aggregation_and_offset_edges_variants():
    """Notebook variants of the aggregation: sum instead of mean, and per-level edge lists with node
    offsets; against torch index_add_ on the CPU."""
    cm, utils = dropin("corrector_model"), dropin("utils")
    g = torch.Generator().manual_seed(2)
    n0, n1, d = 300, 500, 7
    x = torch.randn(n0 + n1, d, generator=g)
    e0 = torch.randint(0, n0, (2, 2000), generator=g)
    e1 = torch.randint(0, n1, (2, 3000), generator=g)
    ei = utils.offset_edge_lists([e0, e1], [n0, n1])
    assert ei.shape == (2, 5000) and int(ei[:, 2000:].min()) >= n0
    model = cm.SimpleCorrector(d, 4, [16], 0.0, aggregation="sum").to(dev())
    h = model.corrector_input(x.to(dev()), ei.to(dev())).cpu()
    agg = torch.zeros_like(x)
    agg.index_add_(0, ei[0], x[ei[1]])
    np.testing.assert_allclose(h[:, :d].numpy(), x.numpy(), rtol=0, atol=0)
    np.testing.assert_allclose(h[:, d:].numpy(), agg.numpy(), rtol=1e-5, atol=1e-5)
    assert model(x.to(dev()), ei.to(dev())).shape == (n0 + n1, 4)
    with pytest.raises(ValueError):
        cm.SimpleCorrector(d, 4, [16], 0.0, aggregation="max")
