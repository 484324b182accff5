"""GPU parity at the BENCHMARK sizes (BASELINE.json configs 4 and 5 shapes) against the CPU oracle:
  * icosphere, 998,562 vertices, k = 32, MLP 82 -> 256 x 6 -> 32   (fp32 parity mode and bf16 perf mode)
  * torus 1024 x 1024 = 1,048,576 vertices, k = 64, MLP 146 -> 256 x 6 -> 64
Thresholds: loss terms and eigenvalues within 1e-5 relative in fp32 mode; bf16 mode within the
tolerance stated in DESIGN.md (2.5e-2 per loss term and per eigenvalue relative to the largest, 5e-3 on the total at these sizes: measured 3e-3 on the k = 64 torus, 6e-4 on the icosphere).
Also the near-convergence case of the one-pass residual expansion."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import pkg, dev, bunny_levels
from oracle import step_port

pytestmark = pytest.mark.gpu
HIDDEN = [256] * 6


def _bench_problem(name):
    import bench
    w = bench.build_host_workload(name)
    k = w["k"]
    U_base = step_port.m_normalize(torch.from_numpy(w["U0"]), w["M"])
    ei = torch.from_numpy(w["edges"])
    lam = torch.linspace(0.0, 1.0, k)
    x = step_port.level_features(w["verts"], U_base, lam, ei, w["K"], w["M"], 0, 1)
    tr = step_port.CorrectorTrainer(x, ei, U_base, [w["K"]], [w["M"]], lam, HIDDEN, k)
    tr.epoch = 2500
    return w, x, ei, U_base, lam, tr


def _engine_for(w, x, ei, U_base, lam, tr, mlp_mode):
    ops, sparse, engine = pkg("ops"), pkg("sparse"), pkg("engine")
    n = x.shape[0]
    adj = sparse.CsrMatrix.from_edge_index(ei, n, dev())
    h = ops.neighbor_mean_concat(x.to(dev()), adj)
    params = engine.FlatParams([p.detach().clone() for p in tr.weights], [p.detach().clone() for p in tr.biases], dev())
    return engine.TrainStepEngine(h, U_base.to(dev()), [sparse.OperatorPair(w["K"], w["M"], dev())], [0], params,
                                  engine.StepConfig(), lam_target=lam.to(dev()), mlp_mode=mlp_mode)


def _compare(w, x, ei, U_base, lam, tr, modes):
    engines = {m: _engine_for(w, x, ei, U_base, lam, tr, m) for m in modes}     # before the oracle updates its weights
    total, l_res, l_orth, lams = tr.step()                                     # oracle: one full training step
    total2, l_res2, l_orth2, lams2, _ = tr.losses()                            # and the loss after it (forward only)
    for mode, eng in engines.items():
        a1 = eng.step(2500).cpu().numpy().copy()
        lam1 = eng.lams[0].cpu().numpy().copy()
        a2 = eng.step(2501).cpu().numpy().copy()
        lam2 = eng.lams[0].cpu().numpy().copy()
        if mode == "fp32":
            t_tot, t_term, t_lam = 1e-5, 1e-5, 1e-5
        else:
            # random-initialised corrector at scale 5: U_pred is noise-dominated, Rayleigh quotients are O(1e4-1e5) and
            # follow the bf16 rounding of the correction linearly (measured 1.7e-2 of the largest)
            t_tot, t_term, t_lam = 5e-3, 2.5e-2, 2.5e-2
        assert a1[5] == pytest.approx(total, rel=t_tot), (mode, a1, total)
        assert a1[0] == pytest.approx(l_res, rel=t_term) and a1[1] == pytest.approx(l_orth, rel=t_term), (mode, a1)
        ref = lams[0].numpy()
        assert np.abs(lam1 - ref).max() <= t_lam * np.abs(ref).max(), mode
        # second step: exercises backward + clip + Adam of the first (a wrong gradient moves the loss elsewhere)
        # (after ONE Adam step of a randomly initialised corrector the loss has moved by three orders of magnitude:
        # last-bit differences of the first step are amplified, hence the wider band)
        loose = 10.0 if mode == "fp32" else 5.0
        assert a2[5] == pytest.approx(float(total2), rel=loose * t_tot), (mode, a2, float(total2))
        ref2 = lams2[0].detach().numpy()
        assert np.abs(lam2 - ref2).max() <= loose * t_lam * np.abs(ref2).max(), mode
        del eng
        torch.cuda.empty_cache()


def test_icosphere_1m_k32_matches_oracle():
    w, x, ei, U_base, lam, tr = _bench_problem("icosphere1m")
    assert x.shape == (998562, 41)
    _compare(w, x, ei, U_base, lam, tr, ["fp32", "bf16"])


def test_torus_1m_k64_matches_oracle():
    w, x, ei, U_base, lam, tr = _bench_problem("torus1m")
    assert x.shape == (1048576, 73)
    _compare(w, x, ei, U_base, lam, tr, ["fp32", "bf16"])


def test_one_pass_residual_near_convergence():
    """U = exact generalised eigenvectors + 1e-4 of higher modes: the residual is ~1e-4 of |KU|, so the expansion
    sKK - 2 lam sKM + lam^2 sMM cancels ~8 digits.  fp64 accumulation keeps the loss within 1e-8 of an fp64
    evaluation of sum (KU - lam MU)^2 on the same KU, MU; against the exact-arithmetic value the only error left is
    the fp32 rounding of the SpMM outputs themselves (6e-8 of |KU| against a residual of 1e-4 |KU|: a few 1e-4
    relative on the squared residual; 1.7e-4 and 0.6e-4 were measured for two start vectors of eigsh)."""
    from scipy.sparse.linalg import eigsh
    ops, sparse = pkg("ops"), pkg("sparse")
    fem, (K, M), _ = bunny_levels()
    n, k = K.shape[0], 16
    vals, vecs = eigsh(K.tocsc(), k=k + 8, M=M.tocsc(), sigma=-0.01, which="LM", v0=np.ones(K.shape[0]))
    order = np.argsort(vals)
    vecs = vecs[:, order]
    rng = np.random.default_rng(3)
    U64 = vecs[:, :k] + 1e-4 * vecs[:, k:] @ rng.standard_normal((8, k))
    U = torch.from_numpy(U64.astype(np.float32)).to(dev())
    pair = sparse.OperatorPair(K, M, dev())
    KU, MU = ops.spmm2(pair, U)
    P = ops.eigen_partials(U, KU, MU)
    acc = torch.zeros(6, dtype=torch.float64, device=dev())
    lam, _ = ops.eigen_finalize(k, n, P, 1000.0, 10.0, acc)
    got_res = acc.cpu().numpy()[0]
    # (a) same fp32 KU / MU, everything else in fp64
    Ud, KUd, MUd = (t.double().cpu().numpy() for t in (U, KU, MU))
    lam64 = (Ud * KUd).sum(0) / ((Ud * MUd).sum(0) + 1e-12)
    ref_a = 1000.0 * ((KUd - MUd * lam64[None, :]) ** 2).mean()
    assert got_res == pytest.approx(ref_a, rel=1e-8)
    np.testing.assert_allclose(lam.cpu().numpy(), lam64, rtol=2e-7, atol=1e-9)
    # (b) exact arithmetic from the fp32 U and the fp32 operator values
    K32, M32 = K.astype(np.float32).astype(np.float64), M.astype(np.float32).astype(np.float64)
    KUx, MUx = K32 @ Ud, M32 @ Ud
    lamx = (Ud * KUx).sum(0) / ((Ud * MUx).sum(0) + 1e-12)
    ref_b = 1000.0 * ((KUx - MUx * lamx[None, :]) ** 2).mean()
    assert got_res == pytest.approx(ref_b, rel=1e-3)
    assert ref_b < 1e-6 * 1000.0 * (KUx ** 2).mean()          # the case really is near convergence
    # (c) the reference's own fp32 evaluation carries rounding noise of this order; ours is the more accurate one
    l_res, _, _ = step_port.residual_ortho_loss(U.cpu(), [K], [M], [0], 1000.0, 10.0, k)
    assert abs(float(l_res) - ref_b) <= 0.25 * ref_b
