"""Notebook loss / model variants (SURVEY 8 a-bis) on the CUDA kernels against the fp64 restatement
(oracle/variants_port.py): values, gradients with respect to U, gradients with respect to network parameters, and a
short training run.  Tolerances: the kernels are fp32 with fp64-accumulated Gram matrices, the oracle is fp64."""
import os
import sys

import numpy as np
import pytest
import scipy.linalg
import torch

from gpu_util import pkg, dev, dropin, bunny_levels

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import variants_port as vp          # noqa: E402

pytestmark = pytest.mark.gpu


def _setup(k, seed=0, coarse=False):
    fem, (K, M), (Kc, Mc) = bunny_levels()
    if coarse:
        K, M, verts = Kc, Mc, fem["coarse_verts"]
    else:
        verts = fem["verts"]
    rng = np.random.default_rng(seed)
    n = K.shape[0]
    U = rng.standard_normal((n, k)) / np.sqrt(n) * 30.0
    return verts, K.tocsr(), M.tocsr(), U.astype(np.float32)


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("k", [8, 16, 50])
def test_dense_rayleigh_loss_value_and_gradient(k):
    v = pkg("variants")
    _, K, M, U0 = _setup(k)
    pair = pkg("sparse").OperatorPair(K, M, dev())
    U = torch.from_numpy(U0).to(dev()).requires_grad_(True)
    loss, l1, dl, ol, lam = v.dense_rayleigh_loss(U, pair)
    loss.backward()
    Ur = torch.from_numpy(U0.astype(np.float64)).requires_grad_(True)
    Kt = vp.to_torch_sparse(K.astype(np.float32).astype(np.float64))
    Mt = vp.to_torch_sparse(M.astype(np.float32).astype(np.float64))
    rloss, rl1, rdl, rol, rlam = vp.dense_rayleigh_loss(Ur, Kt, Mt)
    rloss.backward()
    for a, b in ((loss, rloss), (l1, rl1), (dl, rdl), (ol, rol)):
        assert abs(a.item() - b.item()) <= 2e-5 * abs(b.item()) + 1e-12
    assert _rel(lam.detach().cpu(), rlam.detach()) < 2e-5
    assert _rel(U.grad.cpu(), Ur.grad) < 5e-5


@pytest.mark.parametrize("k", [8, 32])
def test_whitened_subspace_loss_value_and_gradient(k):
    v = pkg("variants")
    _, K, M, U0 = _setup(k, seed=1)
    Kn, Mn, ks, ms = v.frobenius_normalised(K, M)
    rKn, rMn, rks, rms = vp.frobenius_normalised(K, M)
    assert abs(ks - rks) < 1e-9 * rks and abs(ms - rms) < 1e-9 * rms
    pair = pkg("sparse").OperatorPair(Kn, Mn, dev())
    U = torch.from_numpy(U0).to(dev()).requires_grad_(True)
    loss, terms, eigs = v.whitened_subspace_loss(U, pair, lambda_orth=0.1)
    loss.backward()
    Ur = torch.from_numpy(U0.astype(np.float64)).requires_grad_(True)
    Kt = vp.to_torch_sparse(sp_f32(Kn))
    Mt = vp.to_torch_sparse(sp_f32(Mn))
    rloss, rterms, reigs = vp.whitened_subspace_loss(Ur, Kt, Mt, lambda_orth=0.1)
    rloss.backward()
    assert abs(loss.item() - rloss.item()) <= 1e-5 * abs(rloss.item())
    hinge_atol = 1e-5 * reigs.abs().max().item()     # the gap / ordering hinges act on DIFFERENCES of the estimates
    for name in ("zero", "trace", "diversity", "offdiag", "ordering", "stability"):
        atol = hinge_atol if name in ("diversity", "ordering") else 1e-12
        assert abs(terms[name].item() - rterms[name].item()) <= 1e-5 * abs(rterms[name].item()) + atol, name
    assert terms["orth"].item() < 1e-12 and rterms["orth"].item() < 1e-12       # rounding only: W B W = I by construction
    assert _rel(eigs.detach().cpu(), reigs.detach()) < 1e-5
    assert _rel(U.grad.cpu(), Ur.grad) < 1e-4


def sp_f32(A):
    return A.astype(np.float32).astype(np.float64)


def test_single_mode_loss_and_network_gradients():
    v = pkg("variants")
    verts, K, M, _ = _setup(1, coarse=True)
    pair = pkg("sparse").OperatorPair(K, M, dev())
    torch.manual_seed(3)
    ref = vp.EigenfunctionNN(32, 3, initial_eigenvalue=0.7).double()
    net = v.EigenfunctionNN(32, 3, initial_eigenvalue=0.7).to(dev())
    net.load_state_dict({k_: t.float() for k_, t in ref.state_dict().items()})
    X = torch.from_numpy(verts.astype(np.float32))
    prev = [torch.from_numpy(np.random.default_rng(5).standard_normal(K.shape[0]).astype(np.float32)) for _ in range(2)]
    u, lam = net(X.to(dev()))
    total, eig, norm, ortho = v.single_mode_loss(u, lam, pair, [p.to(dev()) for p in prev], ortho_weight=2.0)
    total.backward()
    ru, rlam = ref(X.double())
    Kt, Mt = vp.to_torch_sparse(sp_f32(K)), vp.to_torch_sparse(sp_f32(M))
    rtotal, reig, rnorm, rortho = vp.single_mode_loss(ru, rlam.reshape(()), Kt, Mt, [p.double() for p in prev], ortho_weight=2.0)
    rtotal.backward()
    assert _rel(u.detach().cpu(), ru.detach()) < 1e-5
    for a, b in ((total, rtotal), (eig, reig), (norm, rnorm), (ortho, rortho)):
        assert abs(a.item() - b.item()) <= 1e-4 * abs(b.item()) + 1e-10
    for (name, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert _rel(p.grad.cpu(), q.grad) < 2e-4, name


def test_coordinate_mlp_trains_towards_the_low_modes():
    """The SiLU coordinate network of simplified_loss.ipynb with the whitened functional: a few hundred Adam steps on
    the coarse bunny drive the sorted Rayleigh estimates towards the exact eigenvalues of the normalised pencil, and
    forward / parameter gradients agree with the fp64 network at the start."""
    v = pkg("variants")
    verts, K, M, _ = _setup(1, coarse=True)
    k = 6
    Kn, Mn, ks, ms = v.frobenius_normalised(K, M)
    pair = pkg("sparse").OperatorPair(Kn, Mn, dev())
    torch.manual_seed(0)
    ref = vp.CoordinateMLP(3, k, (64, 64), "silu").double()
    net = v.CoordinateMLP(3, k, (64, 64), "silu").to(dev())
    net.load_state_dict({k_: t.float() for k_, t in ref.state_dict().items()})
    X = torch.from_numpy(verts.astype(np.float32)).to(dev())
    loss, _, _ = v.whitened_subspace_loss(net(X), pair)
    loss.backward()
    rloss, _, _ = vp.whitened_subspace_loss(ref(torch.from_numpy(verts)), vp.to_torch_sparse(sp_f32(Kn)), vp.to_torch_sparse(sp_f32(Mn)))
    rloss.backward()
    assert abs(loss.item() - rloss.item()) <= 1e-4 * abs(rloss.item())
    for (name, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert _rel(p.grad.cpu(), q.grad) < 2e-3, name
    exact = scipy.linalg.eigh(Kn.toarray(), Mn.toarray(), eigvals_only=True, subset_by_index=[0, k - 1])
    opt = torch.optim.Adam(net.parameters(), lr=3e-3)
    first = None
    for it in range(400):
        opt.zero_grad()
        loss, terms, eigs = v.whitened_subspace_loss(net(X), pair)
        loss.backward()
        opt.step()
        first = loss.item() if first is None else first
    assert loss.item() < 0.5 * first
    assert eigs.detach().sum().item() >= exact.sum() * (1 - 1e-4)      # Rayleigh-Ritz estimates bound from above


def test_adaptive_corrector_and_smoothness_terms():
    """AdaptiveCorrector (per-mode scales) + the two Laplacian-energy terms: forward, loss values and every parameter
    gradient (including mode_scales) against the fp64 restatement."""
    v = pkg("variants")
    cm = dropin("corrector_model")
    verts, K, M, _ = _setup(1, coarse=True)
    n, k, d = K.shape[0], 8, 5
    rng = np.random.default_rng(11)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ei = np.stack([np.repeat(np.arange(n), 6), rng.integers(0, n, 6 * n)]).astype(np.int64)
    U_base = (rng.standard_normal((n, k)) / np.sqrt(n)).astype(np.float32)
    torch.manual_seed(2)
    ref = vp.AdaptiveCorrector(d, k, (32, 16), init_scale=0.05).double()
    net = cm.AdaptiveCorrector(d, k, (32, 16), 0.0, init_scale=0.05).to(dev())
    net.load_state_dict({k_: t.float() for k_, t in ref.state_dict().items()})
    pair = pkg("sparse").OperatorPair(K, M, dev())
    corr = net(torch.from_numpy(x).to(dev()), torch.from_numpy(ei).to(dev()))
    U_pred = torch.from_numpy(U_base).to(dev()) + corr
    a, b = v.smoothness_loss(corr, U_pred, pair)
    (3.0 * (a + b)).backward()
    rcorr = ref(torch.from_numpy(x).double(), torch.from_numpy(ei))
    rU = torch.from_numpy(U_base).double() + rcorr
    ra, rb = vp.smoothness_loss(rcorr, rU, vp.to_torch_sparse(sp_f32(K)))
    (3.0 * (ra + rb)).backward()
    assert _rel(corr.detach().cpu(), rcorr.detach()) < 1e-5
    assert abs(a.item() - ra.item()) <= 1e-5 * abs(ra.item()) and abs(b.item() - rb.item()) <= 1e-5 * abs(rb.item())
    scale = max(q.grad.abs().max().item() for q in ref.parameters())
    for (name, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        # (the gradient of the last bias is zero in exact arithmetic - constants are in the null space of L - and
        #  rounding noise on both sides: hence the absolute part of the tolerance)
        err = (p.grad.cpu().double() - q.grad).abs().max().item()
        assert err <= 1e-4 * q.grad.abs().max().item() + 1e-6 * scale, name
