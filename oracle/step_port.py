"""torch-CPU restatement of the eigen-loss training step.  ORACLE — test
infrastructure only (see oracle/__init__.py); never imported by the product.

Arithmetic is fp32 with the same torch CPU operators the reference uses, so the
port agrees with the reference to the last bit or two; tests/test_oracle_golden.py
pins it against fixtures written by the reference itself.

Reference anchors (all under /root/reference/src):
  coo_f32 ................ utils.py:14-20            (scipy fp64 -> torch COO fp32, coalesced)
  neighbour_mean_concat .. corrector_model.py:23-30  (SimpleCorrector aggregation)
  spectral_concat ........ corrector_model.py:76-79  (SpectralCorrector aggregation)
  mlp .................... corrector_model.py:12-21,31
  residual_ortho_loss .... multigrid_model.py:291-324
  eigenvalue_losses ...... multigrid_model.py:326-348
  m_normalize ............ multigrid_model.py:120-130
  rayleigh_ritz .......... multigrid_model.py:386-408
  level_features ......... multigrid_model.py:159-201
  gcn_norm_adjacency ..... utils.py:78-124
  CorrectorTrainer.step .. multigrid_model.py:237-261 (+ :218-224 optimiser/scheduler)
"""
import numpy as np
import torch


def coo_f32(A):
    A = A.tocoo()
    idx = torch.from_numpy(np.vstack((A.row, A.col)).astype(np.int64))
    val = torch.from_numpy(np.asarray(A.data)).to(torch.float32)
    return torch.sparse_coo_tensor(idx, val, A.shape).coalesce()


def neighbour_mean_concat(x, edge_index):
    dst, src = edge_index[0], edge_index[1]
    acc = torch.zeros_like(x)
    acc.index_add_(0, dst, x[src])
    cnt = torch.bincount(dst, minlength=x.shape[0]).to(x.dtype).clamp(min=1.0)
    return torch.cat([x, acc / cnt[:, None]], dim=1)


def spectral_concat(x, A_norm):
    return torch.cat([x, torch.sparse.mm(A_norm, x)], dim=1)


def mlp(h, weights, biases):
    """weights[i]: (out, in) as in nn.Linear; ReLU between layers, none after the last."""
    for li, (W, b) in enumerate(zip(weights, biases)):
        h = torch.nn.functional.linear(h, W, b)
        if li + 1 < len(weights):
            h = torch.relu(h)
    return h


def residual_ortho_loss(U_pred, K_list, M_list, offsets, w_res, w_orth, n_modes):
    total_res = torch.zeros((), dtype=torch.float32)
    total_orth = torch.zeros((), dtype=torch.float32)
    lams = []
    eye = torch.eye(n_modes)
    for K, M, off in zip(K_list, M_list, offsets):
        U = U_pred[int(off):int(off) + K.shape[0]]
        Kt, Mt = coo_f32(K), coo_f32(M)
        MU = torch.sparse.mm(Mt, U)
        KU = torch.sparse.mm(Kt, U)
        lam = (U * KU).sum(0) / ((U * MU).sum(0) + 1e-12)
        lams.append(lam)
        R = KU - MU * lam[None, :]
        total_res = total_res + (R ** 2).mean()
        G = U.t() @ MU
        total_orth = total_orth + ((G - eye) ** 2).sum() / n_modes
    return w_res * total_res, w_orth * total_orth, lams


def eigenvalue_losses(lam0, lam_target, w_proj, w_trace, w_order, w_eigen):
    trace = lam0.mean()
    order = torch.relu(-(lam0[1:] - lam0[:-1])).sum()
    eigen = ((lam0 - lam_target) ** 2).mean() if lam_target is not None else torch.zeros(())
    proj = torch.zeros(())
    return w_proj * proj, w_trace * trace, w_order * order, w_eigen * eigen


def m_normalize(U, M):
    MU = torch.sparse.mm(coo_f32(M), U)
    return U / torch.sqrt((U * MU).sum(0) + 1e-12)[None, :]


def rayleigh_ritz(U_np, K, M):
    from scipy.linalg import eigh
    U = torch.from_numpy(np.asarray(U_np)).to(torch.float32)
    A = (U.t() @ torch.sparse.mm(coo_f32(K), U)).numpy()
    B = (U.t() @ torch.sparse.mm(coo_f32(M), U)).numpy()
    vals, C = eigh(A, B)
    return vals, U.numpy() @ C


def level_features(X, U_norm, lam, edge_index, K, M, level_idx, n_levels):
    n = X.shape[0]
    Xt = torch.from_numpy(np.asarray(X)).to(torch.float32)
    res_level = torch.full((n, 1), float(n_levels - 1 - level_idx))
    deg = torch.bincount(edge_index[0], minlength=n).to(torch.float32)[:, None]
    deg = deg / (deg.max() + 1e-12)
    Kd = torch.from_numpy(np.asarray(K.diagonal())).to(torch.float32)[:, None]
    Md = torch.from_numpy(np.asarray(M.diagonal())).to(torch.float32)[:, None]
    KU = coo_f32(K) @ U_norm
    MU = coo_f32(M) @ U_norm
    rmag = torch.norm(KU - MU * lam[None, :], dim=1, keepdim=True)
    rmag = rmag / (rmag.max() + 1e-12)
    ray = (U_norm * KU).sum(1, keepdim=True) / ((U_norm * MU).sum(1, keepdim=True) + 1e-12)
    ray = ray / (lam.max() + 1e-12)
    return torch.cat([Xt, res_level, deg, Kd, Md, rmag, ray, U_norm], dim=1)


def gcn_norm_adjacency(edge_index, n):
    ones = torch.ones(edge_index.shape[1])
    A = torch.sparse_coo_tensor(edge_index, ones, (n, n)).coalesce()
    d = torch.arange(n).repeat(2, 1)
    A_hat = (A + torch.sparse_coo_tensor(d, torch.ones(n), (n, n)).coalesce()).coalesce()
    deg = torch.bincount(A_hat.indices()[0], minlength=n).float()
    dis = torch.pow(deg.clamp(min=1e-12), -0.5)
    D = torch.sparse_coo_tensor(d, dis, (n, n)).coalesce()
    return torch.sparse.mm(D, torch.sparse.mm(A_hat, D))


class CorrectorTrainer:
    """One-object restatement of the reference epoch body: forward, scale ramp, loss,
    backward, clip, Adam(coupled L2), ReduceLROnPlateau stepped on the raw loss."""

    def __init__(self, x_feats, edge_index, U_base, K_list, M_list, lam_target,
                 hidden, n_modes, model_type="simple", A_norm=None,
                 lr=1e-3, weight_decay=1e-5, corr_scale=10.0, w_res=1000.0, w_orth=10.0,
                 w_proj=0.0, w_trace=0.0, w_order=0.0, w_eigen=0.0, grad_clip=10.0, seed=0):
        self.x, self.ei, self.A_norm = x_feats, edge_index, A_norm
        self.U_base, self.K_list, self.M_list = U_base, K_list, M_list
        self.lam_target = lam_target
        self.model_type, self.k = model_type, n_modes
        self.corr_scale, self.grad_clip = corr_scale, grad_clip
        self.w = (w_res, w_orth, w_proj, w_trace, w_order, w_eigen)
        self.offsets = [0] + list(np.cumsum([K.shape[0] for K in K_list[:-1]]))
        g = torch.Generator().manual_seed(seed)
        dims = [2 * x_feats.shape[1]] + list(hidden) + [n_modes]
        self.weights, self.biases = [], []
        for i in range(len(dims) - 1):
            lin = torch.nn.Linear(dims[i], dims[i + 1])
            with torch.no_grad():
                bound = 1.0 / np.sqrt(dims[i])
                lin.weight.uniform_(-bound, bound, generator=g)
                lin.bias.uniform_(-bound, bound, generator=g)
                if i == len(dims) - 2:
                    lin.weight.normal_(0.0, 0.01, generator=g)
                    lin.bias.zero_()
            self.weights.append(lin.weight)
            self.biases.append(lin.bias)
        self.params = [p for pair in zip(self.weights, self.biases) for p in pair]
        self.opt = torch.optim.Adam(self.params, lr=lr, weight_decay=weight_decay)
        self.sched = torch.optim.lr_scheduler.ReduceLROnPlateau(
            self.opt, mode="min", factor=0.5, patience=2000, min_lr=1e-6)
        self.epoch = 0

    def load_parameters(self, weights, biases):
        with torch.no_grad():
            for p, w in zip(self.weights, weights):
                p.copy_(torch.as_tensor(w))
            for p, b in zip(self.biases, biases):
                p.copy_(torch.as_tensor(b))

    def forward(self):
        if self.model_type == "simple":
            h = neighbour_mean_concat(self.x, self.ei)
        else:
            h = spectral_concat(self.x, self.A_norm)
        return mlp(h, self.weights, self.biases)

    def losses(self, epoch=None):
        epoch = self.epoch if epoch is None else epoch
        scale = self.corr_scale * min(1.0, epoch / 5000.0)
        U_pred = self.U_base + scale * self.forward()
        w_res, w_orth, w_proj, w_trace, w_order, w_eigen = self.w
        l_res, l_orth, lams = residual_ortho_loss(U_pred, self.K_list, self.M_list, self.offsets,
                                                  w_res, w_orth, self.k)
        extra = eigenvalue_losses(lams[0], self.lam_target, w_proj, w_trace, w_order, w_eigen)
        total = l_res + l_orth + sum(extra)
        return total, l_res, l_orth, lams, U_pred

    def step(self):
        self.opt.zero_grad()
        total, l_res, l_orth, lams, _ = self.losses()
        total.backward()
        torch.nn.utils.clip_grad_norm_(self.params, self.grad_clip)
        self.opt.step()
        self.sched.step(total.item())
        self.epoch += 1
        return total.item(), l_res.item(), l_orth.item(), [l.detach() for l in lams]
