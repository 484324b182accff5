"""TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product): fp64 torch-CPU restatement of the loss
functionals and networks of three reference notebooks (SURVEY.md 8 a-bis).

PARITY UNPINNED for this file: the notebooks are scripts with the formulas inline in their training loops - they
cannot be imported, they hold no fixtures, and no reference test covers them - so the formulas are restated here
from the cells cited per function and the CUDA path is compared with this restatement.

    dense_rayleigh_loss      /root/reference/scripts/simplified_loss.ipynb cell 0 (loop body: KU, MU, UKU, UMU,
                             rayleigh, loss_1, UMU_I, off_diag_loss, diag_loss)
    whitened_subspace_loss   /root/reference/scripts/loss_with_rigid_body.ipynb cell 0 (loop body from
                             "M-Orthogonalization via SVD" to "Total Loss"); operator preparation (epsilon, Frobenius
                             scaling) from the same cell
    single_mode_loss         /root/reference/delta_pinns_validation/iterative_eigenvalues_on_cloud.ipynb cell 1
                             (compute_eigenvalue_loss, compute_normalization_loss, compute_orthogonality_loss)
    smoothness_loss, AdaptiveCorrector   /root/reference/delta_pinns_validation/multigrid_gnn_refine_fixed.ipynb cell 4
                             (`train_gnn` loop body: L_corr, L_smooth_corr, L_smooth_total; class AdaptiveCorrector)
    CoordinateMLP, EigenfunctionNN   the `MLP` / `EigenfunctionNN` classes of the same cells
"""
import numpy as np
import scipy.sparse as sp
import torch
from scipy.sparse.linalg import norm as sparse_norm
import torch.nn as nn


def to_torch_sparse(A, dtype=torch.float64):
    A = sp.coo_matrix(A)
    idx = torch.from_numpy(np.vstack([A.row, A.col]).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(A.data).to(dtype), A.shape).coalesce()


def dense_rayleigh_loss(U, K, M, eps=1e-6):
    k = U.shape[1]
    KU, MU = torch.sparse.mm(K, U), torch.sparse.mm(M, U)
    UKU, UMU = U.T @ KU, U.T @ MU
    rayleigh = UKU / (UMU + eps)
    loss_1 = torch.mean(torch.norm((KU - torch.diag(rayleigh) * MU) ** 2))
    UMU_I = (UMU - torch.eye(k, dtype=U.dtype)) ** 2
    off_diag_loss, diag_loss = torch.max(UMU_I), torch.mean(UMU_I)
    return loss_1 + diag_loss + off_diag_loss, loss_1, diag_loss, off_diag_loss, torch.diag(rayleigh)


def frobenius_normalised(K, M, epsilon=1e-4):
    K_reg = sp.csr_matrix(K) + epsilon * sp.identity(K.shape[0], format="csr")
    K_scale, M_scale = sparse_norm(K_reg, "fro"), sparse_norm(sp.csr_matrix(M), "fro")
    return K_reg / K_scale, sp.csr_matrix(M) / M_scale, K_scale, M_scale


def whitened_subspace_loss(U, K, M, lambda_orth=1.0, lambda_zero=100.0, lambda_order=0.05, lambda_stability=0.1,
                           min_gap=1e-4):
    k = U.shape[1]
    identity_k = torch.eye(k, dtype=U.dtype)
    B = U.T @ torch.sparse.mm(M, U)
    V, S, _ = torch.linalg.svd(B)
    B_inv_sqrt = V @ torch.diag_embed(1.0 / torch.sqrt(torch.clamp(S, min=1e-7))) @ V.T
    U_orth = U @ B_inv_sqrt
    rayleigh_matrix = U_orth.T @ torch.sparse.mm(K, U_orth)
    sorted_eigs, _ = torch.sort(torch.diag(rayleigh_matrix))
    zero_eig_loss = sorted_eigs[0] ** 2
    eig_loss_trace = torch.sum(sorted_eigs[1:]) / (k - 1)
    gaps = sorted_eigs[1:] - sorted_eigs[:-1]
    diversity_loss = torch.sum(torch.relu(min_gap - gaps)) / (k - 1)
    eig_loss_offdiag = torch.sum((rayleigh_matrix * (1 - identity_k)) ** 2) / (k * (k - 1))
    eig_loss = lambda_zero * zero_eig_loss + 5.0 * eig_loss_trace + 2.0 * diversity_loss + eig_loss_offdiag
    B_orth = U_orth.T @ torch.sparse.mm(M, U_orth)
    orth_loss = torch.norm(B_orth - identity_k, p="fro") ** 2
    ordering_loss = torch.sum(torch.relu(sorted_eigs[:-1] - sorted_eigs[1:])) / k
    stability_loss = torch.relu(S.max() / (S.min() + 1e-10) - 1e3) / 1e3
    loss = eig_loss + lambda_orth * orth_loss + lambda_order * ordering_loss + lambda_stability * stability_loss
    terms = {"zero": zero_eig_loss, "trace": eig_loss_trace, "diversity": diversity_loss, "offdiag": eig_loss_offdiag,
             "orth": orth_loss, "ordering": ordering_loss, "stability": stability_loss}
    return loss, terms, sorted_eigs


def single_mode_loss(u, eigenvalue, L, M, previous=(), ortho_weight=1.0):
    u_flat = u.squeeze()
    Lu = torch.sparse.mm(L, u_flat.unsqueeze(1)).squeeze()
    Mu = torch.sparse.mm(M, u_flat.unsqueeze(1)).squeeze()
    eig = torch.mean((Lu - eigenvalue * Mu) ** 2)
    norm = (torch.dot(u_flat, Mu) - 1.0) ** 2
    ortho = torch.zeros((), dtype=u.dtype)
    for u_prev in previous:
        ortho = ortho + torch.dot(u_flat, torch.sparse.mm(M, u_prev.reshape(-1, 1)).squeeze()) ** 2
    return eig + norm + ortho_weight * ortho, eig, norm, ortho


def smoothness_loss(corr, U_pred, L, denom=None):
    n, k = U_pred.shape
    denom = float(n * k) if denom is None else float(denom)
    return torch.sum(corr * torch.sparse.mm(L, corr)) / denom, torch.sum(U_pred * torch.sparse.mm(L, U_pred)) / denom


class AdaptiveCorrector(nn.Module):
    def __init__(self, in_dim, out_dim, hidden_sizes=(128, 64, 32), init_scale=0.01):
        super().__init__()
        layers, prev = [], in_dim * 2
        for h in hidden_sizes:
            layers += [nn.Linear(prev, h), nn.ReLU()]
            prev = h
        layers.append(nn.Linear(prev, out_dim))
        self.net = nn.Sequential(*layers)
        self.mode_scales = nn.Parameter(torch.ones(out_dim) * init_scale)

    def forward(self, x, edge_index):
        row, col = edge_index
        agg = torch.zeros_like(x)
        agg.index_add_(0, row, x[col])
        deg = torch.bincount(row, minlength=x.shape[0]).unsqueeze(1).to(x.dtype).clamp(min=1.0)
        return self.net(torch.cat([x, agg / deg], dim=1)) * self.mode_scales.unsqueeze(0)


class Sin(nn.Module):
    def forward(self, x):
        return torch.sin(x)


class CoordinateMLP(nn.Module):
    def __init__(self, in_dim=3, out_dim=50, hidden=(256, 256, 128), activation="silu"):
        super().__init__()
        act = {"silu": nn.SiLU, "sin": Sin, "relu": nn.ReLU, "tanh": nn.Tanh}[activation]
        layers, last = [], in_dim
        for h in hidden:
            layers += [nn.Linear(last, h), act()]
            last = h
        layers.append(nn.Linear(last, out_dim))
        self.net = nn.Sequential(*layers)

    def forward(self, x):
        return self.net(x)


class EigenfunctionNN(nn.Module):
    def __init__(self, hidden_dim=64, input_dim=3, initial_eigenvalue=0.0):
        super().__init__()
        self.activation = Sin()
        self.eigenvalue_layer = nn.Linear(1, 1, bias=False)
        with torch.no_grad():
            self.eigenvalue_layer.weight.fill_(initial_eigenvalue)
        self.fc1 = nn.Linear(input_dim + 1, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim + 1, hidden_dim)
        self.fc3 = nn.Linear(hidden_dim + 1, hidden_dim)
        self.fc4 = nn.Linear(hidden_dim + 1, 1)

    def forward(self, x):
        eigenvalue = torch.abs(self.eigenvalue_layer(torch.ones(1, 1, dtype=x.dtype)))
        e = eigenvalue.expand(x.shape[0], 1)
        h = self.activation(self.fc1(torch.cat([x, e], dim=1)))
        h = self.activation(self.fc2(torch.cat([h, e], dim=1)))
        h = self.activation(self.fc3(torch.cat([h, e], dim=1)))
        return self.fc4(torch.cat([h, e], dim=1)), eigenvalue
