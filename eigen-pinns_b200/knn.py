"""k nearest neighbours, prolongation and Jacobi smoothing on the GPU: the pre-processing that the
reference does with scikit-learn and Python double loops (src/utils.py:39-75, :220-232).

  knn(ref, query, k)            grid hash + fp64 distances, ties by index        -> ep_knn_grid_f64
  knn_graph(X, k)               (2, n k) edge list, self excluded                 (reference build_knn_graph)
  prolongation(Xc, Xf, k)       inverse-distance weights, rows sum to one, CSR    (reference build_prolongation)
  jacobi_smooth(M, K, U, ...)   sweeps on (M + alpha K) U = M U_rough, each sweep one SpMM (ep_spmm_csr_f32)
"""
import ctypes

import numpy as np
import torch

from . import ops
from ._cabi import call, EpError
from .sparse import CsrMatrix


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _pts(points, device):
    if torch.is_tensor(points):
        if not points.is_cuda:
            raise EpError("device point set expected (no CPU fallback exists)")
        return points.to(torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(points, dtype=np.float64)).to(device)


def knn(ref, query, k, device="cuda", with_dist=True):
    """Indices (n_query x k, int64, sorted by distance then index) of the k nearest reference points and distances."""
    R, Q = _pts(ref, device), _pts(query, device)
    n_ref = R.shape[0]
    if k > n_ref:
        raise ValueError("k = %d neighbours requested from %d points" % (k, n_ref))
    lo = R.min(0).values
    ext = (R.max(0).values - lo).clamp_min(1e-12)
    # surface point sets: aim at ~4 points per occupied cell, assuming a 2-D manifold in the bounding box
    area_guess = float(2.0 * (ext[0] * ext[1] + ext[1] * ext[2] + ext[0] * ext[2]) / 3.0)
    cell = max(float(np.sqrt(4.0 * area_guess / max(n_ref, 1))), float(ext.max()) / 512.0)
    dims = torch.clamp((ext / cell).floor().long() + 1, min=1)
    c = torch.minimum(((R - lo) / cell).floor().long().clamp_min(0), dims - 1)
    cid = (c[:, 0] * dims[1] + c[:, 1]) * dims[2] + c[:, 2]
    order = torch.argsort(cid, stable=True)
    n_cells = int(dims.prod().item())
    cell_start = torch.searchsorted(cid[order], torch.arange(n_cells + 1, device=R.device))
    out_idx = torch.empty((Q.shape[0], k), dtype=torch.int64, device=R.device)
    out_dist = torch.empty((Q.shape[0], k), dtype=torch.float64, device=R.device) if with_dist else None
    lo_h = (ctypes.c_double * 3)(*lo.cpu().tolist())
    dims_h = (ctypes.c_int64 * 3)(*dims.cpu().tolist())
    call("ep_knn_grid_f64", Q.shape[0], _p(Q), n_ref, _p(R), _p(order), _p(cell_start), lo_h, float(cell), dims_h, int(k),
         _p(out_idx), _p(out_dist), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    return out_idx, out_dist


def knn_graph(X, k, device="cuda"):
    """(2, n k) int64 edge list: row i lists the k nearest neighbours of point i, self excluded (the nearest hit)."""
    idx, _ = knn(X, X, k + 1, device, with_dist=False)
    n = idx.shape[0]
    rows = torch.arange(n, device=idx.device).repeat_interleave(k)
    return torch.stack([rows, idx[:, 1:].reshape(-1)])


def prolongation(X_coarse, X_fine, k, device="cuda"):
    """CSR interpolation matrix (n_fine x n_coarse): weights 1 / (d + 1e-12), rows normalised to one."""
    idx, dist = knn(X_coarse, X_fine, k, device)
    w = 1.0 / (dist + 1e-12)
    w = w / w.sum(1, keepdim=True)
    n_f, n_c = idx.shape[0], (X_coarse.shape[0])
    col, perm = torch.sort(idx, dim=1)                            # CSR wants ascending columns inside a row
    val = torch.gather(w, 1, perm)
    rowptr = (torch.arange(n_f + 1, device=idx.device) * k).to(torch.int32)
    return CsrMatrix.from_device_arrays(rowptr, col.reshape(-1).to(torch.int32).contiguous(),
                                        val.reshape(-1).to(torch.float32).contiguous(), (n_f, n_c), symmetric=False)


def jacobi_smooth(M: CsrMatrix, K: CsrMatrix, U_rough, alpha=0.05, n_iters=5):
    """U <- U + D^-1 (M U_rough - (M + alpha K) U), D = diag(M) + alpha diag(K); M and K share one sparsity pattern
    (FEM pair), so A = M + alpha K is formed on the value arrays.  fp32 on the device."""
    U0 = U_rough.to(torch.float32).contiguous()
    A = CsrMatrix.from_device_arrays(M.rowptr, M.col, (M.val + alpha * K.val).contiguous(), M.shape, symmetric=True)
    n = M.shape[0]
    counts = (M.rowptr[1:] - M.rowptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(n, device=M.val.device), counts)
    diag = torch.zeros(n, dtype=torch.float32, device=M.val.device)
    on_diag = M.col.long() == rows
    diag.index_add_(0, rows[on_diag], A.val[on_diag])
    d_inv = (1.0 / (diag + 1e-12)).unsqueeze(1)
    rhs = ops.spmm(M, U0)
    U = U0.clone()
    for _ in range(n_iters):
        U = U + d_inv * (rhs - ops.spmm(A, U))
    return U
