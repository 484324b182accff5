"""Graph / linear-algebra helpers around the hot path (pre-processing side).

Drop-in for reference src/utils.py: same function names and argument meaning.  The Python
double loops of the reference (`build_prolongation` :39-60, `build_knn_graph` :63-75) are
vectorised; `scipy_sparse_to_torch_sparse` (:14-20) keeps its contract for callers that want a
torch COO tensor, but the training path never calls it per epoch - operators are moved to the
GPU once as CSR (`device_operator` / `device_pair`).
"""
import numpy as np
import torch
from scipy.sparse import coo_matrix, identity, diags
from scipy.sparse.linalg import eigsh

import _backend

_sparse = _backend.module("sparse")

_pair_cache = {}


def device_pair(K, M, device):
    """OperatorPair (CSR fp32/int32 on the GPU) for a scipy (K, M), converted once per matrix object."""
    key = (id(K), id(M), K.shape, getattr(K, "nnz", None), str(device))
    hit = _pair_cache.get(key)
    if hit is None or hit[0] is not K or hit[1] is not M:
        hit = (K, M, _sparse.OperatorPair(K, M, device))
        _pair_cache[key] = hit
    return hit[2]


def device_operator(A, device):
    return _sparse.CsrMatrix.from_scipy(A, device)


def scipy_sparse_to_torch_sparse(A):
    A = A.tocoo()
    idx = torch.from_numpy(np.vstack((A.row, A.col)).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(A.data.astype(np.float32)), A.shape).coalesce()


def normalize_columns_np(U, eps=1e-12):
    norms = np.linalg.norm(U, axis=0) + eps
    return U / norms, norms


def normalize_columns_torch(U, eps=1e-12):
    norms = torch.norm(U, dim=0) + eps
    return U / norms, norms


def _knn(X_ref, X_query, k):
    """(distances, indices) of the k nearest reference points, sorted by distance.  With a GPU: the grid-hash kernel
    (ep_knn_grid_f64, ties by index); without one (pre-processing on a CPU-only host, CPU test-suite): scikit-learn,
    as in the reference.  Pre-processing is outside the training hot path, which has no CPU path at all."""
    if torch.cuda.is_available():
        knn_mod = _backend.module("knn")
        idx, dist = knn_mod.knn(np.asarray(X_ref, dtype=np.float64), np.asarray(X_query, dtype=np.float64), k)
        return dist.cpu().numpy(), idx.cpu().numpy()
    from sklearn.neighbors import NearestNeighbors
    nbrs = NearestNeighbors(n_neighbors=k, algorithm='auto').fit(X_ref)
    return nbrs.kneighbors(X_query)


def build_prolongation(X_coarse, X_fine, k):
    """Inverse-distance kNN interpolation P (n_fine x n_coarse), rows sum to one."""
    dist, idx = _knn(X_coarse, X_fine, k)
    w = 1.0 / (dist + 1e-12)
    w /= w.sum(axis=1, keepdims=True)
    n_fine, n_coarse = X_fine.shape[0], X_coarse.shape[0]
    rows = np.repeat(np.arange(n_fine), k)
    return coo_matrix((w.ravel(), (rows, idx.ravel())), shape=(n_fine, n_coarse))


def build_knn_graph(X, k):
    """(2, n*k) int64 edge list: row i lists the k nearest neighbours of point i (self excluded)."""
    n = X.shape[0]
    _, idx = _knn(X, X, k + 1)
    rows = np.repeat(np.arange(n, dtype=np.int64), k)
    return torch.from_numpy(np.stack([rows, idx[:, 1:].astype(np.int64).ravel()]))


def offset_edge_lists(edge_index_list, sizes):
    """Concatenate per-level edge lists WITH node offsets (`edge_index + node_offset` of the multigrid notebooks, SURVEY
    8a-bis) - src/multigrid_model.py:147-150 concatenates them un-offset (quirk Q3), which aggregates every level's
    edges into the first rows of the stacked feature matrix."""
    out, off = [], 0
    for ei, n in zip(edge_index_list, sizes):
        out.append(ei + off)
        off += int(n)
    return torch.cat(out, dim=1)


def build_A_norm(edge_index, n_nodes, device):
    """D^-1/2 (A + I) D^-1/2 with A the coalesced (multi-)adjacency and D counting the stored
    entries per row of A + I - the exact recipe of reference :78-124 - as a torch sparse tensor."""
    ei = edge_index.detach().cpu().numpy()
    A = coo_matrix((np.ones(ei.shape[1], dtype=np.float32), (ei[0], ei[1])), shape=(n_nodes, n_nodes)).tocsr()
    A_hat = (A + identity(n_nodes, dtype=np.float32, format="csr")).tocsr()
    deg = np.diff(A_hat.indptr).astype(np.float32)
    dis = np.power(np.clip(deg, 1e-12, None), np.float32(-0.5)).astype(np.float32)
    Dm = diags(dis).tocsr().astype(np.float32)
    A_norm = (Dm @ (A_hat @ Dm)).tocoo()
    idx = torch.from_numpy(np.vstack((A_norm.row, A_norm.col)).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(A_norm.data.astype(np.float32)),
                                   (n_nodes, n_nodes)).coalesce().to(device)


def sparse_block_diag(sparse_tensor_list):
    if not sparse_tensor_list:
        return None
    idx, val, r0, c0 = [], [], 0, 0
    for t in sparse_tensor_list:
        t = t.coalesce()
        shift = torch.tensor([[r0], [c0]], device=t.device)
        idx.append(t.indices() + shift)
        val.append(t.values())
        r0 += t.shape[0]
        c0 += t.shape[1]
    return torch.sparse_coo_tensor(torch.cat(idx, 1), torch.cat(val), (r0, c0)).coalesce()


def solve_eigenvalue_point_cloud(X, n_modes):
    from mesh_helpers import compute_laplacian_and_mass_matrices
    L, M = compute_laplacian_and_mass_matrices(X)
    vals, vecs = eigsh(L, k=n_modes, M=M, sigma=-1e-8, which='LM')
    return vals, np.array(vecs), L, M


def solve_eigenvalue_operators(K, M, n_modes):
    """Smallest generalised eigenpairs of a given (K, M) pair (shift-invert Lanczos; dense for tiny levels)."""
    n = K.shape[0]
    if n_modes >= n - 1:
        from scipy.linalg import eigh
        vals, vecs = eigh(K.toarray(), M.toarray())
        return vals[:n_modes], vecs[:, :n_modes]
    vals, vecs = eigsh(K.tocsc(), k=n_modes, M=M.tocsc(), sigma=-1e-8, which='LM')
    order = np.argsort(vals)
    return vals[order], np.array(vecs[:, order])


def solve_eigenvalue_mesh(mesh, n_modes):
    from mesh_helpers import compute_stiffness_and_mass_matrices
    K, M = compute_stiffness_and_mass_matrices(mesh)
    vals, vecs = eigsh(K.tocsc(), k=n_modes, M=M.tocsc(), sigma=-1e-8, which='LM')   # smallest modes
    return vals, np.array(vecs), K, M


def orthonormalize(U, M):
    """Gram-Schmidt in the M inner product."""
    Q = np.zeros_like(U)
    for i in range(U.shape[1]):
        v = U[:, i].copy()
        for j in range(i):
            v -= (Q[:, j] @ (M @ v)) * Q[:, j]
        Q[:, i] = v / (np.sqrt(v @ (M @ v)) + 1e-12)
    return Q


def jacobi_smooth_device(M, L, U_rough, alpha=0.05, n_iters=5, device="cuda"):
    """Same sweeps on the GPU in fp32 (each sweep = one CSR SpMM); for meshes where the host loop matters."""
    knn_mod = _backend.module("knn")
    pair = _sparse.OperatorPair(L, M, device)                 # K and M on one shared pattern
    U = knn_mod.jacobi_smooth(pair.M, pair.K, torch.from_numpy(np.asarray(U_rough, dtype=np.float32)).to(device),
                              alpha=alpha, n_iters=n_iters)
    return U.cpu().numpy().astype(np.float64)


def jacobi_smooth(M, L, U_rough, alpha=0.05, n_iters=5):
    """A few Jacobi sweeps on (M + alpha L) U = M U_rough."""
    A = (M + alpha * L).tocsr()
    rhs = M @ U_rough
    d_inv = 1.0 / (M.diagonal() + alpha * L.diagonal() + 1e-12)
    U = U_rough.copy()
    for _ in range(n_iters):
        U += d_inv[:, None] * (rhs - A @ U)
    return U
