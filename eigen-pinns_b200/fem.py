"""Sparse linear-FEM operators on triangle meshes (host side, float64).

The reference assembles the stiffness matrix K (cotangent Laplacian) and the
consistent mass matrix M as two *dense* N x N arrays with a Python loop over
triangles (reference src/Mesh.py:348-364, element formulas at :180-198 and
:228-234).  That cannot produce the 1 M / 16 M-vertex operators the hot path is
benchmarked on, so this module restates the same element formulas vectorised
over all triangles and scatters straight to CSR.

Element (p0, p1, p2), local frame e1 = (p1-p0)/|p1-p0|, e2 = unit part of
(p2-p0) orthogonal to e1:
    B = [[y23, y31, y12], [x32, x13, x21]],  J = x13*y23 - y31*x32 (= 2*area)
    k_el = B^T B / (2 J)          m_el = J/12 * [[2,1,1],[1,2,1],[1,1,2]]
"""
import numpy as np
import scipy.sparse as sp


def element_matrices(verts, tris):
    """Per-triangle 3x3 stiffness and mass blocks.  Returns (k_el, m_el, J) with
    shapes (T,3,3), (T,3,3), (T,)."""
    verts = np.asarray(verts, dtype=np.float64)
    tris = np.asarray(tris)
    p0, p1, p2 = verts[tris[:, 0]], verts[tris[:, 1]], verts[tris[:, 2]]
    d10 = p1 - p0
    d20 = p2 - p0
    e1 = d10 / np.linalg.norm(d10, axis=1)[:, None]
    e2 = d20 - np.einsum("ij,ij->i", d20, e1)[:, None] * e1
    e2 = e2 / np.linalg.norm(e2, axis=1)[:, None]

    def dot(a, b):
        return np.einsum("ij,ij->i", a, b)

    x21 = dot(d10, e1)
    x13 = dot(p0 - p2, e1)
    x32 = dot(p2 - p1, e1)
    y23 = dot(p1 - p2, e2)
    y31 = dot(d20, e2)
    y12 = dot(p0 - p1, e2)
    J = x13 * y23 - y31 * x32
    B = np.stack([np.stack([y23, y31, y12], axis=1),
                  np.stack([x32, x13, x21], axis=1)], axis=1)      # (T,2,3)
    k_el = np.einsum("tia,tib->tab", B, B) / (2.0 * J)[:, None, None]
    pattern = np.array([[2.0, 1.0, 1.0], [1.0, 2.0, 1.0], [1.0, 1.0, 2.0]])
    m_el = pattern[None, :, :] * (J / 12.0)[:, None, None]
    return k_el, m_el, J


def assemble_stiffness_mass(verts, tris):
    """CSR (K, M), float64, sorted column indices, identical sparsity pattern
    for both (so the dual-operator SpMM can share rowptr/col).

    Sparse restatement of reference Mesh.computeLaplacian (src/Mesh.py:348-364):
    K[tri[a], tri[b]] += k_el[a, b], same for M.
    """
    verts = np.asarray(verts, dtype=np.float64)
    tris = np.asarray(tris, dtype=np.int64)
    n = verts.shape[0]
    k_el, m_el, _ = element_matrices(verts, tris)
    rows = np.repeat(tris, 3, axis=1).ravel()            # tri[a] for (a,b)
    cols = np.tile(tris, (1, 3)).ravel()                 # tri[b] for (a,b)
    K = sp.coo_matrix((k_el.ravel(), (rows, cols)), shape=(n, n)).tocsr()
    M = sp.coo_matrix((m_el.ravel(), (rows, cols)), shape=(n, n)).tocsr()
    K.sum_duplicates()
    M.sum_duplicates()
    K.sort_indices()
    M.sort_indices()
    # Both come from the same (row, col) list, so the patterns coincide.
    assert K.nnz == M.nnz
    return K, M


def normalize_verts(verts):
    """Centre and divide by the largest per-axis standard deviation
    (reference src/mesh_helpers.py:9-13)."""
    verts = np.asarray(verts, dtype=np.float64)
    centroid = verts.mean(0)
    std_max = verts.std(0).max() + 1e-12
    return (verts - centroid) / std_max


def connectivity_edges(tris):
    """Unique directed edge list (2, E) int64 from triangle connectivity, both
    directions, lexicographically sorted by (row, col) — vectorised equivalent of
    reference mesh_helpers.mesh_to_edge_index (src/mesh_helpers.py:66-90)."""
    tris = np.asarray(tris, dtype=np.int64)
    a = np.concatenate([tris[:, 0], tris[:, 1], tris[:, 2], tris[:, 1], tris[:, 2], tris[:, 0]])
    b = np.concatenate([tris[:, 1], tris[:, 2], tris[:, 0], tris[:, 0], tris[:, 1], tris[:, 2]])
    n = int(tris.max()) + 1
    key = np.unique(a * n + b)
    return np.stack([key // n, key % n])
