"""CPU: host-side containers of the product (no kernels are launched): CSR conversion, shared-pattern pairs,
symmetry detection, edge-list CSR, flat parameter buffer."""
import importlib

import numpy as np
import scipy.sparse as sp
import torch

sparse = importlib.import_module("eigen-pinns_b200.sparse")
engine = importlib.import_module("eigen-pinns_b200.engine")
synthetic = importlib.import_module("eigen-pinns_b200.synthetic")
fem = importlib.import_module("eigen-pinns_b200.fem")


def test_csr_conversion_casts_then_merges_duplicates():
    A = sp.coo_matrix((np.array([1.0, 2.0, 1e-9, 3.0]), ([0, 0, 0, 2], [1, 1, 1, 0])), shape=(3, 3))
    C = sparse.CsrMatrix.from_scipy(A, "cpu")
    assert C.nnz == 2 and C.rowptr.dtype == torch.int32 and C.val.dtype == torch.float32
    assert C.val[0].item() == np.float32(np.float32(1.0) + np.float32(2.0) + np.float32(1e-9))
    assert C.rowptr.tolist() == [0, 1, 1, 2] and C.col.tolist() == [1, 0]


def test_operator_pair_shared_pattern_and_symmetry():
    v, t = synthetic.icosphere(3)
    K, M = fem.assemble_stiffness_mass(v, t)
    pair = sparse.OperatorPair(K, M, "cpu")
    assert pair.shared and pair.symmetric and pair.KT is pair.K and pair.MT is pair.M
    # different patterns are padded to the union with explicit zeros; non-symmetric -> explicit transposes
    A = sp.random(40, 40, density=0.1, random_state=1, format="csr")
    B = sp.random(40, 40, density=0.05, random_state=2, format="csr")
    p2 = sparse.OperatorPair(A, B, "cpu")
    assert p2.shared and not p2.symmetric
    assert torch.equal(p2.K.col, p2.M.col) and p2.K.nnz >= max(A.nnz, B.nnz)
    dense = lambda c: sp.csr_matrix((c.val.numpy(), c.col.numpy(), c.rowptr.numpy()), shape=c.shape).toarray()
    np.testing.assert_allclose(dense(p2.K), A.toarray().astype(np.float32))
    np.testing.assert_allclose(dense(p2.M), B.toarray().astype(np.float32))
    np.testing.assert_allclose(dense(p2.KT), A.T.toarray().astype(np.float32))
    assert torch.equal(p2.KT.col, p2.MT.col)


def test_edge_index_csr_keeps_edge_order_and_duplicates():
    ei = torch.tensor([[2, 0, 2, 1, 2], [5, 1, 3, 0, 5]])
    adj = sparse.CsrMatrix.from_edge_index(ei, 4, "cpu")
    assert adj.rowptr.tolist() == [0, 1, 2, 5, 5]
    assert adj.col.tolist() == [1, 0, 5, 3, 5]           # row 2 keeps the original order 5, 3, 5 (index_add_ order)


def test_flat_params_views_follow_the_flat_buffer():
    lins = [torch.nn.Linear(5, 7), torch.nn.Linear(7, 3)]
    before = [l.weight.detach().clone() for l in lins]
    fp = engine.FlatParams.adopt(lins)
    assert fp.flat.numel() == 5 * 7 + 7 + 7 * 3 + 3 and fp.dims == [5, 7, 3]
    assert all(torch.equal(l.weight, b) for l, b in zip(lins, before))
    fp.flat.mul_(2.0)                                    # what the optimiser kernel does, in place
    assert torch.equal(lins[0].weight, 2 * before[0]) and torch.equal(lins[1].weight, 2 * before[1])
    sd = torch.nn.Sequential(*lins).state_dict()
    assert torch.equal(sd["0.weight"], 2 * before[0])
    assert fp.dW[1].shape == lins[1].weight.shape and fp.dW[1].data_ptr() != fp.W[1].data_ptr()


def test_synthetic_meshes():
    for f in (1, 2, 5):
        v, t = synthetic.icosphere(f)
        assert v.shape == (10 * f * f + 2, 3) and t.shape == (20 * f * f, 3)
        assert np.allclose(np.linalg.norm(v, axis=1), 1.0)
        e = set()
        for a, b in ((0, 1), (1, 2), (2, 0)):
            e.update(zip(np.minimum(t[:, a], t[:, b]).tolist(), np.maximum(t[:, a], t[:, b]).tolist()))
        assert v.shape[0] - len(e) + t.shape[0] == 2                    # Euler characteristic of the sphere
    v, t = synthetic.torus(12, 8)
    K, M = fem.assemble_stiffness_mass(v, t)
    assert K.nnz == 7 * 96 and abs(np.asarray(K.sum(1))).max() < 1e-12
    Y, deg = synthetic.real_spherical_harmonics(synthetic.icosphere(4)[0], 9)
    assert deg.tolist() == [0, 1, 1, 1, 2, 2, 2, 2, 2]
