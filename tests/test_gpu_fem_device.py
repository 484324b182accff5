"""GPU sparse FEM assembly against the host assembly, which is itself pinned to the
reference's dense Mesh.computeLaplacian on bunny.obj."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import pkg, dev

pytestmark = pytest.mark.gpu


def _check(verts, tris):
    fem, femd, ops = pkg("fem"), pkg("fem_device"), pkg("ops")
    K, M = fem.assemble_stiffness_mass(verts, tris)
    pair, vK, vM = femd.assemble(verts, tris, dev(), keep_fp64=True)
    assert np.array_equal(pair.K.rowptr.cpu().numpy(), K.indptr)
    assert np.array_equal(pair.K.col.cpu().numpy(), K.indices)
    assert pair.shared and pair.symmetric
    np.testing.assert_allclose(vK.cpu().numpy(), K.data, rtol=0, atol=2e-13 * np.abs(K.data).max())
    np.testing.assert_allclose(vM.cpu().numpy(), M.data, rtol=0, atol=2e-15 * max(1.0, np.abs(M.data).max()))
    pair2, vK2, vM2 = femd.assemble(verts, tris, dev(), keep_fp64=True)
    assert torch.equal(vK, vK2) and torch.equal(vM, vM2)                      # deterministic (ordered sums)
    U = torch.randn(verts.shape[0], 16, device=dev())
    KU, MU = ops.spmm2(pair, U)
    ref = K.astype(np.float32) @ U.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(KU.cpu().numpy(), ref, rtol=2e-5, atol=2e-4)
    return K, vK


def test_bunny_matches_reference_operator():
    g = load_golden("bunny_fem.npz")
    K, vK = _check(g["verts"], g["tris"])
    # the fixture holds the reference's dense assembly (CSR of it)
    np.testing.assert_allclose(vK.cpu().numpy(), g["K_data"], rtol=0, atol=5e-13 * np.abs(g["K_data"]).max())


@pytest.mark.parametrize("freq", [1, 7, 40])
def test_icosphere(freq):
    syn = pkg("synthetic")
    v, t = syn.icosphere(freq)
    _check(v, t)


def test_torus_device_generated_large():
    """4 M-vertex torus assembled entirely on the device: row sums of K vanish, total mass = 2 * area."""
    femd, ops = pkg("fem_device"), pkg("ops")
    nu = nv = 2048
    i = torch.arange(nu, device=dev()).repeat_interleave(nv)
    j = torch.arange(nv, device=dev()).repeat(nu)
    u, w = 2 * np.pi * i.double() / nu, 2 * np.pi * j.double() / nv
    R, r = 1.0, 0.4
    verts = torch.stack([(R + r * torch.cos(w)) * torch.cos(u), (R + r * torch.cos(w)) * torch.sin(u), r * torch.sin(w)], 1)
    ip, jp = (i + 1) % nu, (j + 1) % nv
    v00, v10, v11, v01 = i * nv + j, ip * nv + j, ip * nv + jp, i * nv + jp
    tris = torch.cat([torch.stack([v00, v10, v11], 1), torch.stack([v00, v11, v01], 1)]).to(torch.int32)
    pair = femd.assemble(verts, tris, dev())
    n = nu * nv
    assert pair.K.nnz == 7 * n
    ones = torch.ones(n, 4, device=dev())
    K1, M1 = ops.spmm2(pair, ones)
    assert K1.abs().max().item() < 1e-2 * pair.K.val.abs().max().item()
    assert M1[:, 0].double().sum().item() == pytest.approx(2 * 4 * np.pi ** 2 * R * r, rel=1e-4)
