"""GPU parity: farthest-point and voxel down-sampling must return bit-identical index sets."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import pkg, dev, dropin
from oracle import samplers_port

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["bunny", "ico8", "cloud", "grid"])
def test_fps_golden_bit_exact(tag):
    sampling = pkg("sampling")
    g = load_golden("samplers.npz")
    hier = [int(h) for h in g[f"fps_{tag}_hier"]]
    out = sampling.fps_levels(g[f"fps_{tag}_pts"], hier, int(g[f"fps_{tag}_start"]))
    assert sorted(out) == list(range(len(hier) + 1))
    for lv, idx in out.items():
        assert idx.dtype == np.int64 and np.array_equal(idx, g[f"fps_{tag}_lv{lv}"]), (tag, lv)


@pytest.mark.parametrize("tag", ["bunny", "cloud", "grid", "ico8"])
def test_voxel_golden_bit_exact(tag):
    sampling = pkg("sampling")
    g = load_golden("samplers.npz")
    hier = [int(h) for h in g[f"vox_{tag}_hier"]]
    out = sampling.voxel_levels(g[f"vox_{tag}_pts"], hier)
    for lv, idx in out.items():
        assert np.array_equal(idx, g[f"vox_{tag}_lv{lv}"]), (tag, lv)
    if tag == "bunny":
        assert [out[i].size for i in range(4)] == [256, 512, 1003, 2503]


@pytest.mark.parametrize("n,samples", [(1, 1), (2, 2), (1025, 40), (50000, 300), (300007, 64)])
def test_fps_order_vs_oracle(n, samples):
    sampling = pkg("sampling")
    pts = np.random.default_rng(n).standard_normal((n, 3))
    if n > 1000:
        pts[::7] = np.round(pts[::7], 1)                       # inject exact ties
    start = int(np.random.default_rng(42).integers(0, n))
    got = sampling.fps_order(pts, samples, start).cpu().numpy()
    ref = samplers_port.fps_order(pts, samples, start)
    assert np.array_equal(got, ref)


def test_fps_global_memory_variant_vs_oracle():
    """> 1.06 M points leaves the shared-memory-resident path; same answers required."""
    sampling = pkg("sampling")
    n = 1_200_000
    pts = np.random.default_rng(7).standard_normal((n, 3))
    got = sampling.fps_order(pts, 24, 5).cpu().numpy()
    assert np.array_equal(got, samplers_port.fps_order(pts, 24, 5))


def test_fps_million_points_properties():
    """BASELINE config 3 size: 1 M-point cloud.  Prefix of the oracle + size-independent properties
    (distinct picks, non-increasing covering radius)."""
    sampling = pkg("sampling")
    n = 1_000_000
    pts = np.random.default_rng(1234).standard_normal((n, 3))
    start = int(np.random.default_rng(42).integers(0, n))
    order = sampling.fps_order(pts, 1024, start).cpu().numpy()
    assert order[0] == start and np.unique(order).size == 1024
    assert np.array_equal(order[:48], samplers_port.fps_order(pts, 48, start))
    sel = pts[order]
    radius = []
    d = np.full(1024, np.inf)
    for i in range(1, 1024):
        d = np.minimum(d, np.linalg.norm(sel - sel[i - 1], axis=1))
        radius.append(d[i])
    assert all(radius[i] >= radius[i + 1] - 1e-12 for i in range(len(radius) - 1))


@pytest.mark.parametrize("n,target", [(20000, 64), (20000, 3000), (200000, 1024)])
def test_voxel_levels_vs_oracle(n, target):
    sampling = pkg("sampling")
    pts = np.random.default_rng(n + target).standard_normal((n, 3)) * np.array([1.0, 2.0, 0.5])
    got = sampling.voxel_levels(pts, [target])
    ref = samplers_port.voxel_levels(pts, [target])
    for lv in ref:
        assert np.array_equal(got[lv], ref[lv])


def test_dropin_sampler_functions_and_quirks():
    samplers = dropin("samplers")
    g = load_golden("samplers.npz")

    class M:
        verts = g["fps_bunny_pts"]
    out = samplers._farthest_point_sampling(M, [256, 512, 1024], start=int(g["fps_bunny_start"]))
    assert all(np.array_equal(out[lv], g[f"fps_bunny_lv{lv}"]) for lv in out)
    assert isinstance(samplers._farthest_point_sampling(M, [100, 5000]), np.ndarray)      # Q2
    vox = samplers._voxel_downsampling(M, [256, 512, 1024])
    assert [vox[i].size for i in range(4)] == [256, 512, 1003, 2503]
    vox2 = samplers._voxel_downsampling(M, [256, 9999])                                     # target >= N level
    assert np.array_equal(vox2[1], np.arange(2503))


def test_fps_host_buffer_entry_point():
    import ctypes
    cabi = pkg("_cabi")
    pts = np.ascontiguousarray(np.random.default_rng(3).standard_normal((5000, 3)))
    out = np.zeros(100, dtype=np.int64)
    cabi.call("ep_fps_f64_host", 5000, pts.ctypes.data_as(ctypes.c_void_p), 100, 17, out.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(out, samplers_port.fps_order(pts, 100, 17))


def test_voxel_million_points_vs_oracle():
    """BASELINE config 3 size for the voxel sampler (target 256, the reference's own probe configuration)."""
    sampling = pkg("sampling")
    pts = np.random.default_rng(1234).standard_normal((1_000_000, 3))
    got = sampling.voxel_levels(pts, [256])
    ref = samplers_port.voxel_levels(pts, [256])
    assert np.array_equal(got[0], ref[0]) and got[0].size <= 256
