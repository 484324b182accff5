"""Synthetic meshes and initial subspaces for the benchmark configurations
(BASELINE.json configs 3-5).  Host side, float64.

* geodesic icosphere of frequency f: V = 10 f^2 + 2, F = 20 f^2  (f=316 -> 998,562)
* torus grid nu x nv, every quad split on the same diagonal, valence 6
* real spherical harmonics (analytic Laplace-Beltrami eigenfunctions of the sphere)
"""
import numpy as np


def _icosahedron():
    t = (1.0 + np.sqrt(5.0)) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0],
                  [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1)[:, None]
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
                  [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                  [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
                  [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    return v, f


def icosphere(freq):
    """Class-I geodesic sphere.  Vertices are shared exactly (integer keys) between
    faces; positions are (i*A + j*B + k*C)/freq projected to the unit sphere.
    Returns (verts float64 (V,3), tris int64 (F,3)) with V = 10 f^2 + 2."""
    f = int(freq)
    assert f >= 1
    V0, F0 = _icosahedron()
    n_corner = 12
    edges = {}
    for tri in F0:
        for a, b in ((tri[0], tri[1]), (tri[1], tri[2]), (tri[2], tri[0])):
            key = (min(a, b), max(a, b))
            if key not in edges:
                edges[key] = len(edges)
    n_edge_pts = f - 1
    n_face_pts = (f - 1) * (f - 2) // 2
    edge_base = n_corner
    face_base = edge_base + len(edges) * n_edge_pts
    n_verts = face_base + 20 * n_face_pts
    verts = np.empty((n_verts, 3), dtype=np.float64)
    verts[:12] = V0
    # edge points, canonical orientation (lo -> hi), t = 1..f-1 steps from lo
    tt = np.arange(1, f, dtype=np.float64)[:, None]
    for (lo, hi), e in edges.items():
        p = ((f - tt) * V0[lo] + tt * V0[hi]) / f
        verts[edge_base + e * n_edge_pts: edge_base + (e + 1) * n_edge_pts] = p

    def edge_vertex(a, b, steps_from_a):
        lo, hi = (a, b) if a < b else (b, a)
        t = steps_from_a if a == lo else f - steps_from_a
        return edge_base + edges[(lo, hi)] * n_edge_pts + (t - 1)

    # local grid ids for every face: (i, j) with weight i on B, j on C, f-i-j on A
    ii, jj = np.meshgrid(np.arange(f + 1), np.arange(f + 1), indexing="ij")
    valid = (ii + jj) <= f
    tris_out = []
    for fi, (A, B, C) in enumerate(F0):
        gid = -np.ones((f + 1, f + 1), dtype=np.int64)
        gid[0, 0], gid[f, 0], gid[0, f] = A, B, C
        if f > 1:
            s = np.arange(1, f)
            gid[s, 0] = [edge_vertex(A, B, int(x)) for x in s]         # A->B edge (j=0)
            gid[0, s] = [edge_vertex(A, C, int(x)) for x in s]         # A->C edge (i=0)
            gid[f - s, s] = [edge_vertex(B, C, int(x)) for x in s]     # B->C edge (i+j=f)
        interior = valid & (ii > 0) & (jj > 0) & (ii + jj < f)
        n_int = int(interior.sum())
        assert n_int == n_face_pts
        if n_int:
            ids = face_base + fi * n_face_pts + np.arange(n_int)
            gid[interior] = ids
            wi = ii[interior].astype(np.float64)[:, None]
            wj = jj[interior].astype(np.float64)[:, None]
            verts[ids] = ((f - wi - wj) * V0[A] + wi * V0[B] + wj * V0[C]) / f
        # upward triangles (i,j),(i+1,j),(i,j+1) for i+j <= f-1
        up = (ii + jj) <= f - 1
        ui, uj = ii[up], jj[up]
        tris_out.append(np.stack([gid[ui, uj], gid[ui + 1, uj], gid[ui, uj + 1]], axis=1))
        # downward triangles (i+1,j),(i+1,j+1),(i,j+1) for i+j <= f-2
        dn = (ii + jj) <= f - 2
        di, dj = ii[dn], jj[dn]
        if di.size:
            tris_out.append(np.stack([gid[di + 1, dj], gid[di + 1, dj + 1], gid[di, dj + 1]], axis=1))
    tris = np.concatenate(tris_out, axis=0)
    assert tris.min() >= 0
    verts /= np.linalg.norm(verts, axis=1)[:, None]
    return verts, tris


def torus(nu, nv, R=1.0, r=0.4):
    """Periodic nu x nv grid on a torus, vertex id = i*nv + j, quads split on the
    (i,j)-(i+1,j+1) diagonal.  V = nu*nv, F = 2*nu*nv, valence 6 everywhere."""
    u = 2.0 * np.pi * np.arange(nu) / nu
    v = 2.0 * np.pi * np.arange(nv) / nv
    U, Vv = np.meshgrid(u, v, indexing="ij")
    x = (R + r * np.cos(Vv)) * np.cos(U)
    y = (R + r * np.cos(Vv)) * np.sin(U)
    z = r * np.sin(Vv)
    verts = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    i, j = i.ravel(), j.ravel()
    ip, jp = (i + 1) % nu, (j + 1) % nv
    v00, v10, v11, v01 = i * nv + j, ip * nv + j, ip * nv + jp, i * nv + jp
    tris = np.concatenate([np.stack([v00, v10, v11], axis=1),
                           np.stack([v00, v11, v01], axis=1)], axis=0).astype(np.int64)
    return verts, tris


def real_spherical_harmonics(unit_xyz, n_modes):
    """First n_modes real spherical harmonics (l = 0,1,2,... each with 2l+1 members)
    evaluated on unit vectors.  Columns are L2(S^2)-orthonormal; they are the exact
    Laplace-Beltrami eigenfunctions of a sphere of radius rho with eigenvalue
    l(l+1)/rho^2.  Returns (Y (N, n_modes) float64, degrees l (n_modes,))."""
    from scipy.special import sph_harm_y
    x, y, z = unit_xyz[:, 0], unit_xyz[:, 1], unit_xyz[:, 2]
    theta = np.arccos(np.clip(z, -1.0, 1.0))       # polar
    phi = np.arctan2(y, x)                         # azimuth
    cols, degs = [], []
    l = 0
    while len(cols) < n_modes:
        for m in range(-l, l + 1):
            Y = sph_harm_y(l, abs(m), theta, phi)
            if m < 0:
                c = np.sqrt(2.0) * (-1) ** m * Y.imag
            elif m == 0:
                c = Y.real
            else:
                c = np.sqrt(2.0) * (-1) ** m * Y.real
            cols.append(np.ascontiguousarray(c, dtype=np.float64))
            degs.append(l)
            if len(cols) == n_modes:
                break
        l += 1
    return np.stack(cols, axis=1), np.array(degs)


def torus_trial_modes(nu, nv, n_modes):
    """Smooth trial functions cos/sin(a u) cos/sin(b v) on the torus grid, lowest
    wavenumbers first — a starting subspace (no closed-form spectrum exists)."""
    u = 2.0 * np.pi * (np.arange(nu * nv) // nv) / nu
    v = 2.0 * np.pi * (np.arange(nu * nv) % nv) / nv
    cols = []
    order = sorted(((a * a + 6.25 * b * b, a, b) for a in range(0, 12) for b in range(0, 6)))
    for _, a, b in order:
        for fu in ((np.cos, np.sin) if a else (np.cos,)):
            for fv in ((np.cos, np.sin) if b else (np.cos,)):
                cols.append(fu(a * u) * fv(b * v))
                if len(cols) == n_modes:
                    return np.stack(cols, axis=1)
    raise ValueError("n_modes too large for torus_trial_modes")
