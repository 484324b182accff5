"""numpy fp64 restatement of the two down-samplers.  ORACLE — test infrastructure
only (see oracle/__init__.py); never imported by the product.

Reference anchors (under /root/reference/src/samplers.py):
  fps_order / fps_levels ......... :97-143   (_farthest_point_sampling)
  voxel_select / voxel_levels .... :9-94     (_voxel_downsampling)

The reference draws the FPS start vertex from an UNSEEDED generator (:113-116); here
it is an explicit argument so results are reproducible.  Bit-exactness notes that the
CUDA kernels must honour: distances are sqrt((dx*dx + dy*dy) + dz*dz) in IEEE double
with no fused multiply-add (numpy: norm = sqrt(add.reduce(x*x, axis=1))), running
minimum, then arg-max / arg-min with the FIRST index winning ties.
"""
import numpy as np

VOXEL_SCALES = (0.7, 0.85, 1.0, 1.15, 1.3, 1.5)


def _dist_to(points, p):
    d = points - p
    return np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])


def fps_order(points, n_samples, start):
    """Selection order of farthest-point sampling (length n_samples, int64)."""
    points = np.ascontiguousarray(points, dtype=np.float64)
    order = np.empty(n_samples, dtype=np.int64)
    order[0] = start
    nearest = np.full(points.shape[0], np.inf)
    for s in range(1, n_samples):
        np.minimum(nearest, _dist_to(points, points[order[s - 1]]), out=nearest)
        order[s] = int(np.argmax(nearest))
    return order


def fps_levels(points, hierarchy, start):
    """dict level -> sorted indices; nested prefixes of one FPS run plus the full set."""
    n = points.shape[0]
    if hierarchy[-1] >= n:
        return np.arange(n)                      # reference quirk Q2: bare array
    order = fps_order(points, hierarchy[-1], start)
    out = {lv: np.sort(order[:cnt]) for lv, cnt in enumerate(hierarchy)}
    out[len(hierarchy)] = np.arange(n)
    return out


def voxel_grid(extent, voxel_size):
    return np.ceil(extent / voxel_size).astype(int) + 1


def voxel_select(points, lo, voxel_size, dims):
    """One representative per occupied voxel (nearest to the voxel centre, first index
    on ties), returned in ascending voxel-id order."""
    cell = ((points - lo) / voxel_size).astype(int)
    cell = np.clip(cell, 0, dims - 1)
    vid = cell[:, 0] * dims[1] * dims[2] + cell[:, 1] * dims[2] + cell[:, 2]
    order = np.argsort(vid, kind="stable")
    vid_sorted = vid[order]
    starts = np.flatnonzero(np.r_[True, vid_sorted[1:] != vid_sorted[:-1]])
    ends = np.r_[starts[1:], vid_sorted.size]
    picks = np.empty(starts.size, dtype=np.int64)
    for s, (a, b) in enumerate(zip(starts, ends)):
        v = vid_sorted[a]
        c3 = np.array([v // (dims[1] * dims[2]), (v // dims[2]) % dims[1], v % dims[2]])
        centre = lo + (c3 + 0.5) * voxel_size
        members = order[a:b]                     # ascending point index (stable sort)
        picks[s] = members[np.argmin(_dist_to(points[members], centre))]
    return picks


def voxel_levels(points, hierarchy):
    points = np.asarray(points, dtype=np.float64)
    n = points.shape[0]
    lo, hi = points.min(axis=0), points.max(axis=0)
    extent = hi - lo
    out = {}
    for lv, target in enumerate(hierarchy):
        if target >= n:
            out[lv] = np.arange(n)
            continue
        base = (np.prod(extent) / (target * 2)) ** (1 / 3)
        best, best_gap = None, float("inf")
        for scale in VOXEL_SCALES:
            vs = base * scale
            picks = voxel_select(points, lo, vs, voxel_grid(extent, vs))
            gap = abs(len(picks) - target)
            if gap < best_gap:
                best_gap, best = gap, picks
            if len(picks) >= target * 0.95:
                break
        out[lv] = best[:target] if len(best) > target else best
    out[len(hierarchy)] = np.arange(n)
    return {lv: np.sort(v) for lv, v in out.items()}
