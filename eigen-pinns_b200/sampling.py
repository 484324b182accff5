"""Host side of the two down-samplers (farthest point, voxel) on top of the fp64 CUDA kernels.

The dependent hot loops (reference src/samplers.py:119-127 and :58-74) run on the GPU;
the scalar control flow around them (voxel-size search, truncation, sorting of <= target
indices) is kept on the host in numpy exactly as the reference orders it, so results are
bit-identical index sets.
"""
import ctypes

import numpy as np
import torch

from ._cabi import call, query, EpError

VOXEL_SCALES = (0.7, 0.85, 1.0, 1.15, 1.3, 1.5)          # reference src/samplers.py:41


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _device_points(points, device):
    if torch.is_tensor(points):
        if not points.is_cuda:
            raise EpError("device point set expected (no CPU fallback exists)")
        return points.to(torch.float64).contiguous()
    pts = np.ascontiguousarray(points, dtype=np.float64)
    if pts.ndim != 2 or pts.shape[1] != 3:
        raise ValueError("points must be (N, 3)")
    return torch.from_numpy(pts).to(device)


def fps_order(points, n_samples, start, device="cuda"):
    """Selection order of farthest-point sampling as a device int64 tensor of length n_samples.
    order[0] = start.  Bit-exact w.r.t. reference src/samplers.py:116-127."""
    pts = _device_points(points, device)
    n = pts.shape[0]
    ws_bytes = query("ep_fps_workspace_bytes", n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=pts.device)
    out = torch.empty(n_samples, dtype=torch.int64, device=pts.device)
    call("ep_fps_f64", n, _p(pts), int(n_samples), int(start), _p(out), _p(ws), ws_bytes, _stream())
    return out


def fps_levels(points, hierarchy, start, device="cuda"):
    """dict level -> sorted index array, nested prefixes of one FPS run plus the full set
    (reference src/samplers.py:97-143; quirk Q2: a bare arange when hierarchy[-1] >= N)."""
    n = points.shape[0]
    if hierarchy[-1] >= n:
        return np.arange(n)
    order = fps_order(points, hierarchy[-1], start, device).cpu().numpy()
    out = {lv: np.sort(order[:cnt]) for lv, cnt in enumerate(hierarchy)}
    out[len(hierarchy)] = np.arange(n)
    return out


def bounds(points_dev):
    lo_hi = torch.empty(6, dtype=torch.float64, device=points_dev.device)
    call("ep_bounds_f64", points_dev.shape[0], _p(points_dev), _p(lo_hi), _stream())
    v = lo_hi.cpu().numpy()
    return v[:3].copy(), v[3:].copy()


_voxel_ws = {}


def voxel_select(points_dev, lo, voxel, dims, max_out=None):
    """One voxel pass: representative point per occupied voxel in ascending voxel id.
    Returns (count, indices[:min(count, max_out)]) as numpy.  The count and the picks come back in ONE device-to-host
    copy (one synchronisation per pass); the workspace is kept between passes."""
    n = points_dev.shape[0]
    dims64 = np.asarray(dims, dtype=np.int64)
    n_vox = int(dims64[0]) * int(dims64[1]) * int(dims64[2])
    max_out = int(max_out) if max_out is not None else min(n, n_vox)
    ws_bytes = query("ep_voxel_workspace_bytes", n, n_vox)
    key = str(points_dev.device)
    ws = _voxel_ws.get(key)
    if ws is None or ws.numel() < ws_bytes:
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=points_dev.device)
        _voxel_ws[key] = ws
    res = torch.zeros(1 + max(max_out, 1), dtype=torch.int64, device=points_dev.device)      # [count | picks]
    lo64 = np.ascontiguousarray(lo, dtype=np.float64)
    call("ep_voxel_select_f64", n, _p(points_dev), lo64.ctypes.data_as(ctypes.c_void_p), float(voxel),
         dims64.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(res.data_ptr() + 8), max_out, _p(res), _p(ws), ws_bytes,
         _stream())
    host = res.cpu().numpy()
    count = int(host[0])
    return count, host[1:1 + min(count, max_out)].copy()


def voxel_levels(points, hierarchy, device="cuda"):
    """Voxel-grid hierarchy, control flow of reference src/samplers.py:9-94."""
    pts_dev = _device_points(points, device)
    n = pts_dev.shape[0]
    lo, hi = bounds(pts_dev)
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
            dims = np.ceil(extent / vs).astype(int) + 1
            count, picks = voxel_select(pts_dev, lo, vs, dims, max_out=target)     # the reference keeps best[:target]
            gap = abs(count - target)
            if gap < best_gap:
                best_gap, best = gap, picks
            if count >= target * 0.95:
                break
        out[lv] = best[:target] if len(best) > target else best
    out = {lv: np.sort(v) for lv, v in out.items()}
    out[len(hierarchy)] = np.arange(n)                     # (already sorted: sorting 10^6 indices cost more than all passes)
    return out
