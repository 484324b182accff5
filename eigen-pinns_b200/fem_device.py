"""Sparse FEM assembly on the GPU: triangles -> CSR stiffness / mass operators without a
host round trip.  Element blocks and the ordered segmented sum are our kernels (csrc/fem.cu); the stable key
sort and run detection between them use torch (library plumbing).  The result is bit-for-bit independent of
thread scheduling: duplicates are summed in triangle order, as the reference loop does (src/Mesh.py:355-361).
"""
import ctypes

import numpy as np
import torch

from ._cabi import call
from .sparse import CsrMatrix, OperatorPair


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def assemble(verts, tris, device="cuda", keep_fp64=False):
    """verts (N, 3) float64, tris (T, 3) int -> OperatorPair(K, M) on `device` (fp32 CSR, shared pattern);
    with keep_fp64 also returns (valK64, valM64) device tensors aligned with the CSR entries."""
    dev = torch.device(device)
    v = (verts if torch.is_tensor(verts) else torch.from_numpy(np.ascontiguousarray(verts, dtype=np.float64)))
    t = (tris if torch.is_tensor(tris) else torch.from_numpy(np.ascontiguousarray(tris).astype(np.int32)))
    v = v.to(device=dev, dtype=torch.float64).contiguous()
    t = t.to(device=dev, dtype=torch.int32).contiguous()
    n, T = v.shape[0], t.shape[0]
    k_el = torch.empty(9 * T, dtype=torch.float64, device=dev)
    m_el = torch.empty(9 * T, dtype=torch.float64, device=dev)
    keys = torch.empty(9 * T, dtype=torch.int64, device=dev)
    call("ep_fem_elements_f64", T, _p(v), _p(t), n, _p(k_el), _p(m_el), _p(keys), _stream())
    skeys, perm = torch.sort(keys, stable=True)
    uniq, counts = torch.unique_consecutive(skeys, return_counts=True)
    starts = torch.cumsum(counts, 0) - counts
    nnz = uniq.numel()
    col = torch.empty(nnz, dtype=torch.int32, device=dev)
    valK = torch.empty(nnz, dtype=torch.float32, device=dev)
    valM = torch.empty(nnz, dtype=torch.float32, device=dev)
    vK64 = torch.empty(nnz, dtype=torch.float64, device=dev) if keep_fp64 else None
    vM64 = torch.empty(nnz, dtype=torch.float64, device=dev) if keep_fp64 else None
    call("ep_fem_segment_sum_f64", nnz, _p(starts), _p(counts), _p(perm), _p(k_el), _p(m_el), _p(uniq), n, _p(col),
         _p(valK), _p(valM), _p(vK64), _p(vM64), _stream())
    rows = torch.div(uniq, n, rounding_mode="floor")
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    rowptr = rowptr.to(torch.int32)
    K = CsrMatrix.from_device_arrays(rowptr, col, valK, (n, n), symmetric=True)
    M = CsrMatrix.from_device_arrays(rowptr, col, valM, (n, n), symmetric=True)
    pair = OperatorPair(K, M, dev, assume_symmetric=True)
    return (pair, vK64, vM64) if keep_fp64 else pair
