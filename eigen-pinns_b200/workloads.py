"""Device-built benchmark workloads (BASELINE config 5: periodic torus grid, valence 6, FEM operators).

Everything is generated and assembled on the GPU (csrc/fem.cu); nothing of size N touches the host.
`build_torus_engine` covers one GPU and the vertex-sharded case: with `world > 1` every rank builds only its slab of
grid rows plus one halo row on either side, assembles the operator rows of its own vertices, and the halo plan
follows from the grid structure (the slab above / below).
"""
import numpy as np
import torch
import torch.nn as nn

from . import ops
from .engine import FlatParams, StepConfig, TrainStepEngine
from .fem_device import assemble
from .sparse import CsrMatrix, OperatorPair

R_MAJOR, R_MINOR = 1.0, 0.4


def _torus_points(rows, size, dev):
    """xyz of grid vertices (i, j) for i in `rows` (1-D long tensor), all j; plus the angles (u, w)."""
    i = rows.repeat_interleave(size)
    j = torch.arange(size, device=dev).repeat(rows.numel())
    u, w = 2 * np.pi * i.double() / size, 2 * np.pi * j.double() / size
    xyz = torch.stack([(R_MAJOR + R_MINOR * torch.cos(w)) * torch.cos(u),
                       (R_MAJOR + R_MINOR * torch.cos(w)) * torch.sin(u), R_MINOR * torch.sin(w)], 1)
    return xyz, u, w, i, j


def _torus_frame(size):
    """mean and max per-axis std of the full torus grid (closed form of mesh_helpers.normalize_mesh)."""
    w = 2 * np.pi * np.arange(size) / size
    rho = R_MAJOR + R_MINOR * np.cos(w)
    var_xy = 0.5 * np.mean(rho ** 2)                 # mean of rho^2 cos^2 u over a full period of u
    var_z = np.mean((R_MINOR * np.sin(w)) ** 2)
    return np.zeros(3), float(np.sqrt(max(var_xy, var_z))) + 1e-12


def _trial_modes(u, w, gid, k):
    """lowest torus harmonics plus a deterministic pseudo-random perturbation that depends only on the GLOBAL vertex
    id (so every rank generates identical values for shared vertices)."""
    cols = []
    order = sorted(((a * a + 6.25 * b * b, a, b) for a in range(0, 12) for b in range(0, 6)))
    for _, a, b in order:
        for fu in ((torch.cos, torch.sin) if a else (torch.cos,)):
            for fv in ((torch.cos, torch.sin) if b else (torch.cos,)):
                if len(cols) < k:
                    cols.append(fu(a * u) * fv(b * w))
    U = torch.stack(cols, 1)
    phase = (gid.double().unsqueeze(1) * 0.6180339887498949 + torch.arange(k, device=u.device).double() * 0.7548776662466927)
    noise = 0.05 * 1.7320508 * (2.0 * torch.frac(phase * 97.0 + torch.frac(phase) * 31.0) - 1.0)     # ~ var 0.05^2
    return (U + noise).float().contiguous()


class _GridPlan:
    """Halo plan of one torus slab; same attributes as partition.LevelPlan where the engine needs them."""

    def __init__(self, size, rank, world):
        from .partition import split_ranges
        self.rank, self.world, self.n_global = rank, world, size * size
        row_ranges = split_ranges(size, world)
        self.row_lo, self.row_hi = row_ranges[rank]
        self.lo, self.hi = self.row_lo * size, self.row_hi * size
        self.n_own = self.hi - self.lo
        below, above = (self.row_lo - 1) % size, self.row_hi % size
        halo = np.concatenate([below * size + np.arange(size), above * size + np.arange(size)])
        self.halo_global = np.unique(halo)
        self.n_halo = self.halo_global.size
        ends = np.array([r[1] * size for r in row_ranges])
        owner = np.searchsorted(ends, self.halo_global, side="right")
        self.recv = {}
        for p in range(world):
            sel = np.flatnonzero(owner == p)
            if sel.size:
                assert sel[-1] - sel[0] + 1 == sel.size
                self.recv[p] = (int(sel[0]), int(sel.size))
        self.send = {}
        # rows of the first and last owned grid row reference the halo; everything between is interior
        self.interior = (size, self.n_own - size) if self.n_own > 2 * size else (0, 0)

    def requests(self):
        return {p: self.halo_global[o:o + c] for p, (o, c) in self.recv.items()}

    def set_send_lists(self, wanted_by_peer):
        self.send = {p: (np.asarray(ids) - self.lo).astype(np.int32) for p, ids in wanted_by_peer.items() if len(ids)}


def build_torus_engine(size, k, dev, mlp_mode, hidden, rank=0, world=1, group=None):
    """Returns (engine, x_feats, adjacency, U_norm, n_global, nnz_global_estimate)."""
    import torch.distributed as dist
    from . import dist_engine
    n_global = size * size
    sharded = world > 1
    centre, scale = _torus_frame(size)
    if sharded:
        plan = dist_engine.resolve_send_lists(_GridPlan(size, rank, world), group)
        n_own, n_halo = plan.n_own, plan.n_halo
        rows_local = (torch.arange(plan.row_lo - 1, plan.row_hi + 1, device=dev) % size)        # below | owned | above
        own_slice = slice(size, size + n_own)
    else:
        plan, n_own, n_halo = None, n_global, 0
        rows_local = torch.arange(size, device=dev)
        own_slice = slice(0, n_global)
    xyz, u, w, gi, gj = _torus_points(rows_local, size, dev)
    xyz = (xyz - torch.tensor(centre, device=dev)) / scale
    gid = gi * size + gj
    n_loc = xyz.shape[0]
    # triangles of the quads whose lower grid row lies in the local block (all rows when not sharded: periodic)
    n_quad_rows = rows_local.numel() if not sharded else rows_local.numel() - 1
    li = torch.arange(n_quad_rows, device=dev).repeat_interleave(size)
    lj = torch.arange(size, device=dev).repeat(n_quad_rows)
    lip = (li + 1) % rows_local.numel() if not sharded else li + 1
    ljp = (lj + 1) % size
    v00, v10, v11, v01 = li * size + lj, lip * size + lj, lip * size + ljp, li * size + ljp
    tris = torch.cat([torch.stack([v00, v10, v11], 1), torch.stack([v00, v11, v01], 1)]).to(torch.int32)
    full = assemble(xyz, tris, dev)
    del tris, v00, v10, v11, v01, li, lj, lip, ljp
    if sharded:
        # keep the rows of owned vertices; columns -> [owned | halo sorted by global id]
        rp = full.K.rowptr.long()
        a, b = int(rp[size].item()), int(rp[size + n_own].item())
        halo_sorted = torch.from_numpy(plan.halo_global).to(dev)
        new_index = torch.empty(n_loc, dtype=torch.int64, device=dev)
        new_index[own_slice] = torch.arange(n_own, device=dev)
        is_halo = torch.ones(n_loc, dtype=torch.bool, device=dev)
        is_halo[own_slice] = False
        new_index[is_halo] = n_own + torch.searchsorted(halo_sorted, gid[is_halo])
        rowptr = (rp[size:size + n_own + 1] - a).to(torch.int32).contiguous()
        col = new_index[full.K.col[a:b].long()].to(torch.int32).contiguous()
        shape = (n_own, n_own + n_halo)
        K = CsrMatrix.from_device_arrays(rowptr, col, full.K.val[a:b].contiguous(), shape, symmetric=True)
        M = CsrMatrix.from_device_arrays(rowptr, col, full.M.val[a:b].contiguous(), shape, symmetric=True)
        pair = OperatorPair(K, M, dev, assume_symmetric=True)
        pair.n = n_own
        order = torch.empty(n_own + n_halo, dtype=torch.int64, device=dev)      # engine row -> local assembly row
        order[new_index] = torch.arange(n_loc, device=dev)
    else:
        pair, order = full, None

    def to_engine_rows(t):
        return t if order is None else t[order].contiguous()

    def allsum(t):
        if sharded:
            dist.all_reduce(t, group=group)
        return t

    def allmax(t):
        if sharded:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return t

    # adjacency (mesh edges) of the owned rows = off-diagonal pattern of K
    counts = (pair.K.rowptr[1:] - pair.K.rowptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(n_own, device=dev), counts)
    offd = pair.K.col.long() != rows
    diag_pos = (~offd).nonzero().squeeze(1)
    adj_counts = torch.bincount(rows[offd], minlength=n_own)
    adj_rowptr = torch.zeros(n_own + 1, dtype=torch.int64, device=dev)
    adj_rowptr[1:] = torch.cumsum(adj_counts, 0)
    adj = CsrMatrix.from_device_arrays(adj_rowptr.to(torch.int32), pair.K.col[offd].contiguous(), None,
                                       (n_own, n_own + n_halo))
    Kd, Md = pair.K.val[diag_pos].unsqueeze(1), pair.M.val[diag_pos].unsqueeze(1)
    del rows, offd, diag_pos
    # trial subspace on [owned | halo] rows, M-normalised with global column sums (reference :120-130)
    U0 = to_engine_rows(_trial_modes(u, w, gid, k))
    MU = ops.spmm(pair.M, U0)
    colsum = allsum((U0[:n_own].double() * MU.double()).sum(0))
    U_norm = (U0 / torch.sqrt(colsum + 1e-12).float().unsqueeze(0)).contiguous()
    del U0, MU
    KU, MU = ops.spmm2(pair, U_norm)
    Uo = U_norm[:n_own]
    A = allsum(ops.eigen_partials(Uo, KU, KU)[:k * k].clone()).view(k, k)      # U^T K U (Gram kernel)
    B = allsum(ops.eigen_partials(Uo, MU, MU)[:k * k].clone()).view(k, k)      # U^T M U
    from scipy.linalg import eigh
    lam = torch.from_numpy(eigh(A.cpu().numpy(), B.cpu().numpy(), eigvals_only=True).astype(np.float32)).to(dev)
    # node features of reference _compute_level_features (:159-201), single level
    deg = adj_counts.float().unsqueeze(1)
    deg = deg / (allmax(deg.max().reshape(1)) + 1e-12)
    rmag = torch.norm(KU - MU * lam.unsqueeze(0), dim=1, keepdim=True)
    rmag = rmag / (allmax(rmag.max().reshape(1)) + 1e-12)
    ray = (Uo * KU).sum(1, keepdim=True) / ((Uo * MU).sum(1, keepdim=True) + 1e-12)
    ray = ray / (lam.max() + 1e-12)
    coords = to_engine_rows(xyz.float())[:n_own]
    x_own = torch.cat([coords, torch.zeros(n_own, 1, device=dev), deg, Kd, Md, rmag, ray, Uo], 1)
    x_feats = torch.zeros((n_own + n_halo, x_own.shape[1]), dtype=torch.float32, device=dev)
    x_feats[:n_own] = x_own
    del KU, MU, rmag, ray, deg, x_own, xyz
    if sharded:
        halo = dist_engine.HaloExchanger(plan, dev, lambda r, idx, out: ops.gather_rows(r, idx, out=out), group)
        halo.exchange(x_feats, n_own)
    h_own = ops.neighbor_mean_concat(x_feats, adj)          # rows of owned vertices (adj has n_own rows)
    h = torch.zeros((n_own + n_halo, h_own.shape[1]), dtype=torch.float32, device=dev)
    h[:n_own] = h_own[:n_own]
    del h_own
    torch.manual_seed(0)                                   # identical initial weights on every rank
    dims = [h.shape[1]] + list(hidden) + [k]
    lins = [nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)]
    nn.init.normal_(lins[-1].weight, mean=0.0, std=0.01)
    nn.init.zeros_(lins[-1].bias)
    lins = [l.to(dev) for l in lins]
    params = FlatParams.adopt(lins)
    cfg = StepConfig()
    if sharded:
        eng = dist_engine.ShardedTrainStepEngine(h, U_norm, [plan], params, cfg, lam_target=lam, mlp_mode=mlp_mode,
                                                 group=group, pairs=[pair])
    else:
        eng = TrainStepEngine(h, U_norm, [pair], [0], params, cfg, lam_target=lam, mlp_mode=mlp_mode)
    eng._modules_keepalive = lins
    return eng, x_feats, adj, U_norm, n_global, 7 * n_global
