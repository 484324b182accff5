"""Vertex sharding and halo plans for the multi-GPU training step (host logic, numpy only).

Each resolution level is cut into contiguous vertex ranges, one per rank.  Contiguous ranges only have
small boundaries when the vertex order is local: structured grids are (row-major slabs), arbitrary meshes
should be permuted first (`z_order` + `permute_mesh`: latitude bands, two neighbours per rank).  A rank owns the rows of K, M, U and of
the corrector input in its range; to apply K and M it also needs the rows of U that its columns
reference outside the range - the halo.  The plan lists, per peer, which owned rows to send and
where received rows land in the local [owned | halo] ordering of U.
"""
import numpy as np
import scipy.sparse as sp


def interior_range(indptr, indices, n_own):
    """Largest contiguous run [a, b) of owned rows that reference NO halo column (column index >= n_own).  With a
    locality-preserving vertex order the boundary rows sit at the two ends of a rank's range, so one run covers almost
    everything; the rows outside it are processed after the halo exchange has completed."""
    row_has_halo = np.zeros(n_own, dtype=bool)
    if indices.size:
        rows = np.repeat(np.arange(n_own), np.diff(indptr[:n_own + 1]))
        row_has_halo[rows[indices[:indptr[n_own]] >= n_own]] = True
    if not row_has_halo.any():
        return 0, n_own
    # longest run of False
    padded = np.concatenate([[True], row_has_halo, [True]])
    edges = np.flatnonzero(padded)
    gaps = np.diff(edges) - 1
    i = int(np.argmax(gaps))
    a = int(edges[i])            # index in `padded` of the True before the run -> run starts at padded index a + 1 = row a
    return a, a + int(gaps[i])


def z_order(verts):
    """Vertex permutation for band partitioning: ascending z (stable).  Contiguous ranges of the permuted mesh are
    latitude bands, so every rank touches at most two neighbours and the halo is two rings of vertices."""
    return np.argsort(np.asarray(verts)[:, 2], kind="stable")


def permute_mesh(verts, tris, perm):
    """Relabel vertices: new vertex i = old vertex perm[i]."""
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    return np.asarray(verts)[perm], inv[np.asarray(tris)]


def split_ranges(n, world):
    """Contiguous, balanced [start, end) ranges."""
    base, rem = divmod(n, world)
    starts = [r * base + min(r, rem) for r in range(world + 1)]
    return [(starts[r], starts[r + 1]) for r in range(world)]


class LevelPlan:
    """Halo plan of one level for one rank.

    local K, M : (n_own) x (n_own + n_halo) CSR with columns remapped to [owned | halo]
    send[p]    : local row indices (int32) of owned rows that peer p needs, in p's halo order
    recv[p]    : (offset, count) slice of the halo block that peer p fills
    """

    def __init__(self, K, M, rank, world):
        K, M = sp.csr_matrix(K), sp.csr_matrix(M)
        n = K.shape[0]
        self.n_global, self.rank, self.world = n, rank, world
        self.ranges = split_ranges(n, world)
        lo, hi = self.ranges[rank]
        self.lo, self.hi, self.n_own = lo, hi, hi - lo
        Kl, Ml = K[lo:hi].tocsr(), M[lo:hi].tocsr()
        cols = np.union1d(Kl.indices, Ml.indices)
        halo = cols[(cols < lo) | (cols >= hi)]                      # sorted global ids
        self.halo_global = halo
        self.n_halo = halo.size
        owner = np.searchsorted(np.array([r[1] for r in self.ranges]), halo, side="right")
        self.recv = {}
        for p in range(world):
            sel = np.flatnonzero(owner == p)
            if sel.size:
                self.recv[p] = (int(sel[0]), int(sel.size))          # halo is sorted => contiguous per owner
        remap = np.full(n, -1, dtype=np.int64)
        remap[lo:hi] = np.arange(self.n_own)
        remap[halo] = self.n_own + np.arange(self.n_halo)

        def localise(A):
            B = sp.csr_matrix((A.data, remap[A.indices], A.indptr), shape=(self.n_own, self.n_own + self.n_halo))
            B.sort_indices()
            return B
        self.K_local, self.M_local = localise(Kl), localise(Ml)
        self.interior = interior_range(self.K_local.indptr, self.K_local.indices, self.n_own)
        self.send = {}                                               # filled by exchange_requests()

    def requests(self):
        """{peer: global ids this rank needs from peer} (ascending)."""
        return {p: self.halo_global[o:o + c] for p, (o, c) in self.recv.items()}

    def set_send_lists(self, wanted_by_peer):
        """wanted_by_peer: {peer: global ids that peer needs from this rank}."""
        self.send = {p: (np.asarray(ids) - self.lo).astype(np.int32) for p, ids in wanted_by_peer.items()
                     if len(ids)}
        for ids in self.send.values():
            assert ids.min() >= 0 and ids.max() < self.n_own


def build_plans(K, M, world):
    """All ranks' plans for one level with the send lists resolved (single-process helper used by the
    tests and by the emulated-partition check; a real run builds its own plan per rank and swaps the
    request lists with one all_to_all)."""
    plans = [LevelPlan(K, M, r, world) for r in range(world)]
    for r, pl in enumerate(plans):
        wanted = {p: plans[p].requests().get(r, np.zeros(0, dtype=np.int64)) for p in range(world) if p != r}
        pl.set_send_lists(wanted)
    return plans
