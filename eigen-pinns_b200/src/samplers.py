"""Hierarchy construction: down-samplers (farthest point / voxel / decimation) and the Sampler
that assembles per-level coordinates, operators, prolongations, initial subspaces and edge lists.

Drop-in for reference src/samplers.py (same function and attribute names).  The two hot
samplers run on the GPU (fp64 kernels, bit-identical index sets):
  _farthest_point_sampling  ->  one persistent cooperative kernel for all dependent iterations
  _voxel_downsampling       ->  voxel keys + atomic arg-min per voxel + ordered compaction
The FPS start vertex is drawn from an unseeded generator like the reference (:113-116) unless
`start` is given.
"""
import numpy as np

import _backend
import mesh_helpers
import utils
from Mesh import Mesh

_sampling = _backend.module("sampling")


def _voxel_downsampling(mesh, hierarchy):
    """dict level -> sorted vertex indices (one representative per occupied voxel, nearest the
    voxel centre), plus a final level holding every vertex."""
    return _sampling.voxel_levels(np.asarray(mesh.verts, dtype=np.float64), list(hierarchy))


def _farthest_point_sampling(mesh, hierarchy, start=None):
    """dict level -> sorted vertex indices: nested prefixes of one farthest-point ordering, plus a
    final level holding every vertex.  Returns a bare arange when hierarchy[-1] >= N (reference Q2)."""
    points = np.asarray(mesh.verts, dtype=np.float64)
    if start is None:
        start = int(np.random.default_rng().integers(0, points.shape[0]))
    return _sampling.fps_levels(points, list(hierarchy), int(start))


def _simplify_mesh_decimation(mesh, hierarchy):
    """Quadric decimation through pyvista/VTK (third-party, optional)."""
    try:
        import pyvista as pv
    except ImportError as exc:
        raise ImportError("mesh decimation needs pyvista (VTK); supply pre-coarsened meshes through "
                          "config.coarse_mesh_files instead") from exc
    faces = np.hstack([np.full((len(mesh.connectivity), 1), 3), mesh.connectivity]).ravel()
    surface = pv.PolyData(mesh.verts, faces)
    out = []
    for target in hierarchy:
        reduction = max(0.0, min(0.99, 1.0 - target / len(mesh.verts)))
        simp = surface.decimate(reduction, volume_preservation=True)
        out.append(Mesh(verts=simp.points, connectivity=simp.faces.reshape(-1, 4)[:, 1:4]))
    return out


class Sampler:
    def __init__(self, config):
        self.sampler_type = config.sampler_type
        self.edge_computation_type = config.edge_computation_type
        self.k_neighbors = config.k_neighbors
        self.prolongation_neighbors = config.prolongation_neighbors
        self.n_modes = config.n_modes
        self.hierarchy = config.hierarchy
        self.coarse_mesh_files = getattr(config, "coarse_mesh_files", None)
        self.fps_start = getattr(config, "fps_start", None)
        self.operator_type = getattr(config, "operator_type", "auto")
        if self.operator_type not in ('auto', 'point_cloud', 'fem'):
            raise ValueError(f"operator_type must be 'auto', 'point_cloud' or 'fem', got '{self.operator_type}'")
        self.meshes, self.X_list, self.K_list, self.M_list = [], [], [], []
        self.P_list, self.U_list, self.actual_hierarchy = [], [], []
        self.edge_index_list, self.indices_per_level = [], []
        if self.edge_computation_type != 'connectivity_based':
            self.edge_computation_type = 'knn_based'
        if self.sampler_type not in ['farthest_point', 'voxel_downsampling', 'graph_coarsening']:
            raise ValueError("sampler_type must be 'farthest_point', 'voxel_downsampling' or "
                             f"'graph_coarsening', got '{self.sampler_type}'")

    def _grid_coarsening(self, mesh, hierarchy):
        if self.sampler_type == 'farthest_point':
            return _farthest_point_sampling(mesh, hierarchy, self.fps_start)
        if self.sampler_type == 'voxel_downsampling':
            return _voxel_downsampling(mesh, hierarchy)

    def _coarse_meshes(self, mesh, hierarchy):
        try:
            return _simplify_mesh_decimation(mesh, hierarchy)
        except ImportError:
            if not self.coarse_mesh_files:
                raise
        # pre-coarsened stand-ins, mapped into the frame of the (already normalised) fine mesh
        raw = getattr(mesh, "raw_frame", None)
        out = []
        for path in self.coarse_mesh_files:
            cm = Mesh(path)
            verts = (cm.verts - raw[0]) / raw[1] if raw is not None else cm.verts
            out.append(Mesh(verts=verts, connectivity=cm.connectivity))
        return out

    def _push_level(self, X, K, M):
        self.X_list.append(X)
        self.K_list.append(K)
        self.M_list.append(M)
        self.actual_hierarchy.append(X.shape[0])

    def _assemble_X_K_M(self, mesh, hierarchy):
        if self.sampler_type == 'graph_coarsening':
            self.meshes = self._coarse_meshes(mesh, hierarchy) + [mesh]
            for m in self.meshes:
                K, M = mesh_helpers.compute_stiffness_and_mass_matrices(m)
                self._push_level(m.verts, K, M)
        else:
            self.indices_per_level = self._grid_coarsening(mesh, hierarchy)
            self.meshes.append(mesh)
            point_cloud = self._use_point_cloud_operators()
            for idx in self.indices_per_level.values():
                X = mesh.verts[idx]
                if point_cloud:
                    K, M = mesh_helpers.compute_laplacian_and_mass_matrices(X)
                else:
                    K, M = self._galerkin_operators(mesh, X)
                self._push_level(X, K, M)

    def _use_point_cloud_operators(self):
        """operator_type 'point_cloud': the reference's robust_laplacian operators (ImportError if the package is
        missing); 'fem': Galerkin restrictions of the mesh's FEM operators; 'auto': the former when robust_laplacian
        is installed, else the latter, so the default configuration runs without the third-party package."""
        if self.operator_type == 'fem':
            return False
        if self.operator_type == 'point_cloud':
            return True
        import importlib.util
        return importlib.util.find_spec("robust_laplacian") is not None

    def _galerkin_operators(self, mesh, X):
        """(P^T K P, P^T M P) for a sub-sampled level: K, M the FEM operators of the full mesh (Mesh.py:348-364), P
        the kNN inverse-distance interpolation from the level's points to the mesh vertices (utils.py:39-60).  P has
        unit row sums, so constants stay in the null space of the coarse stiffness matrix."""
        K, M = mesh_helpers.compute_stiffness_and_mass_matrices(mesh)
        if X.shape[0] == mesh.verts.shape[0]:
            return K, M
        P = utils.build_prolongation(X, mesh.verts, k=min(self.prolongation_neighbors, X.shape[0])).tocsr()
        return (P.T @ K.tocsr() @ P).tocoo(), (P.T @ M.tocsr() @ P).tocoo()

    def _assemble_edge_list(self):
        if self.sampler_type == 'graph_coarsening' and self.edge_computation_type == 'connectivity_based':
            self.edge_index_list = [mesh_helpers.mesh_to_edge_index(m) for m in self.meshes]
        else:
            self.edge_index_list = [utils.build_knn_graph(X, k=self.k_neighbors) for X in self.X_list]

    def _assemble_P_U(self):
        if self.sampler_type == 'graph_coarsening':
            _, U0, _, _ = utils.solve_eigenvalue_mesh(self.meshes[0], self.n_modes)
        elif self._use_point_cloud_operators():
            _, U0, _, _ = utils.solve_eigenvalue_point_cloud(self.X_list[0], self.n_modes)
        else:
            _, U0 = utils.solve_eigenvalue_operators(self.K_list[0], self.M_list[0], self.n_modes)
        self.U_list.append(U0)
        U_prev = U0.copy()
        for lv in range(1, len(self.X_list)):
            P = utils.build_prolongation(self.X_list[lv - 1], self.X_list[lv], k=self.prolongation_neighbors)
            self.P_list.append(P)
            U_lv = utils.jacobi_smooth(self.M_list[lv], self.K_list[lv], P @ U_prev, alpha=0.1, n_iters=10)
            self.U_list.append(U_lv)
            U_prev = U_lv.copy()

    def preprocess_mesh(self, mesh):
        self._assemble_X_K_M(mesh, self.hierarchy)
        self._assemble_edge_list()
        self._assemble_P_U()

    def visualize(self, output_prefix):
        raise NotImplementedError("plotting is outside the hot path (matplotlib is not a dependency)")
