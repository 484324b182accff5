"""Triangle surface mesh: OBJ loading and linear-FEM operators.

Drop-in for the parts of reference src/Mesh.py the eigen-pinns pipeline uses:
`Mesh(filename)` / `Mesh(verts=, connectivity=)`, `.verts`, `.connectivity`, `.normals`,
`.centroids`, `.computeLaplacian()` (:348-364).  The reference assembles dense N x N arrays in a
Python loop over triangles; here the same element matrices (:180-198, :228-234) are evaluated
for all triangles at once and scattered to CSR (`computeLaplacianSparse`), and the dense pair is
only materialised on request for small meshes.  The fractal-tree utilities of the original
class (point projection, geodesics, Laplace solves, tvtk writer: :81-178, :239-346) are outside
the eigen-pinns hot path and are not provided.
"""
import numpy as np

import _backend

_fem = _backend.module("fem")
DENSE_LIMIT = 20000


class Mesh:
    def __init__(self, filename=None, verts=None, connectivity=None):
        if filename is not None:
            verts, connectivity = self.loadOBJ(filename)
        self.verts = np.array(verts, dtype=np.float64)
        self.connectivity = np.array(connectivity, dtype=np.int64)
        tri = self.connectivity
        if tri.size:
            a = self.verts[tri[:, 1]] - self.verts[tri[:, 0]]
            b = self.verts[tri[:, 2]] - self.verts[tri[:, 0]]
            n = np.cross(a, b)
            self.normals = n / np.linalg.norm(n, axis=1)[:, None]
            self.centroids = (self.verts[tri[:, 0]] + self.verts[tri[:, 1]] + self.verts[tri[:, 2]]) / 3.0
        else:
            self.normals = np.zeros((0, 3))
            self.centroids = np.zeros((0, 3))
        self._KM = None

    @staticmethod
    def loadOBJ(filename):
        """Vertices (`v x y z`) and faces (`f a/b/c ...`, 1-based, first index of each group)."""
        verts, faces = [], []
        with open(filename, "r") as handle:
            for line in handle:
                tok = line.split()
                if not tok:
                    continue
                if tok[0] == "v":
                    verts.append([float(t) for t in tok[1:4]])
                elif tok[0] == "f":
                    faces.append([int(t.split("/")[0]) - 1 for t in tok[1:]])
        return verts, faces

    def computeLaplacianSparse(self):
        """(K, M) as scipy CSR float64 with one shared sparsity pattern."""
        if self._KM is None:
            self._KM = _fem.assemble_stiffness_mass(self.verts, self.connectivity)
        return self._KM

    def computeLaplacian(self):
        """Dense (K, M) like the reference; refuses sizes where N x N float64 is unreasonable."""
        n = self.verts.shape[0]
        if n > DENSE_LIMIT:
            raise MemoryError("dense %d x %d operators requested; use computeLaplacianSparse()" % (n, n))
        K, M = self.computeLaplacianSparse()
        return K.toarray(), M.toarray()
