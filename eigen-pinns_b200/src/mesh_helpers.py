"""Mesh helpers: loading / normalisation, operator wrappers, edge lists, VTU output.

Drop-in for reference src/mesh_helpers.py (same function names and argument meaning).
Differences: operators come back sparse (never densified), the edge list is built vectorised,
`.vtu` files are written with the standard library only (meshio is not required), and the
point-cloud Laplacian is delegated to `robust_laplacian` only if that third-party package is
installed (it is outside the hot path; see DESIGN.md "out of scope").
"""
import base64
import struct
import zlib

import numpy as np
from scipy.sparse import coo_matrix

import _backend
from Mesh import Mesh

_fem = _backend.module("fem")


def normalize_mesh(mesh):
    return Mesh(verts=_fem.normalize_verts(mesh.verts), connectivity=mesh.connectivity)


def load_mesh(file, normalize=True):
    mesh = Mesh(file)
    return normalize_mesh(mesh) if normalize else mesh


def compute_stiffness_and_mass_matrices(mesh):
    K, M = mesh.computeLaplacianSparse()
    return coo_matrix(K), coo_matrix(M)


def compute_laplacian_and_mass_matrices(point_cloud):
    try:
        import robust_laplacian
    except ImportError as exc:
        raise ImportError("the point-cloud Laplacian needs the third-party package robust_laplacian, which is "
                          "not installed; use sampler_type 'graph_coarsening' (FEM operators) instead") from exc
    return robust_laplacian.point_cloud_laplacian(point_cloud)


def mesh_to_edge_index(mesh):
    import torch
    return torch.from_numpy(_fem.connectivity_edges(mesh.connectivity))


def _vtu_array(arr):
    """One zlib block, VTK 'vtkZLibDataCompressor' header with UInt32 sizes, base64."""
    raw = np.ascontiguousarray(arr).tobytes()
    comp = zlib.compress(raw)
    head = struct.pack("<IIII", 1, len(raw), len(raw), len(comp))
    return (base64.b64encode(head) + base64.b64encode(comp)).decode("ascii")


def save_eigenfunctions(mesh, U_pred, n_modes, vtu_file):
    """Write vertices (normalised frame), triangles and point data v0..v{n_modes-1} (Float64) as a
    zlib-compressed binary .vtu - the layout meshio produces for the reference (outputs/bunny_model.vtu)."""
    centroid = mesh.verts.mean(0)
    verts = (mesh.verts - centroid) / mesh.verts.std(0).max()
    tris = np.asarray(mesh.connectivity, dtype=np.int64)
    n_pts, n_cells = verts.shape[0], tris.shape[0]
    offsets = (np.arange(n_cells, dtype=np.int64) + 1) * 3
    types = np.full(n_cells, 5, dtype=np.int64)
    parts = ['<?xml version="1.0"?>',
             '<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian" '
             'compressor="vtkZLibDataCompressor">', "<UnstructuredGrid>",
             '<Piece NumberOfPoints="%d" NumberOfCells="%d">' % (n_pts, n_cells),
             '<Points><DataArray type="Float64" Name="Points" NumberOfComponents="3" format="binary">',
             _vtu_array(verts.astype(np.float64)), "</DataArray></Points>", "<Cells>",
             '<DataArray type="Int64" Name="connectivity" format="binary">', _vtu_array(tris), "</DataArray>",
             '<DataArray type="Int64" Name="offsets" format="binary">', _vtu_array(offsets), "</DataArray>",
             '<DataArray type="Int64" Name="types" format="binary">', _vtu_array(types), "</DataArray>",
             "</Cells>", "<PointData>"]
    for i in range(n_modes):
        parts += ['<DataArray type="Float64" Name="v%d" format="binary">' % i,
                  _vtu_array(np.asarray(U_pred[:, i], dtype=np.float64)), "</DataArray>"]
    parts += ["</PointData>", "</Piece>", "</UnstructuredGrid>", "</VTKFile>"]
    with open(vtu_file, "w") as handle:
        handle.write("\n".join(parts))


def read_vtu_point_data(vtu_file):
    """Minimal reader for the files written above / by meshio (zlib, UInt32 or UInt64 headers)."""
    import re
    text = open(vtu_file, "r").read()
    head_t = np.uint64 if 'header_type="UInt64"' in text else np.uint32
    out = {}
    for m in re.finditer(r'<DataArray type="(\w+)" Name="([^"]+)"[^>]*format="binary"[^>]*>\s*([^<]+?)\s*</DataArray>',
                         text):
        dtype, name, payload = m.group(1), m.group(2), m.group(3).strip()
        hs = np.dtype(head_t).itemsize
        nb = int(np.frombuffer(base64.b64decode(payload[:_b64len(hs)])[:hs], dtype=head_t)[0])
        head_len = _b64len(hs * (3 + nb))
        head = np.frombuffer(base64.b64decode(payload[:head_len]), dtype=head_t)
        data = base64.b64decode(payload[head_len:])
        chunks, pos = [], 0
        for c in head[3:3 + nb]:
            chunks.append(zlib.decompress(data[pos:pos + int(c)]))
            pos += int(c)
        out[name] = np.frombuffer(b"".join(chunks), dtype={"Float64": np.float64, "Float32": np.float32,
                                                            "Int64": np.int64, "Int32": np.int32,
                                                            "UInt8": np.uint8}[dtype])
    return out


def _b64len(n_bytes):
    return 4 * ((n_bytes + 2) // 3)


def meshio_to_Mesh(meshio_mesh, normalize=True):
    mesh = Mesh(verts=meshio_mesh.points, connectivity=meshio_mesh.cells_dict['triangle'])
    return normalize_mesh(mesh) if normalize else mesh
