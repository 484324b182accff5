"""Command-line entry of the B200 build: mesh -> hierarchy -> corrector training -> refined eigenvectors -> .vtu.

Same role and call order as the reference's src/main.py:9-35 (run it from a directory that holds ./resources):

    python <this dir>/main.py [parameters.yml]

Without an argument the YAML next to this file is used, unless ./src/parameters.yml exists in the working
directory (the reference's location).
"""
import os
import sys

import diagnostics
import mesh_helpers
import samplers
from config import PINNConfig
from multigrid_model import MultigridGNN

HERE = os.path.dirname(os.path.abspath(__file__))


def pick_config_file(explicit=None):
    if explicit:
        return explicit
    reference_style = os.path.join(".", "src", "parameters.yml")
    return reference_style if os.path.exists(reference_style) else os.path.join(HERE, "parameters.yml")


def load_normalised_mesh(path):
    """The normalised mesh plus the frame (centroid, max axis std) that maps raw coordinates into it - pre-coarsened
    stand-in meshes (config.coarse_mesh_files) live in the raw frame."""
    raw = mesh_helpers.load_mesh(path, normalize=False)
    mesh = mesh_helpers.normalize_mesh(raw)
    mesh.raw_frame = (raw.verts.mean(0), raw.verts.std(0).max() + 1e-12)
    return mesh


def main(config_file=None):
    config = PINNConfig.from_yaml(pick_config_file(config_file))
    print("Loading mesh...")
    mesh = load_normalised_mesh(config.mesh_file)
    print("Preprocessing mesh data...")
    sampler = samplers.Sampler(config)
    sampler.preprocess_mesh(mesh)
    print("Training physics-informed multiresolution GNN...")
    U_refined = MultigridGNN(config).train_multiresolution(sampler)
    print("Saving predicted eigenvectors...")
    out_dir = os.path.dirname(config.vtu_file)
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
    mesh_helpers.save_eigenfunctions(mesh, U_refined, config.n_modes, config.vtu_file)
    print("Running comprehensive diagnostics...")
    main.last_report = diagnostics.comprehensive_diagnostics(U_refined, mesh, sampler, config)
    return U_refined


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
