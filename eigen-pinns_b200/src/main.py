"""Entry point: `python <this dir>/main.py` from a directory that holds ./resources (as the
reference's src/main.py:9-35 expects).  Loads the YAML next to this file unless
./src/parameters.yml exists in the working directory."""
import os
import sys

import numpy as np

import samplers
import mesh_helpers
from multigrid_model import MultigridGNN
from config import PINNConfig


def main(config_file=None):
    if config_file is None:
        here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "parameters.yml")
        config_file = './src/parameters.yml' if os.path.exists('./src/parameters.yml') else here
    config = PINNConfig.from_yaml(config_file)

    print("Loading mesh...")
    raw = mesh_helpers.load_mesh(config.mesh_file, normalize=False)
    mesh = mesh_helpers.normalize_mesh(raw)
    mesh.raw_frame = (raw.verts.mean(0), raw.verts.std(0).max() + 1e-12)

    print("Preprocessing mesh data...")
    sampler = samplers.Sampler(config)
    sampler.preprocess_mesh(mesh)

    print("Training physics-informed multiresolution GNN...")
    solver = MultigridGNN(config)
    U_refined = solver.train_multiresolution(sampler)

    print("Saving predicted eigenvectors...")
    os.makedirs(os.path.dirname(config.vtu_file) or ".", exist_ok=True)
    mesh_helpers.save_eigenfunctions(mesh, U_refined, config.n_modes, config.vtu_file)
    return U_refined


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
