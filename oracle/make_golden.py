"""Run the UNMODIFIED reference (imported from /root/reference/src) on seeded inputs
and write the golden fixtures under tests/golden/.  ORACLE — test infrastructure only.

    python oracle/make_golden.py            # only works where /root/reference exists

Everything the GPU-side parity tests need travels in the .npz files (mesh arrays,
inputs, reference outputs); nothing at test time reads /root/reference.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import reference_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def csr_arrays(A, prefix):
    A = A.tocsr()
    A.sort_indices()
    return {prefix + "_indptr": A.indptr.astype(np.int32), prefix + "_indices": A.indices.astype(np.int32),
            prefix + "_data": A.data.astype(np.float64)}


def make_config(ref, **over):
    cfg = ref.config.PINNConfig.from_yaml(os.path.join(reference_loader.REFERENCE_SRC, "parameters.yml"))
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


def main():
    ref = reference_loader.load()
    os.makedirs(OUT, exist_ok=True)
    R = reference_loader.REFERENCE_ROOT
    quiet = contextlib.redirect_stdout(io.StringIO())

    # ---------------------------------------------------------------- meshes + FEM
    bunny = ref.mesh_helpers.load_mesh(os.path.join(R, "resources/bunny.obj"), normalize=True)
    Kd, Md = bunny.computeLaplacian()                       # dense reference operators
    from scipy.linalg import eigh
    from scipy.sparse import coo_matrix, csr_matrix
    evals, evecs = eigh(Kd, Md, subset_by_index=[0, 9])
    K_b, M_b = coo_matrix(Kd), coo_matrix(Md)
    # a second, coarser level in bunny's frame (FEM on resources/coarse_3.obj)
    raw = ref.Mesh.Mesh(os.path.join(R, "resources/bunny.obj"))
    c3 = ref.Mesh.Mesh(os.path.join(R, "resources/coarse_3.obj"))
    centroid, std_max = raw.verts.mean(0), raw.verts.std(0).max() + 1e-12
    c3n = ref.Mesh.Mesh(verts=(c3.verts - centroid) / std_max, connectivity=c3.connectivity)
    Kc_d, Mc_d = c3n.computeLaplacian()
    K_c, M_c = coo_matrix(Kc_d), coo_matrix(Mc_d)
    fem = dict(verts=bunny.verts, tris=bunny.connectivity.astype(np.int32), eig10=evals,
               evec10=evecs, coarse_verts=c3n.verts, coarse_tris=c3n.connectivity.astype(np.int32),
               K_rowsum_abs=np.abs(Kd).sum(1), M_rowsum=Md.sum(1),
               Kc_rowsum_abs=np.abs(Kc_d).sum(1), Mc_rowsum=Mc_d.sum(1))
    fem.update(csr_arrays(K_b, "K"))
    fem.update(csr_arrays(M_b, "M"))
    np.savez_compressed(os.path.join(OUT, "bunny_fem.npz"), **fem)

    # ---------------------------------------------------------------- eigen-loss (H5/H7)
    loss = {}
    for tag, k, levels in (("k16_1lvl", 16, [(K_b, M_b)]), ("k64_1lvl", 64, [(K_b, M_b)]),
                           ("k16_2lvl", 16, [(K_c, M_c), (K_b, M_b)])):
        cfg = make_config(ref, n_modes=k)
        gnn = ref.multigrid_model.MultigridGNN(cfg)
        gnn.device = torch.device("cpu")
        torch.manual_seed(100 + k + len(levels))
        n_tot = sum(Kx.shape[0] for Kx, _ in levels)
        U = (0.3 * torch.randn(n_tot, k)).requires_grad_(True)
        offs = gnn._compute_node_offsets([np.zeros((Kx.shape[0], 3)) for Kx, _ in levels])
        Ks, Ms = [a for a, _ in levels], [b for _, b in levels]
        l_res, l_orth, lams = gnn._compute_residual_ortho_loss(U, Ks, Ms, offs, 1000.0, 10.0, k)
        lam_t = torch.linspace(0.0, 2.0, k)
        extra = gnn._compute_eigenvalue_losses(Ms, lam_t, lams, 0.0, 0.5, 2.0, 3.0)
        total = l_res + l_orth + sum(extra)
        total.backward()
        loss.update({f"{tag}_U": U.detach().numpy(), f"{tag}_loss_res": l_res.item(),
                     f"{tag}_loss_orth": l_orth.item(), f"{tag}_total": total.item(),
                     f"{tag}_extra": np.array([e.item() for e in extra]),
                     f"{tag}_lam_target": lam_t.numpy(),
                     f"{tag}_grad": U.grad.numpy(), f"{tag}_offsets": np.array(offs, dtype=np.int64)})
        for i, l in enumerate(lams):
            loss[f"{tag}_lam{i}"] = l.detach().numpy()
    np.savez_compressed(os.path.join(OUT, "eigen_loss.npz"), **loss)

    # ---------------------------------------------------------------- corrector + features + short training run
    k = 16
    hidden = [64, 64, 64]
    cfg = make_config(ref, n_modes=k, hidden_layers=hidden, epochs=6, log_every=100)
    sampler = types.SimpleNamespace()
    sampler.X_list = [c3n.verts, bunny.verts]
    sampler.K_list, sampler.M_list = [K_c, K_b], [M_c, M_b]
    with quiet:
        sampler.edge_index_list = [ref.utils.build_knn_graph(X, k=8) for X in sampler.X_list]
    # initial subspaces: exact low modes of each level + noise (no CGC: reference's dense coarse solve is
    # singular on these meshes) -> start from _normalize_eigenvectors onward
    rng = np.random.default_rng(7)
    U_lists, lam_list = [], []
    for Kx, Mx in zip([Kc_d, Kd], [Mc_d, Md]):
        w, V = eigh(Kx, Mx, subset_by_index=[0, k - 1])
        U_lists.append((V + 0.05 * rng.standard_normal(V.shape)).astype(np.float32))
    corr = {}
    for model_type in ("simple", "spectral"):
        cfg.model_type = model_type
        gnn = ref.multigrid_model.MultigridGNN(cfg)
        gnn.device = torch.device("cpu")
        lam_list = []
        for U0, Kx, Mx in zip(U_lists, sampler.K_list, sampler.M_list):
            vals, _ = gnn.refine_eigenvectors(U0, Kx, Mx)
            lam_list.append(torch.FloatTensor(vals))
        U_norm = gnn._normalize_eigenvectors([torch.from_numpy(u) for u in U_lists], sampler.M_list)
        with quiet:
            x_feats, ei, A_norm = gnn._build_features(sampler.X_list, U_norm, lam_list,
                                                      sampler.edge_index_list, sampler.K_list, sampler.M_list)
            torch.manual_seed(11)
            gnn._initialize_model(x_feats.shape[1], k, hidden, 0.0)
        state0 = {n: p.detach().clone().numpy() for n, p in gnn.model.state_dict().items()}
        out0 = gnn._forward_pass(x_feats, ei, A_norm).detach().numpy()
        opt, sched = gnn._create_optimizer(cfg.learning_rate, cfg.weight_decay)
        U_all = torch.cat(U_norm, dim=0)
        offs = gnn._compute_node_offsets(sampler.X_list)
        # the reference ramps the correction scale as epoch/5000; six epochs would barely move, so the
        # fixture drives the reference's own loop body at epochs 2500..2505 (scale = 5.0...)
        losses = []
        for epoch in range(2500, 2506):
            gnn.model.train()
            opt.zero_grad()
            corr_raw = gnn._forward_pass(x_feats, ei, A_norm)
            scale = gnn.corr_scale * min(1.0, epoch / 5000.0)
            U_pred = U_all + scale * corr_raw
            l_res, l_orth, lam_pred = gnn._compute_residual_ortho_loss(
                U_pred, sampler.K_list, sampler.M_list, offs, gnn.w_res, gnn.w_orth, k)
            ex = gnn._compute_eigenvalue_losses(sampler.M_list, lam_list[0], lam_pred, gnn.w_proj,
                                                gnn.w_trace, gnn.w_order, gnn.w_eigen)
            total = l_res + l_orth + sum(ex)
            total.backward()
            torch.nn.utils.clip_grad_norm_(gnn.model.parameters(), gnn.grad_clip)
            opt.step()
            sched.step(total.item())
            losses.append([total.item(), l_res.item(), l_orth.item()])
        state1 = {n: p.detach().clone().numpy() for n, p in gnn.model.state_dict().items()}
        with quiet:
            U_final = gnn._generate_final_predictions(x_feats, ei, A_norm, U_all, U_norm, sampler.M_list)
        t = model_type
        corr[f"{t}_x_feats"] = x_feats.numpy()
        corr[f"{t}_out0"] = out0
        corr[f"{t}_losses"] = np.array(losses)
        corr[f"{t}_U_final"] = U_final
        for n_, v in state0.items():
            corr[f"{t}_init_{n_}"] = v
        for n_, v in state1.items():
            corr[f"{t}_after_{n_}"] = v
        if A_norm is not None:
            A = A_norm.coalesce()
            corr["A_norm_indices"] = A.indices().numpy()
            corr["A_norm_values"] = A.values().numpy()
    corr["edge_index_0"] = sampler.edge_index_list[0].numpy()
    corr["edge_index_1"] = sampler.edge_index_list[1].numpy()
    corr["edge_index_all"] = ei.numpy()
    corr["U0_0"], corr["U0_1"] = U_lists
    corr["U_norm_0"], corr["U_norm_1"] = [u.numpy() for u in U_norm]
    corr["lam_0"], corr["lam_1"] = [l.numpy() for l in lam_list]
    corr["hidden"] = np.array(hidden)
    vals_rr, U_rr = gnn.refine_eigenvectors(U_lists[1], K_b, M_b)
    corr["rr_vals"], corr["rr_U"] = vals_rr, U_rr
    np.savez_compressed(os.path.join(OUT, "corrector_train.npz"), **corr)

    # ---------------------------------------------------------------- samplers (S1/S2)
    samp = {}
    real_rng = np.random.default_rng

    def run_fps(points, hierarchy, seed=42):
        ref.samplers.np.random.default_rng = lambda: real_rng(seed)     # reference start is unseeded (Q1)
        try:
            out = ref.samplers._farthest_point_sampling(types.SimpleNamespace(verts=points), hierarchy)
        finally:
            ref.samplers.np.random.default_rng = real_rng
        start = int(real_rng(seed).integers(0, points.shape[0]))
        return out, start

    sys.path.insert(0, os.path.join(ROOT, "eigen-pinns_b200"))
    import synthetic
    ico, _ = synthetic.icosphere(8)                                       # tie-heavy symmetric set
    cloud = real_rng(1234).standard_normal((6000, 3))
    grid = np.stack(np.meshgrid(*[np.arange(9.0)] * 3, indexing="ij"), -1).reshape(-1, 3)   # exact ties
    for tag, pts, hier in (("bunny", bunny.verts, [256, 512, 1024]), ("ico8", ico, [16, 64, 200]),
                           ("cloud", cloud, [32, 128, 700]), ("grid", grid, [8, 27, 300])):
        out, start = run_fps(pts, hier)
        samp[f"fps_{tag}_pts"], samp[f"fps_{tag}_hier"], samp[f"fps_{tag}_start"] = pts, np.array(hier), start
        for lv, idx in out.items():
            samp[f"fps_{tag}_lv{lv}"] = idx
    for tag, pts, hier in (("bunny", bunny.verts, [256, 512, 1024]), ("cloud", cloud, [50, 400, 2000]),
                           ("grid", grid, [8, 100]), ("ico8", ico, [16, 100])):
        out = ref.samplers._voxel_downsampling(types.SimpleNamespace(verts=pts), hier)
        samp[f"vox_{tag}_pts"], samp[f"vox_{tag}_hier"] = pts, np.array(hier)
        for lv, idx in out.items():
            samp[f"vox_{tag}_lv{lv}"] = idx
    np.savez_compressed(os.path.join(OUT, "samplers.npz"), **samp)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
