"""Trained-accuracy evidence (BASELINE "lambda rel err"): the same multiresolution problem trained for E epochs by
  * the CPU oracle (oracle/step_port.py, the reference's arithmetic)        --oracle
  * the B200 engine in fp32 parity mode and in bf16 tensor-core mode        --gpu
from identical initial weights, then evaluated the way the reference does (src/multigrid_model.py:359-408, :452-475):
final prediction with the full correction scale, per-level M-normalisation, Rayleigh-Ritz on the finest level,
eigenvalues against the exact generalised eigenvalues of the FEM operators (fixture bunny_fem.npz, eig via eigsh).

Problem: coarse FEM level (1057 vertices) + bunny (2503 vertices), k = 16, MLP 50 -> 256 x 6 -> 16 (the reference's
default width and depth), kNN-8 aggregation graph, prolongated + Jacobi-smoothed coarse eigenvectors as U_base
(the reference CGC is singular on these meshes, SURVEY Q12, so it is skipped identically on both sides).

    python tests/trained_accuracy_run.py --oracle --epochs 3000 --out profiles/r02_trained_accuracy_oracle.json
    python tests/trained_accuracy_run.py --gpu    --epochs 3000 --out profiles/r02_trained_accuracy_gpu.json
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
HIDDEN = [256] * 6


def problem(k=16):
    from conftest import load_golden, csr_from_golden
    from oracle import step_port
    from scipy.sparse.linalg import eigsh
    fem_mod = importlib.import_module("eigen-pinns_b200.fem")
    fem = load_golden("bunny_fem.npz")
    n = fem["verts"].shape[0]
    K, M = csr_from_golden(fem, "K", n), csr_from_golden(fem, "M", n)
    Kc, Mc = fem_mod.assemble_stiffness_mass(fem["coarse_verts"], fem["coarse_tris"])
    from sklearn.neighbors import NearestNeighbors

    def knn_graph(X, kk):
        _, idx = NearestNeighbors(n_neighbors=kk + 1).fit(X).kneighbors(X)
        rows = np.repeat(np.arange(X.shape[0], dtype=np.int64), kk)
        return torch.from_numpy(np.stack([rows, idx[:, 1:].astype(np.int64).ravel()]))

    def prolongation(Xc, Xf, kk):
        dist, idx = NearestNeighbors(n_neighbors=kk).fit(Xc).kneighbors(Xf)
        w = 1.0 / (dist + 1e-12)
        w /= w.sum(1, keepdims=True)
        return sp.coo_matrix((w.ravel(), (np.repeat(np.arange(Xf.shape[0]), kk), idx.ravel())),
                             shape=(Xf.shape[0], Xc.shape[0]))
    Xs = [fem["coarse_verts"], fem["verts"]]
    vals_c, U0 = eigsh(Kc.tocsc(), k=k, M=Mc.tocsc(), sigma=-1e-6, which="LM")
    U0 = U0[:, np.argsort(vals_c)]
    P = prolongation(Xs[0], Xs[1], 8)
    A = (M + 0.1 * K).tocsr()
    U1 = P @ U0
    d_inv = 1.0 / (M.diagonal() + 0.1 * K.diagonal() + 1e-12)
    rhs = M @ U1
    for _ in range(10):                                          # utils.jacobi_smooth (reference utils.py:220-232)
        U1 = U1 + d_inv[:, None] * (rhs - A @ U1)
    exact, _ = eigsh(K.tocsc(), k=k, M=M.tocsc(), sigma=-1e-6, which="LM")
    exact = np.sort(exact)
    Ks, Ms = [Kc.tocoo(), K.tocoo()], [Mc.tocoo(), M.tocoo()]
    U_norm = [step_port.m_normalize(torch.from_numpy(U.astype(np.float32)), Mm) for U, Mm in zip([U0, U1], Ms)]
    eis = [knn_graph(X, 8) for X in Xs]
    lams = [torch.from_numpy(step_port.rayleigh_ritz(U.numpy(), Kk, Mm)[0].astype(np.float32))
            for U, Kk, Mm in zip(U_norm, Ks, Ms)]
    feats = [step_port.level_features(X, U, lam, ei, Kk, Mm, i, 2)
             for i, (X, U, lam, ei, Kk, Mm) in enumerate(zip(Xs, U_norm, lams, eis, Ks, Ms))]
    x = torch.cat(feats, 0)
    ei_all = torch.cat(eis, 1)                                   # un-offset like the reference
    return dict(x=x, ei=ei_all, U_base=torch.cat(U_norm, 0), Ks=Ks, Ms=Ms, lam0=lams[0], exact=exact, k=k,
                n_fine=n, n_coarse=Xs[0].shape[0], U_norm=U_norm)


def evaluate(pb, corr, corr_scale=10.0):
    """reference :359-384 + :452-475 on the finest level, in fp64 on the host."""
    from oracle import step_port
    U_pred = pb["U_base"] + corr_scale * corr
    nc = pb["n_coarse"]
    Uf = step_port.m_normalize(U_pred[nc:].contiguous(), pb["Ms"][1])
    vals, U_ref = step_port.rayleigh_ritz(Uf.numpy(), pb["Ks"][1], pb["Ms"][1])
    exact = pb["exact"]
    rel = np.abs(vals[1:] - exact[1:]) / np.abs(exact[1:])
    return {"ritz_values": vals.tolist(), "exact": exact.tolist(), "lambda_rel_err_mean": float(rel.mean()),
            "lambda_rel_err_max": float(rel.max()), "lambda_rel_err_first10_max": float(rel[:9].max())}


def run_oracle(pb, epochs, log):
    from oracle import step_port
    tr = step_port.CorrectorTrainer(pb["x"], pb["ei"], pb["U_base"], pb["Ks"], pb["Ms"], pb["lam0"], HIDDEN, pb["k"])
    w0 = ([w.detach().clone() for w in tr.weights], [b.detach().clone() for b in tr.biases])
    base = evaluate(pb, torch.zeros_like(pb["U_base"]))
    hist = []
    t0 = time.time()
    for e in range(epochs):
        hist.append(tr.step()[0])
        if e % log == 0:
            print("oracle epoch %5d loss %.6f  (%.1f s)" % (e, hist[-1], time.time() - t0), flush=True)
    with torch.no_grad():
        corr = tr.forward()
    return w0, dict(loss=hist[::max(1, epochs // 200)], final_loss=hist[-1], untrained=base, trained=evaluate(pb, corr),
                    seconds=time.time() - t0)


def run_gpu(pb, epochs, mlp_mode, w0, log):
    ops, sparse, engine = (importlib.import_module("eigen-pinns_b200." + m) for m in ("ops", "sparse", "engine"))
    dev = torch.device("cuda", 0)
    n_tot = pb["x"].shape[0]
    adj = sparse.CsrMatrix.from_edge_index(pb["ei"], n_tot, dev)
    h = ops.neighbor_mean_concat(pb["x"].to(dev), adj)
    params = engine.FlatParams(w0[0], w0[1], dev)
    pairs = [sparse.OperatorPair(K, M, dev) for K, M in zip(pb["Ks"], pb["Ms"])]
    eng = engine.TrainStepEngine(h, pb["U_base"].to(dev), pairs, [0, pb["n_coarse"]], params, engine.StepConfig(),
                                 lam_target=pb["lam0"].to(dev), mlp_mode=mlp_mode)
    reader = engine.LossReader(depth=4)
    hist, pend = [], []
    torch.cuda.synchronize()
    t0 = time.time()
    lr, best, stale = 1e-3, float("inf"), 0
    for e in range(epochs):
        if e == 3:
            eng.enable_graph()
        pend.append(reader.push(eng.step(e, lr=lr)))
        if len(pend) > 1:
            hist.append(float(reader.get(pend.pop(0))[5]))
            if e % log == 0:
                print("%s epoch %5d loss %.6f" % (mlp_mode, e, hist[-1]), flush=True)
    hist.append(float(reader.get(pend.pop(0))[5]))
    torch.cuda.synchronize()
    secs = time.time() - t0
    eng.enable_graph(False)
    corr = eng.mlp.forward(eng.h)
    if mlp_mode == "bf16":
        eng.mlp.want_corr = True
        corr = eng.mlp.forward(eng.h)
    corr = corr.detach().cpu()
    return dict(loss=hist[::max(1, epochs // 200)], final_loss=hist[-1], trained=evaluate(pb, corr), seconds=secs,
                steps_per_s=epochs / secs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=3000)
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--log", type=int, default=500)
    args = ap.parse_args()
    pb = problem()
    from oracle import step_port
    tr0 = step_port.CorrectorTrainer(pb["x"], pb["ei"], pb["U_base"], pb["Ks"], pb["Ms"], pb["lam0"], HIDDEN, pb["k"])
    w0 = ([w.detach().clone() for w in tr0.weights], [b.detach().clone() for b in tr0.biases])
    out = {"epochs": args.epochs, "vertices": [pb["n_coarse"], pb["n_fine"]], "k": pb["k"], "hidden": HIDDEN,
           "untrained": evaluate(pb, torch.zeros_like(pb["U_base"]))}
    if args.oracle:
        _, out["oracle_cpu_fp32"] = run_oracle(pb, args.epochs, args.log)
    if args.gpu:
        for mode in ("fp32", "bf16"):
            out["b200_" + mode] = run_gpu(pb, args.epochs, mode, w0, args.log)
    txt = json.dumps(out, indent=1)
    if args.out:
        open(args.out, "w").write(txt)
    print(json.dumps({k_: (v["trained"] if isinstance(v, dict) and "trained" in v else None) for k_, v in out.items()
                      if isinstance(v, dict)}, indent=1))


if __name__ == "__main__":
    main()
