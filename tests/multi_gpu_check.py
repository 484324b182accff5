"""torchrun entry: the vertex-sharded engine on WORLD_SIZE GPUs must reproduce the single-GPU engine
(same losses step by step, same weights afterwards).  Launched by tests/test_gpu_multi.py."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eigen-pinns_b200", "src"))


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    if mode.startswith("torus"):
        return torus_check(mode, rank, world, dev)
    label = mode
    cabi_ok = True
    if mode == "cabi":
        # the multi-GPU entry points of the C ABI (ep_halo_exchange_f32, ep_allreduce_sum_*) on torch's communicator:
        # all-reduces against torch.distributed, then the whole sharded step with the C-ABI halo exchange
        cabi_ok = cabi_allreduce_check(rank, world, dev)
        os.environ["EP_HALO_CABI"] = "1"
        mode = "fp32"
    import bench
    import config as cfg_mod
    import multigrid_model
    de = importlib.import_module("eigen-pinns_b200.dist_engine")
    w = bench.build_host_workload("icosphere10k", band_order=True)
    k = w["k"]
    cfg = cfg_mod.PINNConfig.from_yaml(os.path.join(ROOT, "eigen-pinns_b200", "src", "parameters.yml"))
    cfg.n_modes, cfg.mlp_mode, cfg.seed, cfg.hidden_layers = k, mode, 0, [256, 256]
    sys.stdout = open(os.devnull, "w")
    results = []
    for sharded in (False, True):
        gnn = multigrid_model.MultigridGNN(cfg)
        edges = torch.from_numpy(w["edges"])
        U_norm = gnn._normalize_eigenvectors([w["U0"]], [w["M"]])
        lam0 = torch.linspace(0, 1, k)
        x_feats, edge_all, A_norm = gnn._build_features([w["verts"]], U_norm, [lam0], [edges], [w["K"]], [w["M"]])
        gnn._initialize_model(x_feats.shape[1], k, cfg.hidden_layers, 0.0)
        opt, _ = gnn._create_optimizer(gnn.lr, gnn.weight_decay)
        if sharded:
            eng = de.make_sharded_engine(gnn, x_feats, edge_all, U_norm[0], w["K"], w["M"], lam0, opt, rank, world)
            if label == "cabi":
                cabi_ok = cabi_ok and type(eng.halo[0]).__name__ == "CabiHaloExchanger"
        else:
            eng = gnn._make_engine(x_feats, edge_all, A_norm, U_norm[0], [w["K"]], [w["M"]], lam0, [0], opt)
        losses = [eng.step(2500 + i).cpu().numpy().copy() for i in range(2)]
        if sharded and len(sys.argv) > 2 and sys.argv[2] == "graph":
            eng.enable_graph()
        losses += [eng.step(2502 + i).cpu().numpy().copy() for i in range(2)]
        results.append((np.array(losses), eng.params.flat.clone(), [l.clone() for l in eng.lams]))
    sys.stdout = sys.__stdout__
    (l1, p1, lam1), (l2, p2, lam2) = results
    tol = 1e-4 if mode == "fp32" else 2e-3
    ok = np.allclose(l1, l2, rtol=tol, atol=1e-9)
    perr = (p1 - p2).abs().max().item()
    ok = ok and perr <= (2e-5 if mode == "fp32" else 2e-3)
    ok = ok and torch.allclose(lam1[0], lam2[0], rtol=tol, atol=1e-6) and cabi_ok
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "OK" if flag.item() == 1.0 else "FAIL", "world", world, "mode", label,
              "loss_single", l1[:, 5].tolist(), "loss_sharded", l2[:, 5].tolist(), "param_err", perr)
    _leave(flag.item() == 1.0)


def cabi_allreduce_check(rank, world, dev):
    import ctypes
    cabi = importlib.import_module("eigen-pinns_b200._cabi")
    de = importlib.import_module("eigen-pinns_b200.dist_engine")
    warm = torch.ones(1, device=dev)
    dist.all_reduce(warm)                                   # creates the communicator
    comm = de.nccl_comm_ptr()
    if comm is None or cabi.query("ep_dist_nccl_version") == 0:
        return False
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    ok = True
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for dtype, entry in ((torch.float64, "ep_allreduce_sum_f64"), (torch.float32, "ep_allreduce_sum_f32")):
        for count in (1, 1109, 362_529):
            a = torch.randn(count, device=dev, dtype=dtype, generator=g)
            b = a.clone()
            dist.all_reduce(a)
            cabi.call(entry, ctypes.c_void_p(comm), count, ctypes.c_void_p(b.data_ptr()), st)
            torch.cuda.synchronize()
            ok = ok and torch.allclose(a, b, rtol=1e-12 if dtype == torch.float64 else 1e-5, atol=1e-12 if dtype == torch.float64 else 1e-6)
    return ok


def _leave(ok):
    """Tearing the NCCL communicator down while captured CUDA graphs still reference it can block for minutes
    (torch 2.11 / NCCL 2.28): synchronise, flush and leave without the orderly shutdown."""
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0 if ok else 1)


def torus_check(mode, rank, world, dev):
    """Device-built torus: the sharded engine (per-rank slab assembly + grid halo plan) against the single-GPU engine."""
    wl = importlib.import_module("eigen-pinns_b200.workloads")
    mlp_mode = "bf16" if mode.endswith("bf16") else "fp32"
    size, k, hidden = 192, 32, [128, 128]
    res = []
    for w_, r_ in ((1, 0), (world, rank)):
        eng = wl.build_torus_engine(size, k, dev, mlp_mode, hidden, r_, w_)[0]
        res.append((np.array([eng.step(2500 + i).cpu().numpy().copy() for i in range(4)]), eng.params.flat.clone(),
                    eng.lams[0].clone()))
    (l1, p1, lam1), (l2, p2, lam2) = res
    tol = 1e-4 if mlp_mode == "fp32" else 3e-3
    perr = (p1 - p2).abs().max().item()
    ok = np.allclose(l1, l2, rtol=tol, atol=1e-9) and perr <= 1e-4 + (2e-3 if mlp_mode == "bf16" else 0.0)
    ok = ok and torch.allclose(lam1, lam2, rtol=tol, atol=1e-6)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "OK" if flag.item() == 1.0 else "FAIL", "torus world", world, mlp_mode,
              "loss_single", l1[:, 5].tolist(), "loss_sharded", l2[:, 5].tolist(), "param_err", perr)
    _leave(flag.item() == 1.0)


if __name__ == "__main__":
    main()
