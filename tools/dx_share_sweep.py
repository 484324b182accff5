"""MLP backward time (CUDA events) against the share of SMs given to the dX kernel of each concurrent dX || dW pair."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
engine = importlib.import_module("eigen-pinns_b200.engine")
tcm = importlib.import_module("eigen-pinns_b200.mlp_tc")
n, k = 998562, 32
dev = torch.device("cuda", 0)
dims = [2 * (9 + k)] + [256] * 6 + [k]
g = torch.Generator().manual_seed(0)
Ws = [torch.randn(dims[i + 1], dims[i], generator=g) / np.sqrt(dims[i]) for i in range(len(dims) - 1)]
bs = [0.1 * torch.randn(dims[i + 1], generator=g) for i in range(len(dims) - 1)]
h = torch.randn(n, dims[0], device=dev); U = torch.randn(n, k, device=dev); up = torch.empty_like(U)
d_out = torch.randn(n, k, device=dev) / n
p = engine.FlatParams(Ws, bs, dev)
m = tcm.TcMlp(n, p, dev, h)
m.want_corr = False
m.forward(h, U, 0.5, up)
ref = None
for share in [float(x) for x in (sys.argv[1:] or ["0.5", "0.55", "0.6", "0.65", "0.7", "0.5"])]:
    m.dx_share = share
    for _ in range(3):
        m.backward(h, d_out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        m.backward(h, d_out)
    e1.record()
    torch.cuda.synchronize()
    gr = p.grad.clone()
    ref = gr if ref is None else ref
    print("dx share %.2f: backward %.3f ms   grads equal to first run: %s" % (share, e0.elapsed_time(e1) / 20, torch.equal(gr, ref)), flush=True)
