"""One forward of the chain kernel (pair kernel unless EP_OLD=1) at the benchmark size - the target of an ncu capture."""
import ctypes, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
engine = importlib.import_module("eigen-pinns_b200.engine")
tcm = importlib.import_module("eigen-pinns_b200.mlp_tc")
cabi = importlib.import_module("eigen-pinns_b200._cabi")
n, k = int(sys.argv[1]) if len(sys.argv) > 1 else 998562, 32
dev = torch.device("cuda", 0)
if os.environ.get("EP_OLD"):
    cabi.call("ep_tune_set", 6, 1)
dims = [2 * (9 + k)] + [256] * 6 + [k]
g = torch.Generator().manual_seed(0)
Ws = [torch.randn(dims[i + 1], dims[i], generator=g) / np.sqrt(dims[i]) for i in range(len(dims) - 1)]
bs = [0.1 * torch.randn(dims[i + 1], generator=g) for i in range(len(dims) - 1)]
h = torch.randn(n, dims[0], device=dev); U = torch.randn(n, k, device=dev); up = torch.empty_like(U)
m = tcm.TcMlp(n, engine.FlatParams(Ws, bs, dev), dev, h, chain=True)
m.want_corr = False
for _ in range(3):
    m.forward(h, U, 0.5, up)
torch.cuda.synchronize()
print("ok")
