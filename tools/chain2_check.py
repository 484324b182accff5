"""Compares the CTA-pair chain kernel (tc_chain2_kernel, tcgen05 cta_group::2) with the single-CTA chain kernel:
bit-exact outputs (U_pred, every stored activation, every ReLU mask, every dZ) and CUDA-event timings.
    python tools/chain2_check.py [sizes...]      # ep_tune_set key 6: 0 = CTA pairs (default), 1 = single CTA"""
import ctypes
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
engine = importlib.import_module("eigen-pinns_b200.engine")
tcm = importlib.import_module("eigen-pinns_b200.mlp_tc")
cabi = importlib.import_module("eigen-pinns_b200._cabi")


def run(n, k, which, reps):
    dev = torch.device("cuda", 0)
    cabi.call("ep_tune_set", 6, which)
    dims = [2 * (9 + k)] + [256] * 6 + [k]
    g = torch.Generator().manual_seed(0)
    Ws = [torch.randn(dims[i + 1], dims[i], generator=g) / np.sqrt(dims[i]) for i in range(len(dims) - 1)]
    bs = [0.1 * torch.randn(dims[i + 1], generator=g) for i in range(len(dims) - 1)]
    g2 = torch.Generator(device=dev).manual_seed(1)
    h = torch.randn(n, dims[0], device=dev, generator=g2)
    U = torch.randn(n, k, device=dev, generator=g2)
    d_out = torch.randn(n, k, device=dev, generator=g2) / n
    up = torch.empty_like(U)
    p = engine.FlatParams(Ws, bs, dev)
    m = tcm.TcMlp(n, p, dev, h, chain=True)
    m.want_corr = True
    for _ in range(2):
        m.forward(h, U, 0.5, up)
        m.backward(h, d_out)
    torch.cuda.synchronize()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    e0, e1, e2 = ev(), ev(), ev()
    tf = tb = 0.0
    for _ in range(reps):
        e0.record()
        m.forward(h, U, 0.5, up)
        e1.record()
        m.backward(h, d_out)
        e2.record()
        torch.cuda.synchronize()
        tf += e0.elapsed_time(e1)
        tb += e1.elapsed_time(e2)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    e0.record()
    for _ in range(reps):
        cabi.call("ep_tc_chain_dx_bf16", m.n, m.L - 1, m._t_bpd, P(m.dz_out), m._t_WT, m._t_bmasks, m._t_dzs, st)
    e1.record()
    torch.cuda.synchronize()
    tdx = e0.elapsed_time(e1) / reps
    e0.record()
    for _ in range(reps):
        cabi.call("ep_tc_chain_fwd_bf16", m.n, m.L, m._t_pd, m._t_out, P(m.x0), m._t_Wp, m._t_b, m._t_acts, m._t_masks,
                  None, k, P(U), 0.5, None, P(up), k, st)
    e1.record()
    torch.cuda.synchronize()
    tfk = e0.elapsed_time(e1) / reps
    if which == 0 and os.environ.get("EP_TRACE"):
        trace = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
        ptr = trace.data_ptr()
        lo = ptr & 0xFFFFFFFF
        cabi.call("ep_tune_set", 3, ctypes.c_int(lo if lo < 2**31 else lo - 2**32))
        cabi.call("ep_tune_set", 4, ctypes.c_int(ptr >> 32))
        m.forward(h, U, 0.5, up)
        torch.cuda.synchronize()
        cabi.call("ep_tune_set", 3, 0)
        cabi.call("ep_tune_set", 4, 0)
        tr = trace.view(64, 8).cpu().numpy()
        t0 = tr[0, 0]
        print("step (layer, slot): ready issued acc_full drained arrived  (clk since first)")
        for gi in range(30):
            r = tr[gi] - t0
            print("%2d (l%d u%d): %7d %7d %7d %7d %7d   issue %5d tail %5d epi %5d arrive %4d" % (
                gi, (gi // 2) % m.L, gi % 2, r[0], r[1], r[2], r[3], r[4], r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3]))
    outs = [up.clone(), m.corr.clone()] + [a.clone() for a in m.acts] + [a.clone() for a in m.masks] + \
           [a.clone() for a in m.dzs] + [m.p.grad.clone() if hasattr(m.p, "grad") and m.p.grad is not None else torch.zeros(1)]
    cabi.call("ep_tune_set", 6, 0)
    return outs, (tf / reps, tb / reps, tdx, tfk)


def main():
    sizes = [int(s) for s in sys.argv[1:]] or [1000, 33000, 998562]
    if os.environ.get("EP_FLAGS"):
        cabi.call("ep_tune_set", 8, int(os.environ["EP_FLAGS"]))
    for n in sizes:
        for k in (32, 64) if n < 100000 else (32,):
            reps = 3 if n < 100000 else 10
            new, t_new = run(n, k, 0, reps)
            old, t_old = run(n, k, 1, reps)
            bad = [i for i, (x, y) in enumerate(zip(new, old)) if not torch.equal(x, y)]
            print("n=%d k=%d  pair kernel fwd %.3f bwd %.3f dx %.3f fwd-kernel %.3f ms | single-CTA fwd %.3f bwd %.3f dx %.3f fwd-kernel %.3f ms | %s"
                  % (n, k, *t_new, *t_old, "BIT-EXACT" if not bad else "MISMATCH in outputs %s" % bad), flush=True)
            if bad:
                i = bad[0]
                d = (new[i].float() - old[i].float()).abs()
                print("   first mismatch: output %d, max abs diff %g, count %d of %d" % (i, d.max().item(), (d > 0).sum().item(), d.numel()))


if __name__ == "__main__":
    main()
