"""Corrector networks: one aggregation step, then a ReLU MLP on [x | agg].

Drop-in for reference src/corrector_model.py: same constructors, same `.net` (an
nn.Sequential of nn.Linear / nn.ReLU / nn.Dropout, so `state_dict` keys are `net.<i>.weight|bias`
and `model.net[-1].weight` works), same forward signatures.  The arithmetic runs in the
hand-written CUDA kernels:
  * SimpleCorrector   mean over incoming edges   -> ep_neighbor_mean_concat_f32  (:23-30)
  * SpectralCorrector agg = A_norm @ x           -> ep_spmm_concat_f32           (:76-79)
  * every Linear(+ReLU) pair                      -> ep_linear_fwd_f32 / ep_linear_bwd_f32
The graph structure and (for inputs that do not require grad) the concatenated input are cached
between calls, because the training loop feeds the same features every epoch.
"""
from typing import List

import torch
import torch.nn as nn

import _backend

_ops = _backend.module("ops")
_sparse = _backend.module("sparse")


def _build_net(in_dim, out_dim, hidden_layers, dropout):
    layers, prev = [], 2 * in_dim
    for width in hidden_layers:
        layers += [nn.Linear(prev, width), nn.ReLU(inplace=True)]
        if dropout > 0.0:
            layers.append(nn.Dropout(dropout))
        prev = width
    layers.append(nn.Linear(prev, out_dim))
    return nn.Sequential(*layers)


def _run_net(net, h, training):
    """Walk the Sequential, fusing each Linear with a directly following ReLU into one kernel."""
    mods = list(net)
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Linear):
            fuse = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
            h = _ops.linear(h, m.weight, m.bias, relu=fuse)
            i += 2 if fuse else 1
        elif isinstance(m, nn.ReLU):
            h = torch.relu(h)
            i += 1
        elif isinstance(m, nn.Dropout):
            h = torch.nn.functional.dropout(h, m.p, training)
            i += 1
        else:
            raise TypeError("unsupported layer in corrector net: %r" % (m,))
    return h


class _GraphCache:
    def __init__(self):
        self.key, self.csr, self.hkey, self.h = None, None, None, None

    @staticmethod
    def _tkey(t):
        return (t.data_ptr(), tuple(t.shape), t._version, str(t.device))


class SimpleCorrector(nn.Module):
    def __init__(self, in_dim, out_dim, hidden_layers, dropout):
        super().__init__()
        self.net = _build_net(in_dim, out_dim, hidden_layers, dropout)
        self._cache = _GraphCache()

    def corrector_input(self, x, edge_index):
        c = self._cache
        gkey = c._tkey(edge_index) + (x.shape[0],)
        if c.key != gkey:
            c.csr, c.key, c.hkey = _sparse.CsrMatrix.from_edge_index(edge_index, x.shape[0], x.device), gkey, None
        if x.requires_grad:
            raise NotImplementedError("gradients w.r.t. the node features are not part of the hot path")
        hkey = c._tkey(x)
        if c.hkey != hkey:
            c.h, c.hkey = _ops.neighbor_mean_concat(x, c.csr), hkey
        return c.h

    def forward(self, x, edge_index):
        return _run_net(self.net, self.corrector_input(x, edge_index), self.training)


class SpectralCorrector(nn.Module):
    def __init__(self, in_dim: int, out_dim: int, hidden_layers: List[int], dropout: float):
        super().__init__()
        self.net = _build_net(in_dim, out_dim, hidden_layers, dropout)
        self._cache = _GraphCache()

    def corrector_input(self, x, A_norm_sparse):
        c = self._cache
        if isinstance(A_norm_sparse, _sparse.CsrMatrix):
            csr, gkey = A_norm_sparse, id(A_norm_sparse)
        else:
            gkey = (id(A_norm_sparse), tuple(A_norm_sparse.shape))
            csr = c.csr if c.key == gkey else _sparse.CsrMatrix.from_torch_sparse(A_norm_sparse, x.device)
        if c.key != gkey:
            c.csr, c.key, c.hkey = csr, gkey, None
        hkey = c._tkey(x)
        if c.hkey != hkey:
            if x.requires_grad:
                raise NotImplementedError("gradients w.r.t. the node features are not part of the hot path")
            c.h, c.hkey = _ops.spmm_concat(x.detach(), c.csr), hkey
        return c.h

    def forward(self, x: torch.Tensor, A_norm_sparse) -> torch.Tensor:
        return _run_net(self.net, self.corrector_input(x, A_norm_sparse), self.training)
