"""Drop-in graph-convolution layers with the reference's signatures
(recbole_gnn/model/layers.py:8-67): ``LightGCNConv``, ``BipartiteGCNConv``, ``BiGNNConv``.

``edge_index`` is polymorphic exactly as in the reference: either an int64 ``[2, nnz]`` tensor (row 0 =
source ids, row 1 = destination ids, PyG ``source_to_target`` flow) with a float ``edge_weight``, or a
sparse object (here a :class:`GraphHandle`) with ``edge_weight=None``.  Raw tensors are converted to a
resident CSR once and cached on (tensor identity, version, shape).  All arithmetic runs in libb200gcn;
CPU tensors raise.
"""
from __future__ import annotations

import weakref
from collections import OrderedDict
from typing import Optional, Tuple, Union

import torch
import torch.nn as nn

from . import _lib
from . import functional as F_
from .graph import GraphHandle

Tensor = torch.Tensor

_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_CACHE_CAP = 16


def handle_from_edges(edge_index: Tensor, edge_weight: Optional[Tensor], n_dst: int, n_src: int) -> GraphHandle:
    """Resident CSR for a raw ``(edge_index, edge_weight)`` pair, cached per tensor identity + version."""
    _lib.require_cuda(edge_index, edge_weight, what="edge_index/edge_weight")
    if edge_index.dim() != 2 or edge_index.size(0) != 2 or edge_index.dtype != torch.int64:
        raise TypeError("edge_index must be an int64 [2, nnz] tensor")
    key = (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape),
           None if edge_weight is None else (edge_weight.data_ptr(), edge_weight._version), n_dst, n_src,
           edge_index.device.index)
    hit = _CACHE.get(key)
    if hit is not None:
        ref_ei, ref_ew, handle = hit
        if ref_ei() is edge_index and (edge_weight is None or ref_ew() is edge_weight):
            _CACHE.move_to_end(key)
            return handle
    if edge_weight is not None and edge_weight.requires_grad:
        # the reference's message() would propagate dL/d edge_weight; the engine treats weights as constants
        raise NotImplementedError("edge_weight.requires_grad: the engine has no gradient w.r.t. edge weights "
                                  "(SURVEY §8a a12: weights are constants on this path); detach() it")
    w = None if edge_weight is None else edge_weight.float().contiguous()
    # A[dst, src] = w  ->  row = destination, col = source
    handle = GraphHandle(row=edge_index[1].contiguous(), col=edge_index[0].contiguous(), value=w,
                         sparse_sizes=(n_dst, n_src)).to(edge_index.device)
    _CACHE[key] = (weakref.ref(edge_index), None if edge_weight is None else weakref.ref(edge_weight), handle)
    # per-epoch / per-step augmented graphs (dropout_adj, SGL-style raw edge_index): drop the resident CSR (and its
    # transpose / hub scratch) as soon as the tensor it was built from dies, not when 16 newer graphs arrived
    weakref.finalize(edge_index, _CACHE.pop, key, None)
    while len(_CACHE) > _CACHE_CAP:
        _CACHE.popitem(last=False)
    return handle


def _resolve(edge_index, edge_weight, n_dst: int, n_src: int) -> GraphHandle:
    if isinstance(edge_index, GraphHandle):
        if not edge_index.is_resident:
            raise RuntimeError("GraphHandle is not resident; call .to(device) as GeneralGraphRecommender does")
        if edge_index.sparse_sizes() != (n_dst, n_src):
            raise ValueError(f"graph is {edge_index.sparse_sizes()}, expected {(n_dst, n_src)}")
        return edge_index
    return handle_from_edges(edge_index, edge_weight, n_dst, n_src)


class _Propagating(nn.Module):
    """The part of PyG's ``MessagePassing`` surface the reference's layers define or call
    (``propagate`` / ``message`` / ``message_and_aggregate``, layers.py:14-20,32-35,55-64), over the engine:
    ``propagate`` IS the fused gather-reduce (no ``[nnz, D]`` message tensor exists), ``message`` is kept for
    subclasses / callers that evaluate it directly."""

    def __init__(self):
        super().__init__()
        self.aggr = "add"

    def propagate(self, edge_index, x, edge_weight=None, size=None):
        if isinstance(x, (tuple, list)):
            x_src, n_dst = x[0], (int(size[1]) if size is not None else x[1].size(0))
        else:
            x_src, n_dst = x, (int(size[1]) if size is not None else x.size(0))
        g = _resolve(edge_index, edge_weight, n_dst, x_src.size(0))
        return F_.spmm(g, x_src)

    def message(self, x_j, edge_weight):
        return edge_weight.view(-1, 1) * x_j

    def message_and_aggregate(self, adj_t, x):
        if not isinstance(adj_t, GraphHandle):
            raise TypeError("message_and_aggregate takes the engine's sparse object (GraphHandle)")
        return F_.spmm(adj_t, x[0] if isinstance(x, (tuple, list)) else x)


class LightGCNConv(_Propagating):
    """``out = A_hat x`` (layers.py:8-23)."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, x: Tensor, edge_index: Union[Tensor, GraphHandle], edge_weight: Optional[Tensor]) -> Tensor:
        return self.propagate(edge_index, x=x, edge_weight=edge_weight)

    def __repr__(self):
        return '{}({})'.format(self.__class__.__name__, self.dim)


class BipartiteGCNConv(_Propagating):
    """Rectangular propagation (layers.py:26-38): ``x = (x_src, x_dst)``, ``size = (n_src, n_dst)``; only
    ``x_src`` is read; ``edge_index[0]`` holds source ids, ``edge_index[1]`` destination ids."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, x: Union[Tensor, Tuple[Tensor, Tensor]], edge_index, edge_weight, size: Tuple[int, int]) -> Tensor:
        x_src = x[0] if isinstance(x, (tuple, list)) else x
        n_src, n_dst = int(size[0]), int(size[1])
        if x_src.size(0) != n_src:
            raise ValueError(f"x_src has {x_src.size(0)} rows, size[0] = {n_src}")
        g = _resolve(edge_index, edge_weight, n_dst, n_src)
        return F_.spmm(g, x_src)

    def __repr__(self):
        return '{}({})'.format(self.__class__.__name__, self.dim)


class BiGNNConv(_Propagating):
    r"""NGCF layer (layers.py:41-67):  output = (L+I) E W_1 + (L E) \otimes E W_2."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin1 = torch.nn.Linear(in_features=in_channels, out_features=out_channels)
        self.lin2 = torch.nn.Linear(in_features=in_channels, out_features=out_channels)

    def forward(self, x: Tensor, edge_index, edge_weight) -> Tensor:
        g = _resolve(edge_index, edge_weight, x.size(0), x.size(0))
        x_prop = F_.spmm(g, x)
        needs_grad = torch.is_grad_enabled() and (
            x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if self.in_channels % 4 or self.out_channels % 4 or max(self.in_channels, self.out_channels) > 256:
            # shapes outside the fused tail: dense torch ops; the SpMM above is the engine's kernel either way
            x_trans = self.lin1(x_prop + x)
            x_inter = self.lin2(torch.mul(x_prop, x))
            return x_trans + x_inter
        if needs_grad:
            # fused forward (slope 1, no normalise: the layer returns the pre-activation), hand-written backward
            return F_.bignn_tail_autograd(x_prop, x, self.lin1.weight, self.lin1.bias, self.lin2.weight,
                                          self.lin2.bias, slope=1.0, normalize=False)
        return F_.bignn_tail(x_prop, x, self.lin1.weight, self.lin1.bias, self.lin2.weight, self.lin2.bias,
                             activate=False, normalize=False,
                             pre_out=torch.empty(x.size(0), self.out_channels, dtype=torch.float32, device=x.device))

    def __repr__(self):
        return '{}({},{})'.format(self.__class__.__name__, self.in_channels, self.out_channels)
