"""Dispatcher entries for the engine (SURVEY §8b "suggested extension ops"): ``torch.ops.b200gcn.*`` registered with
``torch.library.custom_op`` over the C-ABI launches, each with a fake (meta) implementation and an autograd formula,
so that the propagation is visible to ``torch.compile`` / ``torch.export`` as an opaque op instead of a Python call
into ctypes.  The ops take the CSR arrays of a resident :class:`GraphHandle` (``handle.csr()``) as plain tensors:

    y      = torch.ops.b200gcn.spmm(rowptr, col, val, x, n_src, symmetric)
    u, i   = torch.ops.b200gcn.lightgcn_propagate(rowptr, col, val, xu, xi, n_layers)        # symmetric graphs
    out    = torch.ops.b200gcn.bignn_tail(p, x, w1, b1, w2, b2, slope, normalize)

The kernels only exist for CUDA tensors (``device_types="cuda"``): there is no CPU implementation to dispatch to.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import functional as F_
from .graph import GraphHandle

_HANDLES: Dict[tuple, GraphHandle] = {}


def _handle(rowptr: Tensor, col: Tensor, val: Optional[Tensor], n_src: int, symmetric: bool) -> GraphHandle:
    """A GraphHandle view over CSR tensors handed to an op (cached: building one plans the hub rows, a sync)."""
    key = (rowptr.data_ptr(), col.data_ptr(), None if val is None else val.data_ptr(), int(n_src), bool(symmetric))
    h = _HANDLES.get(key)
    if h is None or h.csr()[0] is not rowptr:
        if len(_HANDLES) > 32:
            _HANDLES.clear()
        h = GraphHandle(rowptr=rowptr, col=col, value=val, sparse_sizes=(rowptr.numel() - 1, int(n_src)),
                        symmetric=bool(symmetric))
        _HANDLES[key] = h
    return h


@torch.library.custom_op("b200gcn::spmm", mutates_args=(), device_types="cuda")
def spmm(rowptr: Tensor, col: Tensor, val: Optional[Tensor], x: Tensor, n_src: int, symmetric: bool) -> Tensor:
    g = _handle(rowptr, col, val, n_src, symmetric)
    xx = F_._f32_rows(x, "x")
    y = torch.empty(g.size(0), xx.size(1), dtype=torch.float32, device=x.device)
    F_.spmm_raw(g, xx, y=y)
    return y


@spmm.register_fake
def _(rowptr, col, val, x, n_src, symmetric):
    return x.new_empty(rowptr.numel() - 1, x.size(1))


def _spmm_setup(ctx, inputs, output):
    rowptr, col, val, x, n_src, symmetric = inputs
    ctx.save_for_backward(rowptr, col) if val is None else ctx.save_for_backward(rowptr, col, val)
    ctx.n_src, ctx.symmetric, ctx.has_val = n_src, symmetric, val is not None


def _spmm_backward(ctx, gy):
    saved = ctx.saved_tensors
    rowptr, col = saved[0], saved[1]
    val = saved[2] if ctx.has_val else None
    gt = _handle(rowptr, col, val, ctx.n_src, ctx.symmetric).t()          # symmetric graphs return themselves
    rp_t, col_t, val_t = gt.csr()
    gx = torch.ops.b200gcn.spmm(rp_t, col_t, val_t, gy.contiguous(), rowptr.numel() - 1, ctx.symmetric)
    return None, None, None, gx, None, None


spmm.register_autograd(_spmm_backward, setup_context=_spmm_setup)


@torch.library.custom_op("b200gcn::lightgcn_propagate", mutates_args=(), device_types="cuda")
def lightgcn_propagate(rowptr: Tensor, col: Tensor, val: Optional[Tensor], xu: Tensor, xi: Tensor,
                       n_layers: int) -> Tuple[Tensor, Tensor]:
    g = _handle(rowptr, col, val, rowptr.numel() - 1, True)
    out = F_._propagate_layers(g, F_._f32_rows(xu, "user table"), F_._f32_rows(xi, "item table"), int(n_layers), True)
    U = xu.size(0)
    return out[:U].clone(), out[U:].clone()


@lightgcn_propagate.register_fake
def _(rowptr, col, val, xu, xi, n_layers):
    return xu.new_empty(xu.shape), xi.new_empty(xi.shape)


def _lgcn_setup(ctx, inputs, output):
    rowptr, col, val, xu, xi, n_layers = inputs
    ctx.save_for_backward(rowptr, col) if val is None else ctx.save_for_backward(rowptr, col, val)
    ctx.n_layers, ctx.has_val = n_layers, val is not None


def _lgcn_backward(ctx, gu, gi):
    saved = ctx.saved_tensors
    val = saved[2] if ctx.has_val else None
    # the layer-mean operator of a symmetric graph is symmetric: the backward is the same op on the gradients
    du, di = torch.ops.b200gcn.lightgcn_propagate(saved[0], saved[1], val, gu.contiguous(), gi.contiguous(), ctx.n_layers)
    return None, None, None, du, di, None


lightgcn_propagate.register_autograd(_lgcn_backward, setup_context=_lgcn_setup)


@torch.library.custom_op("b200gcn::bignn_tail", mutates_args=(), device_types="cuda")
def bignn_tail(p: Tensor, x: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, slope: float,
               normalize: bool) -> Tensor:
    return F_.bignn_tail(p, x, w1, b1, w2, b2, slope=slope, normalize=normalize)


@bignn_tail.register_fake
def _(p, x, w1, b1, w2, b2, slope, normalize):
    return x.new_empty(x.size(0), w1.size(0))
