"""The fused training step around the propagation: what one iteration of RecBole's
``Trainer._train_epoch`` does for LightGCN — ``calculate_loss`` (lightgcn.py:83-110), ``loss.backward()`` and
``optimizer.step()`` (torch.optim.Adam) — as engine launches with no autograd graph:

    forward   K fused SpMM launches (layer mean in the epilogue)            functional._propagate_layers
    loss      ONE kernel: BPR rows gathered, both scores, loss terms, and the gradient rows scattered straight into
              the (zeroed) gradient table of the propagation output           b200gcn_bpr_loss
    backward  the same K launches on A^T applied to that gradient table (the layer-mean operator is linear)
    reg       EmbLoss gradient rows added to the ego gradient (second, rows-only call of the loss kernel)
    update    one Adam kernel per embedding table                             b200gcn_adam_step

and its row-sharded twin (``ShardedLightGCNTrainer``): every rank owns the rows / optimiser state of its users and
items, forward and backward are one chain-kernel launch each (sharded.ShardedPropagator), the batch-sized row exchange
is one small all-reduce, and the Adam update touches only the rank's own rows.  Both produce the loss curve of
``model.calculate_loss`` + ``torch.optim.Adam`` (tests/test_train_gpu.py).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from . import functional as F_

Tensor = torch.Tensor


def bpr_loss_fused(u_all: Tensor, i_all: Tensor, reg_u: Tensor, reg_i: Tensor, user: Tensor, pos: Tensor,
                   neg: Tensor, *, reg_weight: float, require_pow: bool = False, gamma: float = 1e-10,
                   g_u_all: Optional[Tensor] = None, g_i_all: Optional[Tensor] = None,
                   g_reg_u: Optional[Tensor] = None, g_reg_i: Optional[Tensor] = None) -> Tensor:
    """``b200gcn_bpr_loss``: returns the 5-float stats tensor ``[loss, mf, ||U_b||, ||P_b||, ||N_b||]`` (on the
    device, no host sync) and ACCUMULATES the gradients into the given tables (None pairs are skipped)."""
    _lib.require_cuda(u_all, i_all, reg_u, reg_i, user, pos, neg, g_u_all, g_i_all, g_reg_u, g_reg_i, what="bpr operand")
    for t in (user, pos, neg):
        if t.dtype != torch.int64 or t.dim() != 1 or not t.is_contiguous():
            raise TypeError("user/pos/neg must be contiguous 1-D int64 id tensors")
    B, D = user.numel(), u_all.size(1)
    dev = u_all.device
    out = torch.empty(5, dtype=torch.float32, device=dev)
    lib = _lib.load()
    need = C.c_size_t(0)
    _lib.check(lib.b200gcn_bpr_loss_workspace(B, C.byref(need)))
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
    ld = F_._ld
    with torch.cuda.device(dev):
        _lib.check(lib.b200gcn_bpr_loss(
            u_all.data_ptr(), ld(u_all), i_all.data_ptr(), ld(i_all), reg_u.data_ptr(), ld(reg_u), reg_i.data_ptr(),
            ld(reg_i), user.data_ptr(), pos.data_ptr(), neg.data_ptr(), B, D, float(gamma), float(reg_weight),
            int(bool(require_pow)), _lib.ptr(g_u_all), ld(g_u_all) if g_u_all is not None else 0, _lib.ptr(g_i_all),
            ld(g_i_all) if g_i_all is not None else 0, _lib.ptr(g_reg_u), ld(g_reg_u) if g_reg_u is not None else 0,
            _lib.ptr(g_reg_i), ld(g_reg_i) if g_reg_i is not None else 0, out.data_ptr(), ws.data_ptr(), need.value,
            _lib.stream_ptr(dev)))
    return out


def adam_step(param: Tensor, grad: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, *, lr: float, step: int,
              betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0) -> None:
    """``b200gcn_adam_step``: torch.optim.Adam's update of one contiguous fp32 tensor, in place."""
    _lib.require_cuda(param, grad, exp_avg, exp_avg_sq, what="adam operand")
    for t in (param, grad, exp_avg, exp_avg_sq):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != param.numel():
            raise ValueError("adam_step takes contiguous fp32 tensors of one size")
    with torch.cuda.device(param.device):
        _lib.check(_lib.load().b200gcn_adam_step(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(),
                                                 exp_avg_sq.data_ptr(), param.numel(), float(lr), float(betas[0]),
                                                 float(betas[1]), float(eps), float(weight_decay), int(step),
                                                 _lib.stream_ptr(param.device)))


class LightGCNTrainStep:
    """Fused single-GPU training step of a :class:`recbole_gnn_b200.models.LightGCN` (its tables are updated in
    place; the Adam state lives here).  ``step(interaction)`` returns the loss as a device scalar."""

    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        self.model, self.lr, self.betas, self.eps, self.wd = model, lr, betas, eps, weight_decay
        self.t = 0
        self.state = {name: (torch.zeros_like(p.data), torch.zeros_like(p.data))
                      for name, p in (("u", model.user_embedding.weight), ("i", model.item_embedding.weight))}

    @torch.no_grad()
    def step(self, interaction: Dict[str, Tensor]) -> Tensor:
        m = self.model
        m._clear_restore()
        user, pos, neg = (interaction[k].contiguous() for k in (m.USER_ID, m.ITEM_ID, m.NEG_ITEM_ID))
        xu, xi = m.user_embedding.weight.data, m.item_embedding.weight.data
        U, L, g = xu.size(0), m.n_layers, m._graph()
        out = F_._propagate_layers(g, xu, xi, L, True)                          # forward
        g_out = torch.zeros_like(out)
        stats = bpr_loss_fused(out[:U], out[U:], xu, xi, user, pos, neg, reg_weight=m.reg_weight,
                               require_pow=m.require_pow, g_u_all=g_out[:U], g_i_all=g_out[U:])
        gx = F_._propagate_layers(g.t(), g_out[:U], g_out[U:], L, True)         # backward: same kernels on A^T
        bpr_loss_fused(out[:U], out[U:], xu, xi, user, pos, neg, reg_weight=m.reg_weight,
                       require_pow=m.require_pow, g_reg_u=gx[:U], g_reg_i=gx[U:])   # + EmbLoss rows (ego tables)
        self.t += 1
        for name, p, gr in (("u", xu, gx[:U]), ("i", xi, gx[U:])):
            adam_step(p, gr, *self.state[name], lr=self.lr, step=self.t, betas=self.betas, eps=self.eps,
                      weight_decay=self.wd)
        return stats[0]


class ShardedLightGCNTrainer:
    """Row-sharded LightGCN training: rank p holds ``xu_loc`` / ``xi_loc`` (its users' / items' rows), their Adam
    state, and a :class:`sharded.ShardedPropagator`.  ``step(user, pos, neg)`` takes the GLOBAL mini-batch ids (the
    same on every rank) and must be called collectively.  Numerically the single-GPU step, row for row."""

    def __init__(self, prop, xu_loc: Tensor, xi_loc: Tensor, n_layers: int, *, reg_weight: float = 1e-5,
                 require_pow: bool = False, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        import torch.distributed as dist
        self.dist = dist
        self.prop, self.L = prop, int(n_layers)
        self.xu, self.xi = xu_loc.contiguous(), xi_loc.contiguous()
        self.reg_weight, self.require_pow = reg_weight, require_pow
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.t = 0
        self.state = {"u": (torch.zeros_like(self.xu), torch.zeros_like(self.xu)),
                      "i": (torch.zeros_like(self.xi), torch.zeros_like(self.xi))}
        plan, r = prop.plan, prop.rank
        self.lo_u, self.hi_u, self.lo_i, self.hi_i = plan.ub[r], plan.ub[r + 1], plan.ib[r], plan.ib[r + 1]

    @torch.no_grad()
    def step(self, user: Tensor, pos: Tensor, neg: Tensor) -> Tensor:
        prop, dist = self.prop, self.dist
        dev, D, B = prop.device, prop.dim, user.numel()
        uc = self.xu.size(0)
        out = prop.forward(self.xu, self.xi, self.L)                             # ONE chain launch
        # ---- batch-sized exchange: every rank contributes the rows it owns (propagated and ego), one all-reduce
        own = [(user >= self.lo_u) & (user < self.hi_u), (pos >= self.lo_i) & (pos < self.hi_i),
               (neg >= self.lo_i) & (neg < self.hi_i)]
        loc = [user - self.lo_u, pos - self.lo_i, neg - self.lo_i]
        rows = torch.zeros(2, 3, B, D, device=dev)
        for k, (tbl_out, tbl_ego) in enumerate(((out[:uc], self.xu), (out[uc:], self.xi), (out[uc:], self.xi))):
            idx = loc[k][own[k]]
            rows[0, k, own[k]] = tbl_out[idx]
            rows[1, k, own[k]] = tbl_ego[idx]
        dist.all_reduce(rows, group=prop.group)
        # ---- the same loss kernel on the gathered rows (each sample owns its three rows: no atomics collide)
        ar = torch.arange(B, device=dev)
        i_all, reg_i = rows[0, 1:].reshape(2 * B, D), rows[1, 1:].reshape(2 * B, D)
        g_rows, g_reg = torch.zeros(3, B, D, device=dev), torch.zeros(3, B, D, device=dev)
        stats = bpr_loss_fused(rows[0, 0], i_all, rows[1, 0], reg_i, ar, ar, ar + B, reg_weight=self.reg_weight,
                               require_pow=self.require_pow, g_u_all=g_rows[0], g_i_all=g_rows[1:].reshape(2 * B, D),
                               g_reg_u=g_reg[0], g_reg_i=g_reg[1:].reshape(2 * B, D))
        # ---- scatter the gradient rows this rank owns, propagate them back (ONE chain launch), add the EmbLoss rows
        g_out = torch.zeros(prop.n_loc, D, device=dev)
        for k, off in enumerate((0, uc, uc)):
            g_out.index_add_(0, loc[k][own[k]] + off, g_rows[k][own[k]])
        gx = prop.forward(g_out[:uc], g_out[uc:], self.L)
        for k, off in enumerate((0, uc, uc)):
            gx.index_add_(0, loc[k][own[k]] + off, g_reg[k][own[k]])
        self.t += 1
        for name, p, gr in (("u", self.xu, gx[:uc]), ("i", self.xi, gx[uc:])):
            adam_step(p, gr.contiguous(), *self.state[name], lr=self.lr, step=self.t, betas=self.betas, eps=self.eps,
                      weight_decay=self.wd)
        return stats[0]
