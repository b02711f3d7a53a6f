"""Propagation entry points over a resident :class:`GraphHandle`, all dispatching to libb200gcn.

* :func:`spmm` — ``y = A x`` with autograd (backward = the same kernel on ``A^T``); the body of
  ``LightGCNConv`` / ``BipartiteGCNConv`` / ``BiGNNConv.propagate`` (recbole_gnn/model/layers.py:13-20,31-35,55).
* :func:`lightgcn_propagate` — the whole of ``LightGCN.forward`` (lightgcn.py:70-81): K layers with the
  layer mean fused into the SpMM epilogue, user/item tables read in place (no cat/stack/mean/split).
* :func:`simgcl_propagate` — ``SimGCL.forward(perturbed)`` (simgcl.py:24-38) with the sign-noise
  perturbation fused (caller-supplied ``rand_like`` draws, or in-kernel Philox).
* :func:`bignn_tail` / :func:`ngcf_forward` — NGCF layer tail and layer stack (layers.py:56-58, ngcf.py:92-102).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from .graph import GraphHandle

Tensor = torch.Tensor

# tuning word forwarded as b200gcn_spmm_args.flags (see spmm.cu `dispatch`); 0 = engine defaults
DEFAULT_FLAGS = int(__import__("os").environ.get("B200GCN_FLAGS", "0"), 0)


def _f32_rows(t: Tensor, what: str) -> Tensor:
    """fp32, 2-D, unit inner stride, 16-byte aligned rows (row stride % 4 == 0); copies only if needed."""
    if t.dtype != torch.float32:
        raise TypeError(f"{what} must be float32 (the reference's embedding dtype), got {t.dtype}")
    if t.dim() != 2:
        raise ValueError(f"{what} must be 2-D [rows, dim]")
    if t.size(1) % 4 != 0:
        raise ValueError(f"{what}: dim={t.size(1)} must be a multiple of 4")
    if t.stride(1) != 1 or t.stride(0) % 4 != 0 or t.data_ptr() % 16 != 0 or (t.size(0) > 1 and t.stride(0) < t.size(1)):
        t = t.contiguous()
    return t


def _ld(t: Tensor) -> int:
    return t.stride(0) if t.size(0) > 1 else max(t.stride(0), t.size(1))


class PeerTables:
    """Where the fused exchange epilogue stores finished rows: the next-layer gather table of every rank
    (peer-mapped pointers, ``ptrs_dev`` = device array of ``n_peers`` pointers) or one NVSwitch multicast
    address (``mc_ptr``); ``row0`` = first row of this rank's block, ``ld`` = leading dimension."""

    def __init__(self, ptrs_dev: int, n_peers: int, row0: int, ld: int, mc_ptr: Optional[int] = None,
                 need: Optional[Tensor] = None):
        self.ptrs_dev, self.n_peers, self.row0, self.ld = ptrs_dev, n_peers, row0, ld
        self.mc_ptr = mc_ptr
        self.need = need          # int32 [n_local_rows]: bit q = rank q reads the row (halo-only exchange); None = all
        if mc_ptr:
            self.ptrs_dev, self.n_peers = None, 0


class LaunchTimer:
    """Brackets every engine kernel launch with CUDA events on the launching stream while active
    (used by bench.py for the per-launch roofline; a few microseconds per launch)."""

    _active: Optional["LaunchTimer"] = None

    def __init__(self):
        self.pairs = []

    def __enter__(self):
        LaunchTimer._active = self
        return self

    def __exit__(self, *exc):
        LaunchTimer._active = None

    @property
    def count(self) -> int:
        return len(self.pairs)

    def durations_ms(self):
        torch.cuda.synchronize()
        return [s.elapsed_time(e) for s, e in self.pairs]


def spmm_args(g: Optional[GraphHandle], x: Tensor, *, x2: Optional[Tensor] = None, y: Optional[Tensor] = None,
              noise: Optional[Tensor] = None, eps: float = 0.0, seed: int = 0,
              acc_in: Optional[Tensor] = None, acc_in2: Optional[Tensor] = None,
              acc_out: Optional[Tensor] = None, acc_scale: float = 1.0,
              peers: Optional["PeerTables"] = None, rows: Optional[Tuple[int, int]] = None,
              peer_row_offset: int = 0, acc_extra: Sequence[Tensor] = ()) -> "_lib.SpmmArgs":
    """Validated ``b200gcn_spmm_args`` for one launch (or one phase of a chain); see :func:`spmm_raw`."""
    _lib.require_cuda(x, x2, y, noise, acc_in, acc_in2, acc_out, what="spmm operand")
    if g is None:
        rowptr = col = val = None
        n_rows = n_src = x.size(0) + (x2.size(0) if x2 is not None else 0)
    else:
        if not isinstance(g, GraphHandle) or not g.is_resident:
            raise RuntimeError("spmm needs a resident GraphHandle (call .to('cuda'))")
        rowptr, col, val = g.csr()
        n_rows, n_src = g.sparse_sizes()
    r0 = 0
    if rows is not None:
        if g is None or g._n_hubs > 0:
            raise ValueError("rows= needs a graph without a hub plan")
        r0, r1 = int(rows[0]), int(rows[1])
        if not (0 <= r0 <= r1 <= n_rows):
            raise ValueError("rows out of range")
        n_rows = r1 - r0
    have = x.size(0) + (x2.size(0) if x2 is not None else 0)
    if have != n_src:
        raise ValueError(f"x holds {have} rows but the graph has {n_src} source nodes")
    D = x.size(1)
    a = _lib.SpmmArgs()
    a.n_rows, a.dim, a.flags = n_rows, D, DEFAULT_FLAGS
    a.rowptr, a.col, a.val = (None if rowptr is None else rowptr.data_ptr() + 8 * r0), _lib.ptr(col), _lib.ptr(val)
    a.x, a.x2, a.x_split, a.ldx = x.data_ptr(), _lib.ptr(x2), x.size(0), _ld(x)
    if x2 is not None and (_ld(x2) != _ld(x) or x2.size(1) != D):
        raise ValueError("x and x2 must share dim and row stride")
    for name, t in (("y", y), ("noise", noise), ("acc_out", acc_out)):
        if t is not None and (t.size(0) != n_rows or t.size(1) != D):
            raise ValueError(f"{name} must be [{n_rows}, {D}]")
    a.y, a.ldy = _lib.ptr(y), (_ld(y) if y is not None else 0)
    a.noise, a.ldn = _lib.ptr(noise), (_ld(noise) if noise is not None else 0)
    a.eps, a.acc_scale, a.seed = float(eps), float(acc_scale), int(seed) & (2 ** 64 - 1)
    a.acc_in, a.acc_in2 = _lib.ptr(acc_in), _lib.ptr(acc_in2)
    a.acc_split = acc_in.size(0) if acc_in is not None else 0
    a.ld_acc_in = _ld(acc_in) if acc_in is not None else 0
    if acc_in is not None:
        tot = acc_in.size(0) + (acc_in2.size(0) if acc_in2 is not None else 0)
        if tot != n_rows or acc_in.size(1) != D:
            raise ValueError("acc_in must cover the destination rows")
        if acc_in2 is not None and _ld(acc_in2) != _ld(acc_in):
            raise ValueError("acc_in and acc_in2 must share the row stride")
    a.acc_out, a.ld_acc_out = _lib.ptr(acc_out), (_ld(acc_out) if acc_out is not None else 0)
    if acc_extra:
        if len(acc_extra) > 3 or acc_out is None:
            raise ValueError("at most 3 acc_extra tensors, and acc_out is required")
        _lib.require_cuda(*acc_extra, what="acc_extra")
        for k, t in enumerate(acc_extra):
            if t.shape != (n_rows, D) or _ld(t) != _ld(acc_extra[0]):
                raise ValueError("acc_extra tensors must be [n_rows, dim] with one common row stride")
            a.acc_extra[k] = t.data_ptr()
        a.n_acc_extra, a.ld_acc_extra = len(acc_extra), _ld(acc_extra[0])
    if peers is not None:
        a.y_peers, a.n_peers = peers.ptrs_dev, peers.n_peers
        a.y_mc, a.y_peer_row0, a.ld_peer = peers.mc_ptr, peers.row0 + r0 + int(peer_row_offset), peers.ld
        if peers.need is not None and not peers.mc_ptr:
            a.peer_need = peers.need.data_ptr() + 4 * (r0 + int(peer_row_offset))
    return a


def spmm_raw(g: Optional[GraphHandle], x: Tensor, **kw) -> None:
    """One launch of ``b200gcn_spmm_planned`` (no autograd, no allocation).  ``x2``/``acc_in2`` are the
    second (item) tables of the two-table form; the split is ``x.size(0)`` / ``acc_in.size(0)``.
    ``g=None`` selects the identity mode (p = x): only the epilogues run.  ``rows=(r0, r1)`` restricts the launch
    to destination rows [r0, r1) of the graph; ``y`` / ``noise`` / ``acc_*`` are then the [r1-r0, D] tensors of
    those rows and ``peer_row_offset`` shifts the peer-table row (graphs with a hub plan are not row-split)."""
    a = spmm_args(g, x, **kw)
    D = x.size(1)
    dev = x.device
    timer = LaunchTimer._active
    with torch.cuda.device(dev):
        if timer is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        if g is None:
            _lib.check(_lib.load().b200gcn_spmm(C.byref(a), _lib.stream_ptr(dev)))
        else:
            # rows up to long_row entries: row kernel; hub rows: chunked multi-CTA path
            _lib.check(_lib.load().b200gcn_spmm_planned(C.byref(a), g._long_row, None, 0, _lib.stream_ptr(dev)))
            if g._n_hubs > 0:
                hp = g._hub_plan(D)
                _lib.check(_lib.load().b200gcn_spmm_hubs(C.byref(a), C.byref(hp), _lib.stream_ptr(dev)))
        if timer is not None:
            ev[1].record()
            timer.pairs.append(ev)


def spmm_chain(phases: Sequence["_lib.SpmmArgs"], sync: "_lib.ChainSync", device) -> None:
    """``b200gcn_spmm_chain``: the phases as ONE persistent cooperative kernel ordered by device-side (cross-GPU)
    flags — the row-sharded K-layer forward in a single launch (include/b200gcn.h)."""
    n = len(phases)
    arr = (_lib.SpmmArgs * n)(*phases)
    timer = LaunchTimer._active
    with torch.cuda.device(device):
        if timer is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        _lib.check(_lib.load().b200gcn_spmm_chain(arr, n, C.byref(sync), _lib.stream_ptr(device)))
        if timer is not None:
            ev[1].record()
            timer.pairs.append(ev)


class _SpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, g: GraphHandle) -> Tensor:
        x = _f32_rows(x, "x")
        y = torch.empty(g.size(0), x.size(1), dtype=torch.float32, device=x.device)
        spmm_raw(g, x, y=y)
        ctx.g = g
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        gy = _f32_rows(gy, "grad")
        gt = ctx.g.t()                       # symmetric graphs return themselves
        gx = torch.empty(gt.size(0), gy.size(1), dtype=torch.float32, device=gy.device)
        spmm_raw(gt, gy, y=gx)
        return gx, None


def spmm(g: GraphHandle, x: Tensor) -> Tensor:
    """``torch_sparse.matmul(adj_t, x, reduce='add')`` (layers.py:19-20) on the engine."""
    _lib.require_cuda(x, what="x")
    return _SpMM.apply(x, g)


# ------------------------------------------------------------------------------------------------
# fused K-layer propagation (LightGCN / SimGCL)
# ------------------------------------------------------------------------------------------------
def _propagate_layers(g: GraphHandle, xu: Tensor, xi: Optional[Tensor], n_layers: int, include_ego: bool,
                      eps: float = 0.0, noises: Optional[Sequence[Tensor]] = None, seed: int = 0) -> Tensor:
    """out = mean over {x_0 (if include_ego), x_1..x_L}, x_{l+1} = perturb(A x_l); returns [N, D]."""
    N = g.size(0)
    D = xu.size(1)
    dev = xu.device
    n_terms = n_layers + (1 if include_ego else 0)
    if n_layers == 0:
        return torch.cat([xu, xi], 0) if xi is not None else xu.clone()
    out = torch.empty(N, D, dtype=torch.float32, device=dev)
    new = lambda: torch.empty(N, D, dtype=torch.float32, device=dev)
    n_mid = n_layers - 1                                   # layer outputs that feed the combine as addends
    if n_mid - (0 if include_ego else 1) <= 3:
        # the earlier layers write only y; the LAST layer forms the mean from x_0 (in place, two tables) and the
        # stored y_1..y_{L-1} in its epilogue: 1 GB fewer writes per 3-layer step than a running sum
        ys = []
        cur, cur2 = xu, xi
        for l in range(1, n_layers + 1):
            last = l == n_layers
            noise = None if noises is None else noises[l - 1]
            if not last:
                y = new()
                spmm_raw(g, cur, x2=cur2, y=y, noise=noise, eps=eps, seed=seed + l)
                ys.append(y)
                cur, cur2 = y, None
            else:
                if include_ego:
                    acc_in, acc_in2, extra = xu, xi, ys
                else:
                    acc_in, acc_in2, extra = (ys[0] if ys else None), None, ys[1:]
                spmm_raw(g, cur, x2=cur2, noise=noise, eps=eps, seed=seed + l, acc_in=acc_in, acc_in2=acc_in2,
                         acc_extra=extra, acc_out=out, acc_scale=1.0 / n_terms)
        return out
    bufs = [new() for _ in range(min(2, n_layers - 1))]
    cur, cur2 = xu, xi
    for l in range(1, n_layers + 1):
        last = l == n_layers
        y = None if last else bufs[(l - 1) % 2]
        if l == 1:
            acc_in, acc_in2 = (xu, xi) if include_ego else (None, None)
        else:
            acc_in, acc_in2 = out, None
        noise = None if noises is None else noises[l - 1]
        spmm_raw(g, cur, x2=cur2, y=y, noise=noise, eps=eps, seed=seed + l,
                 acc_in=acc_in, acc_in2=acc_in2, acc_out=out, acc_scale=(1.0 / n_terms) if last else 1.0)
        cur, cur2 = y, None
    return out


class _LayerMean(torch.autograd.Function):
    """K-layer propagation + layer mean; grad = the same recurrence on A^T (the perturbation of SimGCL
    has unit Jacobian almost everywhere: d(e + sign(e) n eps)/de = 1)."""

    @staticmethod
    def forward(ctx, xu, xi, g, n_layers, include_ego, eps, noises, seed):
        xu, xi = _f32_rows(xu, "user table"), _f32_rows(xi, "item table")
        out = _propagate_layers(g, xu, xi, n_layers, include_ego, eps, noises, seed)
        ctx.g, ctx.n_layers, ctx.include_ego, ctx.U = g, n_layers, include_ego, xu.size(0)
        return out[: xu.size(0)], out[xu.size(0):]

    @staticmethod
    def backward(ctx, gu, gi):
        U = ctx.U
        gt = ctx.g.t()
        if gu is None and gi is None:
            return (None,) * 8
        if gu is None:
            gu = torch.zeros(U, gi.size(1), dtype=torch.float32, device=gi.device)
        if gi is None:
            gi = torch.zeros(gt.size(0) - U, gu.size(1), dtype=torch.float32, device=gu.device)
        gu, gi = _f32_rows(gu, "grad_users"), _f32_rows(gi, "grad_items")
        gx = _propagate_layers(gt, gu, gi, ctx.n_layers, ctx.include_ego)
        return gx[:U], gx[U:], None, None, None, None, None, None


def lightgcn_propagate(g: GraphHandle, user_weight: Tensor, item_weight: Tensor, n_layers: int) -> Tuple[Tensor, Tensor]:
    """``LightGCN.forward`` (lightgcn.py:70-81): returns ``(user_all_embeddings, item_all_embeddings)``."""
    _lib.require_cuda(user_weight, item_weight, what="embedding table")
    return _LayerMean.apply(user_weight, item_weight, g, int(n_layers), True, 0.0, None, 0)


def simgcl_propagate(g: GraphHandle, user_weight: Tensor, item_weight: Tensor, n_layers: int, eps: float,
                     perturbed: bool = False, noises: Optional[Sequence[Tensor]] = None,
                     seed: Optional[int] = None) -> Tuple[Tensor, Tensor]:
    """``SimGCL.forward(perturbed)`` (simgcl.py:24-38).  ``noises``: per-layer ``[N, D]`` U[0,1) draws (what
    ``torch.rand_like`` returns in the reference); if None while ``perturbed``, the kernel draws them from
    Philox keyed by ``seed`` (default: a fresh seed from torch's generator)."""
    _lib.require_cuda(user_weight, item_weight, what="embedding table")
    if not perturbed:
        return _LayerMean.apply(user_weight, item_weight, g, int(n_layers), False, 0.0, None, 0)
    if noises is not None:
        noises = [_f32_rows(n, "noise") for n in noises]
        if len(noises) != n_layers:
            raise ValueError("one noise tensor per layer")
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    return _LayerMean.apply(user_weight, item_weight, g, int(n_layers), False, float(eps), noises, int(seed))


class _SimGCLViews(torch.autograd.Function):
    """The three forwards of one SimGCL training step (simgcl.py:48-55: one clean via
    ``super().calculate_loss`` -> ``forward()``, two perturbed) in 1 + 3(L-1) SpMMs instead of 3L: every view
    starts from the same ``A x0``, so the first layer is computed once and the perturbed first-layer rows are
    derived from it by the identity mode of the kernel.  All three views are the same linear map of x0 (the
    perturbation has unit Jacobian a.e.), so the backward is ONE propagation of the summed upstream grads."""

    @staticmethod
    def forward(ctx, xu, xi, g, n_layers, eps, noises1, noises2, seed1, seed2):
        xu, xi = _f32_rows(xu, "user table"), _f32_rows(xi, "item table")
        N, D, dev, L = g.size(0), xu.size(1), xu.device, n_layers
        new = lambda: torch.empty(N, D, dtype=torch.float32, device=dev)
        y1 = new()
        accs = [new() for _ in range(3)]
        last = L == 1
        sc = 1.0 / L
        spmm_raw(g, xu, x2=xi, y=y1)                                              # shared first layer
        views = []
        for v, (noises, seed) in enumerate(((None, 0), (noises1, seed1), (noises2, seed2))):
            pert = v > 0
            yv = None if last else (y1 if not pert else new())
            if not pert:
                # clean view: acc = y1 (* 1/L when L == 1)
                if last:
                    accs[0].copy_(y1).mul_(sc)
                else:
                    accs[0].copy_(y1)
            else:
                spmm_raw(None, y1, y=yv, noise=None if noises is None else noises[0], eps=eps, seed=seed + 1,
                         acc_out=accs[v], acc_scale=sc if last else 1.0)
            cur = yv
            bufs = [new() for _ in range(min(2, L - 2))] if L > 2 else []
            for l in range(2, L + 1):
                fin = l == L
                y = None if fin else bufs[(l - 2) % 2]
                spmm_raw(g, cur, y=y, noise=None if (not pert or noises is None) else noises[l - 1],
                         eps=eps if pert else 0.0, seed=seed + l, acc_in=accs[v], acc_out=accs[v],
                         acc_scale=sc if fin else 1.0)
                cur = y
            views.append(accs[v])
        ctx.g, ctx.L, ctx.U = g, L, xu.size(0)
        U = xu.size(0)
        return tuple(t for a in views for t in (a[:U], a[U:]))

    @staticmethod
    def backward(ctx, *grads):
        U, gt = ctx.U, ctx.g.t()
        gu = sum(g_ for g_ in grads[0::2] if g_ is not None)
        gi = sum(g_ for g_ in grads[1::2] if g_ is not None)
        gu, gi = _f32_rows(gu, "grad_users"), _f32_rows(gi, "grad_items")
        gx = _propagate_layers(gt, gu, gi, ctx.L, False)
        return gx[:U], gx[U:], None, None, None, None, None, None, None


def simgcl_views(g: GraphHandle, user_weight: Tensor, item_weight: Tensor, n_layers: int, eps: float,
                 noises1: Optional[Sequence[Tensor]] = None, noises2: Optional[Sequence[Tensor]] = None,
                 seeds: Optional[Tuple[int, int]] = None):
    """All three forwards of a SimGCL step: returns ``(u, i), (u', i'), (u'', i'')`` = clean, perturbed 1,
    perturbed 2 (simgcl.py:49,54-55).  ``noises1/2``: per-layer rand_like draws (parity runs); otherwise Philox
    noise keyed by ``seeds``."""
    _lib.require_cuda(user_weight, item_weight, what="embedding table")
    if n_layers < 1:
        raise ValueError("n_layers >= 1")
    if seeds is None:
        r = torch.randint(0, 2 ** 62, (2,))
        seeds = (int(r[0]), int(r[1]))
    if noises1 is not None:
        noises1 = [_f32_rows(n, "noise") for n in noises1]
    if noises2 is not None:
        noises2 = [_f32_rows(n, "noise") for n in noises2]
    o = _SimGCLViews.apply(user_weight, item_weight, g, int(n_layers), float(eps), noises1, noises2,
                           int(seeds[0]), int(seeds[1]))
    return (o[0], o[1]), (o[2], o[3]), (o[4], o[5])


# ------------------------------------------------------------------------------------------------
# NGCF
# ------------------------------------------------------------------------------------------------
def bignn_tail(p: Tensor, x: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, *, slope: float = 0.2,
               keep: Optional[Tensor] = None, drop_p: float = 0.0, normalize: bool = True,
               out: Optional[Tensor] = None, out2: Optional[Tensor] = None, pre_out: Optional[Tensor] = None,
               activate: bool = True) -> Optional[Tensor]:
    """Everything of an NGCF layer after the SpMM in one pass (layers.py:56-58 + ngcf.py:96-98).
    With ``activate=False, normalize=False`` and ``pre_out`` it returns what ``BiGNNConv.forward`` returns."""
    _lib.require_cuda(p, x, w1, b1, w2, b2, keep, out, out2, pre_out, what="bignn_tail operand")
    p, x = _f32_rows(p, "p"), _f32_rows(x, "x")
    w1, w2 = w1.contiguous(), w2.contiguous()
    b1, b2 = b1.contiguous(), b2.contiguous()
    n, d_in = x.shape
    d_out = w1.size(0)
    if out is None and pre_out is None:
        out = torch.empty(n, d_out, dtype=torch.float32, device=x.device)
    if keep is not None:
        keep = keep.to(torch.uint8).contiguous()
        if keep.shape != (n, d_out):
            raise ValueError("keep must be [n, d_out]")
    dev = x.device
    with torch.cuda.device(dev):
        _lib.check(_lib.load().b200gcn_bignn_tail(
            p.data_ptr(), _ld(p), x.data_ptr(), _ld(x), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
            n, d_in, d_out, float(slope) if activate else 1.0, _lib.ptr(keep), float(drop_p), int(bool(normalize)),
            _lib.ptr(out), _ld(out) if out is not None else 0, _lib.ptr(out2), _ld(out2) if out2 is not None else 0,
            _lib.ptr(pre_out),
            _ld(pre_out) if pre_out is not None else 0, _lib.stream_ptr(dev)))
    return out if out is not None else pre_out


def bignn_tail_backward(p: Tensor, x: Tensor, w1: Tensor, w2: Tensor, t: Tensor, out: Tensor,
                        keep_scale: Optional[Tensor], slope: float, normalize: bool, g_out: Tensor):
    """Gradients of the fused NGCF layer tail (device-agnostic torch algebra; the GEMMs are library calls):
        t = (p + x) W1^T + b1 + (p * x) W2^T + b2 ; z = leaky_relu(t) * keep_scale ; out = z / max(||z||, 1e-12)
    given ``g_out`` = dL/d out.  ``keep_scale`` = keep / (1 - p_drop) as float, or None.  ``t`` (pre-activation)
    and ``out`` are what the forward kernel wrote.  Returns (g_p, g_x, g_w1, g_b1, g_w2, g_b2)."""
    act = torch.where(t > 0, torch.ones_like(t), torch.full_like(t, slope))
    if keep_scale is not None:
        act = act * keep_scale
    if normalize:
        z = torch.where(t > 0, t, t * slope)
        if keep_scale is not None:
            z = z * keep_scale
        n = z.norm(dim=1, keepdim=True)
        live = n > 1e-12                                   # below F.normalize's eps the map is z / eps
        dot = (out * g_out).sum(dim=1, keepdim=True)
        g_z = torch.where(live, (g_out - out * dot) / n.clamp_min(1e-12), g_out / 1e-12)
    else:
        g_z = g_out
    g_t = g_z * act
    a, m = p + x, p * x
    g_a, g_m = g_t @ w1, g_t @ w2
    g_w1, g_w2 = g_t.t() @ a, g_t.t() @ m
    g_b = g_t.sum(dim=0)
    return g_a + g_m * x, g_a + g_m * p, g_w1, g_b, g_w2, g_b.clone()


def bignn_tail_backward_fused(p: Tensor, x: Tensor, w1: Tensor, w2: Tensor, t: Tensor, keep: Optional[Tensor],
                              drop_p: float, slope: float, normalize: bool, g_out: Tensor):
    """``b200gcn_bignn_tail_backward`` (d_in = d_out = 64): the row-local backward of normalise / dropout / LeakyReLU,
    the ``g_t [W1 | W2]`` contraction on tcgen05 and ``g_p`` / ``g_x`` in one pass; the weight gradient is then one
    plain library GEMM over the ``am = [p + x | p * x]`` rows the kernel also wrote.
    Returns (g_p, g_x, g_w1, g_b1, g_w2, g_b2)."""
    n, d = x.shape
    dev = x.device
    g_p, g_x, g_t = (torch.empty(n, d, dtype=torch.float32, device=dev) for _ in range(3))
    am = torch.empty(n, 2 * d, dtype=torch.float32, device=dev)
    g_out = _f32_rows(g_out, "grad")
    with torch.cuda.device(dev):
        _lib.check(_lib.load().b200gcn_bignn_tail_backward(
            p.data_ptr(), _ld(p), x.data_ptr(), _ld(x), w1.data_ptr(), w2.data_ptr(), t.data_ptr(), _ld(t),
            _lib.ptr(keep), float(drop_p) if keep is not None else 0.0, float(slope), int(bool(normalize)),
            g_out.data_ptr(), _ld(g_out), n, d, d, g_p.data_ptr(), g_x.data_ptr(), g_t.data_ptr(), am.data_ptr(),
            _lib.stream_ptr(dev)))
    g_w = g_t.t() @ am                                   # [64, n] x [n, 128]: a plain library GEMM
    g_b = g_t.sum(dim=0)
    return g_p, g_x, g_w[:, :d].contiguous(), g_b, g_w[:, d:].contiguous(), g_b.clone()


class _BiGNNTail(torch.autograd.Function):
    """Forward = the fused tail kernel (one pass, also writes the pre-activation t that the backward needs);
    backward = the fused tcgen05 backward kernel for 64 x 64 layers (:func:`bignn_tail_backward_fused`), the dense
    algebra of :func:`bignn_tail_backward` for other shapes."""

    @staticmethod
    def forward(ctx, p, x, w1, b1, w2, b2, slope, keep, drop_p, normalize):
        p, x = _f32_rows(p, "p"), _f32_rows(x, "x")
        n, d_out = x.size(0), w1.size(0)
        out = torch.empty(n, d_out, dtype=torch.float32, device=x.device)
        t = torch.empty(n, d_out, dtype=torch.float32, device=x.device)
        bignn_tail(p, x, w1, b1, w2, b2, slope=slope, keep=keep, drop_p=drop_p if keep is not None else 0.0,
                   normalize=normalize, out=out, pre_out=t)
        ctx.save_for_backward(p, x, w1, w2, t, out, keep if keep is not None else torch.empty(0, device=x.device))
        ctx.slope, ctx.drop_p, ctx.normalize, ctx.has_keep = float(slope), float(drop_p), bool(normalize), keep is not None
        return out

    @staticmethod
    def backward(ctx, g_out):
        p, x, w1, w2, t, out, keep = ctx.saved_tensors
        if x.size(1) == 64 and w1.size(0) == 64 and (not ctx.has_keep or keep.data_ptr() % 4 == 0):
            g_p, g_x, g_w1, g_b1, g_w2, g_b2 = bignn_tail_backward_fused(
                p, x, w1.contiguous(), w2.contiguous(), t, keep if ctx.has_keep else None, ctx.drop_p, ctx.slope,
                ctx.normalize, g_out)
            return g_p, g_x, g_w1, g_b1, g_w2, g_b2, None, None, None, None
        ks = keep.to(torch.float32) * (1.0 / (1.0 - ctx.drop_p)) if ctx.has_keep else None
        g_p, g_x, g_w1, g_b1, g_w2, g_b2 = bignn_tail_backward(p, x, w1, w2, t, out, ks, ctx.slope, ctx.normalize,
                                                               g_out.contiguous())
        return g_p, g_x, g_w1, g_b1, g_w2, g_b2, None, None, None, None


def bignn_tail_autograd(p: Tensor, x: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, *, slope: float = 0.2,
                        keep: Optional[Tensor] = None, drop_p: float = 0.0, normalize: bool = True) -> Tensor:
    """Differentiable fused NGCF layer tail (layers.py:56-58 + ngcf.py:96-98) for training."""
    _lib.require_cuda(p, x, w1, b1, w2, b2, keep, what="bignn_tail operand")
    if keep is not None:
        keep = keep.to(torch.uint8).contiguous()
    return _BiGNNTail.apply(p, x, w1, b1, w2, b2, float(slope), keep, float(drop_p), bool(normalize))


def ngcf_forward(g: GraphHandle, user_weight: Tensor, item_weight: Tensor,
                 weights: Sequence[Tuple[Tensor, Tensor, Tensor, Tensor]], *, slope: float = 0.2,
                 message_dropout: float = 0.0, keep_masks: Optional[Sequence[Tensor]] = None) -> Tuple[Tensor, Tensor]:
    """Inference-mode ``NGCF.forward`` with node_dropout == 0 (ngcf.py:92-104): per layer one SpMM and one
    fused tail kernel that writes its normalised output straight into the column slice of the
    ``[N, sum(dims)]`` concat buffer.  (Training goes through ``BiGNNConv`` so that autograd sees the
    dense tail; the SpMM there is the same kernel.)"""
    _lib.require_cuda(user_weight, item_weight, what="embedding table")
    xu, xi = _f32_rows(user_weight.detach(), "user table"), _f32_rows(item_weight.detach(), "item table")
    U, D0 = xu.shape
    N = U + xi.size(0)
    dims = [D0] + [w[0].size(0) for w in weights]
    dev = xu.device
    out = torch.empty(N, sum(dims), dtype=torch.float32, device=dev)
    out[:U, :D0].copy_(xu)
    out[U:, :D0].copy_(xi)
    # gather tables stay CONTIGUOUS (ping-pong); the concat slices are written as second outputs of the tail:
    # gathering straight from a [N, 256]-strided slice costs ~60 % more (quarter of the L2 sets / DRAM banks)
    x, x2 = xu, xi
    xcat = None
    off = 0
    n_l = len(weights)
    for l, (w1, b1, w2, b2) in enumerate(weights):
        p = torch.empty(N, dims[l], dtype=torch.float32, device=dev)
        spmm_raw(g, x, x2=x2, y=p)
        x_in = out[:, off:off + dims[l]]          # the layer input as one [N, d] view (for the p + x / p * x terms)
        off += dims[l]
        keep = None if keep_masks is None else keep_masks[l]
        nxt = None if l == n_l - 1 else torch.empty(N, dims[l + 1], dtype=torch.float32, device=dev)
        bignn_tail(p, x_in, w1.detach(), b1.detach(), w2.detach(), b2.detach(), slope=slope, keep=keep,
                   drop_p=message_dropout if keep is not None else 0.0, normalize=True,
                   out=out[:, off:off + dims[l + 1]], out2=nxt)
        x, x2 = nxt, None
    return out[:U], out[U:]


# ------------------------------------------------------------------------------------------------
# full-sort evaluation (lightgcn.py:123-133, ngcf.py:138-150)
# ------------------------------------------------------------------------------------------------
def full_sort_scores(u: Tensor, items: Tensor) -> Tensor:
    """``torch.matmul(u_embeddings, restore_item_e.transpose(0, 1))`` (lightgcn.py:131): the dense
    ``[batch, n_items]`` score matrix the reference's ``full_sort_predict`` returns — tcgen05 kernel with the
    fp32-accurate TF32 split (``b200gcn_fullsort_scores``); dims that are not a multiple of 8 are zero-padded."""
    _lib.require_cuda(u, items, what="full_sort operand")
    u, items = _pad8(_f32_rows(u.detach(), "user rows")), _pad8(_f32_rows(items.detach(), "item table"))
    out = torch.empty(u.size(0), items.size(0), dtype=torch.float32, device=u.device)
    with torch.cuda.device(u.device):
        _lib.check(_lib.load().b200gcn_fullsort_scores(u.data_ptr(), _ld(u), u.size(0), items.data_ptr(), _ld(items),
                                                       items.size(0), u.size(1), out.data_ptr(), out.stride(0) if
                                                       out.size(0) > 1 else out.size(1), _lib.stream_ptr(u.device)))
    return out


def _pad8(t: Tensor) -> Tensor:
    d = t.size(1)
    return t if d % 8 == 0 else torch.nn.functional.pad(t, (0, 8 - d % 8))


def full_sort_topk(u: Tensor, items: Tensor, k: int, history=None, first_item: int = 1) -> Tuple[Tensor, Tensor]:
    """Top-``k`` ``(scores, item ids)`` per user row without the ``[batch, n_items]`` matrix
    (``b200gcn_fullsort_topk``).  ``history = (row_idx, item_idx)``: seen interactions to exclude, the form RecBole's
    full-sort evaluator holds them in; ``first_item = 1`` excludes the [PAD] item like RecBole's ``scores[:, 0] = -inf``."""
    _lib.require_cuda(u, items, what="full_sort operand")
    u, items = _pad8(_f32_rows(u.detach(), "user rows")), _pad8(_f32_rows(items.detach(), "item table"))
    B, I, dev = u.size(0), items.size(0), u.device
    if k > 64:   # beyond the fused kernel's candidate list: the engine's dense scores, then the library selection
        s = full_sort_scores(u, items)
        if history is not None and history[0].numel() > 0:
            s[history[0].to(dev).long(), history[1].to(dev).long()] = float("-inf")
        s[:, :first_item] = float("-inf")
        return torch.topk(s, k, dim=1)
    hist_ptr = hist_items = None
    if history is not None and history[0].numel() > 0:
        rows, its = history[0].to(dev).long(), history[1].to(dev).long()
        order = torch.argsort(rows * I + its)
        hist_items = its[order].contiguous()
        hist_ptr = torch.zeros(B + 1, dtype=torch.int64, device=dev)
        hist_ptr[1:] = torch.cumsum(torch.bincount(rows, minlength=B), 0)
    scores = torch.empty(B, k, dtype=torch.float32, device=dev)
    ids = torch.empty(B, k, dtype=torch.int64, device=dev)
    lib = _lib.load()
    need = C.c_size_t(0)
    _lib.check(lib.b200gcn_fullsort_topk_workspace(B, I, k, C.byref(need)))
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.b200gcn_fullsort_topk(u.data_ptr(), _ld(u), B, items.data_ptr(), _ld(items), I, u.size(1), int(k),
                                             int(first_item), _lib.ptr(hist_ptr), _lib.ptr(hist_items),
                                             scores.data_ptr(), ids.data_ptr(), ws.data_ptr(), need.value,
                                             _lib.stream_ptr(dev)))
    return scores, ids
