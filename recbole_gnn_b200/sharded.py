"""Row-sharded multi-GPU propagation: one process per GPU, embedding tables and CSR rows split 1-D over the
ranks of one NVLink/NVSwitch box, one exchange per layer.  The reference has no multi-GPU
path at all (single process, `quick_start.py:41`); this is the engine's extension of
`LightGCN.forward` (lightgcn.py:70-81) to P ranks with identical numerics.

Partition (``ShardPlan``): rank p owns users [pU/P, (p+1)U/P) and items [pI/P, (p+1)I/P) — both halves of
the bipartite graph stay balanced.  Node ids are relabelled rank-major: ``new = p * n_pad + local`` with the
rank's users first, then its items, so the gathered table of a layer is the plain concatenation of the
ranks' output blocks and every rank's block is one contiguous slab.

Exchange per layer — implementations with the same result:

* ``"chain"`` (CUDA default): the WHOLE K-layer forward is one persistent cooperative kernel per step
  (``b200gcn_spmm_chain``).  Every finished row goes straight from the registers of the SpMM epilogue into the
  next-layer table of EVERY rank (``multimem.st`` through the NVSwitch multicast address, or P peer-mapped
  stores); layers are ordered by device-side flag words in symmetric memory instead of host-ordered barriers.
  The bipartite structure gives two independent dependency chains (user rows gather only item rows and vice
  versa): ``I0 -> U1 -> I2 -> U3`` and ``U0 -> I1 -> U2 -> I3``.  The kernel runs the half-layers in the order
  ``I0 U0 | U1 I1 | I2 U2 | U3 I3`` so that every half-layer depends on the one TWO phases back: the NVLink
  flight time and flag latency of one phase hide behind the tiles of the next.
* ``"fused"`` / ``"fused-split"``: one launch per layer with the same epilogue, a symmetric-memory barrier
  between launches (round-1 path; still used for graphs with hub rows).
* ``"allgather"``: local SpMM into the own block, then ``all_gather_into_tensor`` (NCCL on GPUs; gloo in the
  CPU tests of the host logic, where the local product is injected by the test).

The K-layer mean operator ``M = (I + A + .. + A^L)/(L+1)`` is symmetric (``A`` is), so the backward of the
sharded forward is the sharded forward of the row-sharded gradient (``propagate`` registers it with autograd).
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist

Tensor = torch.Tensor


class ShardPlan:
    """Pure host-side partition arithmetic (no device work; covered by the CPU tests)."""

    def __init__(self, user_num: int, item_num: int, world_size: int):
        self.U, self.I, self.P = int(user_num), int(item_num), int(world_size)
        P = self.P
        self.ub = [p * self.U // P for p in range(P + 1)]
        self.ib = [p * self.I // P for p in range(P + 1)]
        self.u_cnt = [self.ub[p + 1] - self.ub[p] for p in range(P)]
        self.i_cnt = [self.ib[p + 1] - self.ib[p] for p in range(P)]
        self.n_loc = [self.u_cnt[p] + self.i_cnt[p] for p in range(P)]
        self.n_pad = max(self.n_loc)          # rows per rank block in the gathered table
        self.n_full = self.n_pad * P

    def _owner(self, ids: Tensor, bounds: List[int]) -> Tensor:
        b = torch.tensor(bounds[1:-1], dtype=ids.dtype, device=ids.device)
        return torch.bucketize(ids, b, right=True)

    def relabel_users(self, u: Tensor) -> Tensor:
        p = self._owner(u, self.ub)
        ub = torch.tensor(self.ub, dtype=u.dtype, device=u.device)
        return p * self.n_pad + (u - ub[p])

    def relabel_items(self, i: Tensor) -> Tensor:
        p = self._owner(i, self.ib)
        ib = torch.tensor(self.ib, dtype=i.dtype, device=i.device)
        uc = torch.tensor(self.u_cnt, dtype=i.dtype, device=i.device)
        return p * self.n_pad + uc[p] + (i - ib[p])

    def local_edges(self, rank: int, uid: Tensor, iid: Tensor, w: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        """Entries of the rank's CSR rows: ``(dst_local, src_relabelled, weight)``.  ``w[k]`` is the
        gcn_norm weight of interaction k (same for both directions of the symmetric adjacency)."""
        mu = (uid >= self.ub[rank]) & (uid < self.ub[rank + 1])
        mi = (iid >= self.ib[rank]) & (iid < self.ib[rank + 1])
        dst_u = uid[mu] - self.ub[rank]                                  # user rows <- item sources
        src_u = self.relabel_items(iid[mu])
        dst_i = self.u_cnt[rank] + (iid[mi] - self.ib[rank])             # item rows <- user sources
        src_i = self.relabel_users(uid[mi])
        return torch.cat([dst_u, dst_i]), torch.cat([src_u, src_i]), torch.cat([w[mu], w[mi]])

    def scatter_tables(self, rank: int, xu: Tensor, xi: Tensor) -> Tuple[Tensor, Tensor]:
        """The rank's slices of the full (unsharded) embedding tables."""
        return xu[self.ub[rank]:self.ub[rank + 1]], xi[self.ib[rank]:self.ib[rank + 1]]

    def half_layer_order(self, n_layers: int) -> List[Tuple[int, str]]:
        """Phase order of the chain kernel after the two ego publishes: ``[(layer, 'U'|'I'), ...]``.  Odd layers
        run user rows first, even layers item rows first, so each half-layer's gather table was published two
        phases earlier (user rows of layer l gather the item rows of layer l-1 and vice versa)."""
        order = []
        for l in range(1, n_layers + 1):
            order += [(l, "U"), (l, "I")] if l % 2 == 1 else [(l, "I"), (l, "U")]
        return order


    def chain_schedule(self, n_layers: int) -> List[dict]:
        """The phase list of the chain kernel with its two dependency kinds, as pure data (the CPU tests replay it):
        ``wait`` = phase that must be complete on EVERY rank (its rows are gathered over NVLink-written tables),
        ``wait_local`` = phase that must be complete on this rank only (it wrote the rows of the running layer sum
        this phase reads and is the direct predecessor; older local phases are implied by ``wait``)."""
        # the user-half ego publish is needed only by the phase AFTER the first half-layer: its tiles are interleaved
        # into that half-layer's tiles (merge_next) instead of blocking every CTA on NVLink stores in front of it
        sched = [{"kind": "publish", "layer": 0, "half": "I", "wait": -1, "wait_local": -1, "merge_next": 0},
                 {"kind": "publish", "layer": 0, "half": "U", "wait": -1, "wait_local": -1,
                  "merge_next": int(n_layers >= 1 and bool(int(os.environ.get("B200GCN_CHAIN_MERGE", "1"))))}]
        acc_writer = {}
        for l, half in self.half_layer_order(n_layers):
            p = len(sched)
            wl = acc_writer.get(half, -1)
            sched.append({"kind": "spmm", "layer": l, "half": half, "wait": p - 2,
                          "wait_local": wl if wl > p - 2 else -1, "merge_next": 0})
            acc_writer[half] = p
        return sched


def interaction_weights_device(uid: Tensor, iid: Tensor, user_num: int, item_num: int) -> Tensor:
    """gcn_norm weight of every interaction, ``deg_u^-1/2 * deg_i^-1/2`` over the FULL graph (dataset.py:74-77
    semantics), by the device kernel behind ``get_bipartite_inter_mat(row_norm=False)``."""
    from . import _lib

    _lib.require_cuda(uid, iid, what="uid/iid")
    dev = uid.device
    w = torch.empty(uid.numel(), dtype=torch.float32, device=dev)
    ws = torch.empty(max((user_num + item_num) * 4, 4), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().b200gcn_bipartite_norm_coo(
            uid.data_ptr(), iid.data_ptr(), uid.numel(), user_num, item_num, 0, w.data_ptr(), ws.data_ptr(),
            ws.numel(), _lib.stream_ptr(dev)))
    return w


def synth_local_edges(plan: ShardPlan, rank: int, n_inter: int, dev, seed: int = 0, chunk: int = 50_000_000):
    """The rank's CSR entries of the seeded synthetic graph of ``bench.synth_graph_device`` WITHOUT holding the
    whole interaction list: the same generator stream is drawn chunk by chunk on every rank, full-graph degrees are
    accumulated from every chunk, and only the interactions touching the rank's rows are kept (2/P of them).
    Weights are ``deg_u^-1/2 deg_i^-1/2`` with IEEE ``1/sqrt`` (= gcn_norm, dataset.py:74-77)."""
    U, I = plan.U, plan.I
    gen = torch.Generator(device=dev).manual_seed(seed)
    deg_u = torch.zeros(U, dtype=torch.int64, device=dev)
    deg_i = torch.zeros(I, dtype=torch.int64, device=dev)
    ku, ki, kd = [], [], []
    lo_u, hi_u, lo_i, hi_i = plan.ub[rank], plan.ub[rank + 1], plan.ib[rank], plan.ib[rank + 1]
    for s in range(0, n_inter, chunk):
        n = min(chunk, n_inter - s)
        u = torch.randint(1, U, (n,), generator=gen, device=dev)
        i = torch.randint(1, I, (n,), generator=gen, device=dev)
        deg_u += torch.bincount(u, minlength=U)
        deg_i += torch.bincount(i, minlength=I)
        mu = (u >= lo_u) & (u < hi_u)
        mi = (i >= lo_i) & (i < hi_i)
        ku.append(torch.cat([u[mu], u[mi]]))
        ki.append(torch.cat([i[mu], i[mi]]))
        kd.append(torch.cat([torch.ones(int(mu.sum()), dtype=torch.bool, device=dev),
                             torch.zeros(int(mi.sum()), dtype=torch.bool, device=dev)]))
        del u, i, mu, mi
    u, i, is_user_row = torch.cat(ku), torch.cat(ki), torch.cat(kd)
    del ku, ki, kd
    dis_u = 1.0 / torch.sqrt(deg_u.to(torch.float32))
    dis_i = 1.0 / torch.sqrt(deg_i.to(torch.float32))
    dis_u[deg_u == 0] = 0
    dis_i[deg_i == 0] = 0
    w = dis_u[u] * dis_i[i]
    dst = torch.where(is_user_row, u - lo_u, plan.u_cnt[rank] + (i - lo_i))
    src = torch.where(is_user_row, plan.relabel_items(i), plan.relabel_users(u))
    return dst, src, w


class ShardedPropagator:
    """K-layer LightGCN propagation of one rank.  ``forward(xu_loc, xi_loc, L)`` returns the rank's rows
    ``[n_loc, D]`` (own users then own items) of ``mean(x_0 .. x_L)`` in a tensor the caller owns."""

    def __init__(self, plan: ShardPlan, rank: int, dst_local: Tensor, src_new: Tensor, w: Tensor, dim: int,
                 device, group=None, exchange: Optional[str] = None,
                 local_spmm: Optional[Callable[[Tensor, Tensor], Tensor]] = None, halo: Optional[bool] = None):
        self.plan, self.rank, self.dim = plan, rank, int(dim)
        self.device = torch.device(device)
        self.group = group if group is not None else dist.group.WORLD
        self.n_loc = plan.n_loc[rank]
        is_cuda = self.device.type == "cuda"
        if exchange is None:
            exchange = os.environ.get("B200GCN_EXCHANGE", "chain" if is_cuda else "allgather")
        if exchange not in ("chain", "fused", "fused-split", "allgather"):
            raise ValueError("exchange must be 'chain', 'fused', 'fused-split' or 'allgather'")
        if exchange != "allgather" and not is_cuda:
            raise RuntimeError("the fused exchanges need CUDA peer memory")
        self._local_spmm = local_spmm
        self.handle = None
        if local_spmm is None:
            if not is_cuda:
                raise RuntimeError("recbole_gnn_b200.sharded: no CPU propagation exists; CPU tests of the host "
                                   "logic must inject `local_spmm`")
            from .graph import GraphHandle
            # matrix rows = local destinations, columns = relabelled global sources
            self.handle = GraphHandle(row=dst_local, col=src_new, value=w,
                                      sparse_sizes=(self.n_loc, plan.n_full)).to(self.device)
        else:
            self._edges = (dst_local, src_new, w)
        if exchange == "chain":
            # the chain kernel has no hub path: graphs with hub rows on ANY rank take the per-launch exchange
            hubs = torch.tensor([self.handle._n_hubs], device=self.device)
            dist.all_reduce(hubs, op=dist.ReduceOp.MAX, group=self.group)
            if int(hubs.item()) > 0:
                exchange = "fused"
        self.split = exchange == "fused-split"
        if self.split:
            exchange = "fused"
        self.exchange = exchange
        self.shape = (plan.n_full, self.dim)
        self.bufs: List[Tensor] = []
        self.hdls = []
        self._peer = {}
        self.use_multicast = False
        if exchange in ("chain", "fused"):
            self._grow_tables(2)
        else:
            self.bufs = [torch.zeros(self.shape, dtype=torch.float32, device=self.device) for _ in range(2)]
        if exchange == "chain":
            import torch.distributed._symmetric_memory as symm_mem
            from . import _lib
            self._flags = symm_mem.empty((_lib.CHAIN_MAX_PHASES * _lib.CHAIN_MAX_RANKS,), dtype=torch.int32,
                                         device=self.device)
            self._flags.zero_()
            self._flags_hdl = symm_mem.rendezvous(self._flags, self.group)
            self._scratch = torch.zeros(_lib.CHAIN_SCRATCH_BYTES // 4, dtype=torch.int32, device=self.device)
            self._epoch = 0
            self._last_phase = -1
        # Halo-only exchange ("halo rows only where the graph partitions naturally"): a row travels only to
        # the ranks whose CSR references it.  On a uniform random graph that is every rank (nothing saved, and the
        # multicast store is the better tool); on graphs with locality it removes the NVLink traffic of interior rows.
        if halo is None:
            halo = bool(int(os.environ.get("B200GCN_HALO", "0")))
        self.halo = bool(halo) and exchange in ("chain", "fused") and self.handle is not None
        self._need = None
        self.halo_traffic_fraction = 1.0
        if self.halo:
            ref = torch.zeros(plan.n_full, dtype=torch.uint8, device=self.device)
            ref[self.handle.csr()[1].long().unique()] = 1
            allref = torch.empty(plan.P, plan.n_full, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(allref, ref, group=self.group)
            r0 = rank * plan.n_pad
            mine = allref[:, r0:r0 + self.n_loc].to(torch.int32)
            shifts = torch.arange(plan.P, device=self.device, dtype=torch.int32).view(-1, 1)
            self._need = (mine << shifts).sum(0).to(torch.int32).contiguous()
            others = mine.sum(0) - mine[rank]
            self.halo_traffic_fraction = float(others.float().mean().item()) / max(plan.P - 1, 1)
            self.use_multicast = False
        self.acc = None
        if self.split and (self.handle is None or self.handle._n_hubs > 0):
            self.split = False          # row-split launches need a hub-free graph
        self.side = torch.cuda.Stream(self.device) if self.split else None
        self._barrier()

    # ------------------------------------------------------------------ pieces
    def _grow_tables(self, n: int) -> None:
        """Symmetric-memory gather tables (collective: every rank grows to the same count)."""
        import torch.distributed._symmetric_memory as symm_mem
        while len(self.bufs) < n:
            t = symm_mem.empty(self.shape, dtype=torch.float32, device=self.device)
            t.zero_()
            self.bufs.append(t)
            self.hdls.append(symm_mem.rendezvous(t, self.group))
        self.use_multicast = bool(int(os.environ.get("B200GCN_MULTICAST", "1"))) and not getattr(self, "halo", False) \
            and all(getattr(h, "has_multicast_support", False) and h.multicast_ptr for h in self.hdls)
        torch.cuda.synchronize(self.device)

    def _barrier(self):
        if self.exchange in ("chain", "fused"):
            self.hdls[0].barrier(channel=0)
        else:
            dist.barrier(group=self.group)

    def _block(self, buf: Tensor) -> Tensor:
        r0 = self.rank * self.plan.n_pad
        return buf[r0:r0 + self.n_loc]

    def _publish_ego(self, x0_loc: Tensor, buf_idx: int) -> None:
        """Layer-0 exchange: every rank's own rows into everybody's gather table."""
        plan = self.plan
        if self.exchange == "fused":
            from .functional import spmm_raw
            # identity mode of the SpMM kernel: p = x0_loc, epilogue stores it into every rank's table
            spmm_raw(None, x0_loc, peers=self._peers(buf_idx))
            self.hdls[buf_idx].barrier(channel=0)
        else:
            pad = torch.zeros(plan.n_pad, self.dim, dtype=torch.float32, device=self.device)
            pad[: self.n_loc] = x0_loc
            dist.all_gather_into_tensor(self.bufs[buf_idx], pad, group=self.group)

    def _peers(self, buf_idx: int):
        from .functional import PeerTables
        if buf_idx not in self._peer:
            hdl = self.hdls[buf_idx]
            mc = int(hdl.multicast_ptr) if self.use_multicast else None
            self._peer[buf_idx] = PeerTables(int(hdl.buffer_ptrs_dev), self.plan.P, self.rank * self.plan.n_pad,
                                             self.dim, mc_ptr=mc, need=self._need)
        return self._peer[buf_idx]

    # ------------------------------------------------------------------ forward: one persistent kernel
    def _forward_chain(self, xu_loc: Tensor, xi_loc: Tensor, n_layers: int) -> Tensor:
        from . import _lib
        from .functional import spmm_args, spmm_chain
        plan, uc, n = self.plan, self.plan.u_cnt[self.rank], self.n_loc
        if 2 + 2 * n_layers > _lib.CHAIN_MAX_PHASES:
            raise ValueError(f"n_layers={n_layers}: the chain kernel takes at most "
                             f"{(_lib.CHAIN_MAX_PHASES - 2) // 2} layers")
        if len(self.bufs) < n_layers:
            self._grow_tables(n_layers)
        acc = torch.empty(n, self.dim, dtype=torch.float32, device=self.device)
        scale = 1.0 / (n_layers + 1)
        rows = {"U": (0, uc), "I": (uc, n)}
        x0 = {"U": xu_loc, "I": xi_loc}
        sched = plan.chain_schedule(n_layers)
        phases = []
        for ph in sched:
            half, l = ph["half"], ph["layer"]
            if ph["kind"] == "publish":                  # identity mode: own ego rows -> T0 of every rank
                phases.append(spmm_args(None, x0[half], peers=self._peers(0), peer_row_offset=rows[half][0]))
                continue
            r0, r1 = rows[half]
            last = l == n_layers
            phases.append(spmm_args(self.handle, self.bufs[l - 1], rows=(r0, r1),
                                    acc_in=x0[half] if l == 1 else acc[r0:r1], acc_out=acc[r0:r1],
                                    acc_scale=scale if last else 1.0, peers=None if last else self._peers(l)))
        wait, wait_local = [ph["wait"] for ph in sched], [ph["wait_local"] for ph in sched]
        self._epoch += 1
        s = _lib.ChainSync()
        s.n_ranks, s.rank, s.epoch, s.start_wait_phase = plan.P, self.rank, self._epoch, self._last_phase
        s.flags, s.flags_peers = self._flags.data_ptr(), int(self._flags_hdl.buffer_ptrs_dev)
        s.scratch = self._scratch.data_ptr()
        for k in range(_lib.CHAIN_MAX_PHASES):
            s.wait_phase[k] = wait[k] if k < len(wait) else -1
            s.wait_local[k] = wait_local[k] if k < len(wait_local) else -1
            s.merge_next[k] = sched[k]["merge_next"] if k < len(sched) else 0
        spmm_chain(phases, s, self.device)
        self._last_phase = len(phases) - 1
        # the inputs are read by the kernel after this call returns: keep them alive on this stream
        xu_loc.record_stream(torch.cuda.current_stream(self.device))
        xi_loc.record_stream(torch.cuda.current_stream(self.device))
        return acc

    def phase_times_us(self) -> List[float]:
        """Diagnostics of the LAST chain launch (synchronises): microseconds from the first tile taken to the moment
        the last CTA of this rank left each phase ([I0, U0, then the half-layers in ``half_layer_order``])."""
        torch.cuda.synchronize(self.device)
        t = self._scratch[16:].view(torch.int64).cpu().tolist()
        return [(v - t[0]) / 1e3 for v in t[1:2 + self._last_phase]]

    # ------------------------------------------------------------------ forward: one launch per layer
    def _forward_split(self, xu_loc: Tensor, xi_loc: Tensor, n_layers: int) -> Tensor:
        """Fused exchange with the bipartite dependency structure exploited: user rows gather only ITEM rows and
        vice versa, so layer 1 of the user rows needs only the item half of the ego tables.  The item half is
        published first; the user half is published on a side stream WHILE layer 1 of the user rows runs."""
        from .functional import spmm_raw
        uc, n = self.plan.u_cnt[self.rank], self.n_loc
        cur_stream = torch.cuda.current_stream(self.device)
        scale = 1.0 / (n_layers + 1)
        last1 = n_layers == 1
        acc = torch.empty(n, self.dim, dtype=torch.float32, device=self.device)
        self.hdls[0].barrier(channel=0)                       # peers done with the previous call's tables
        spmm_raw(None, xi_loc, peers=self._peers(0), peer_row_offset=uc)          # publish ego items
        self.hdls[0].barrier(channel=0)
        ev = torch.cuda.Event()
        ev.record(cur_stream)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ev)
            spmm_raw(None, xu_loc, peers=self._peers(0))                           # publish ego users (side stream)
            ev2 = torch.cuda.Event()
            ev2.record(self.side)
        p1 = None if last1 else self._peers(1)
        spmm_raw(self.handle, self.bufs[0], rows=(0, uc), acc_in=xu_loc, acc_out=acc[:uc],
                 acc_scale=scale if last1 else 1.0, peers=p1)                       # layer 1, user rows
        cur_stream.wait_event(ev2)
        self.hdls[0].barrier(channel=0)
        spmm_raw(self.handle, self.bufs[0], rows=(uc, n), acc_in=xi_loc, acc_out=acc[uc:],
                 acc_scale=scale if last1 else 1.0, peers=p1)                       # layer 1, item rows
        for l in range(2, n_layers + 1):
            self.hdls[(l - 1) % 2].barrier(channel=0)
            last = l == n_layers
            spmm_raw(self.handle, self.bufs[(l - 1) % 2], acc_in=acc, acc_out=acc,
                     acc_scale=scale if last else 1.0, peers=None if last else self._peers(l % 2))
        return acc

    def forward(self, xu_loc: Tensor, xi_loc: Tensor, n_layers: int) -> Tensor:
        plan = self.plan
        if n_layers == 0:
            return torch.cat([xu_loc, xi_loc], 0)
        if self.exchange == "chain":
            return self._forward_chain(xu_loc.contiguous(), xi_loc.contiguous(), n_layers)
        if self.split:
            return self._forward_split(xu_loc.contiguous(), xi_loc.contiguous(), n_layers)
        x0_loc = torch.cat([xu_loc, xi_loc], 0)
        if self.exchange == "fused":
            self.hdls[0].barrier(channel=0)      # peers finished reading the tables of the previous call
        self._publish_ego(x0_loc, 0)
        scale = 1.0 / (n_layers + 1)
        acc = None
        for l in range(1, n_layers + 1):
            cur, nxt = self.bufs[(l - 1) % 2], self.bufs[l % 2]
            last = l == n_layers
            acc_in = x0_loc if l == 1 else acc
            if self._local_spmm is not None:            # CPU tests of the host logic
                y = self._local_spmm(self._edges, cur)
                acc = (acc_in + y) * (scale if last else 1.0)
                if not last:
                    pad = torch.zeros(plan.n_pad, self.dim, dtype=torch.float32, device=self.device)
                    pad[: self.n_loc] = y
                    dist.all_gather_into_tensor(nxt, pad, group=self.group)
                continue
            from .functional import spmm_raw
            if acc is None:
                acc = torch.empty(self.n_loc, self.dim, dtype=torch.float32, device=self.device)
            if self.exchange == "fused":
                spmm_raw(self.handle, cur, acc_in=acc_in, acc_out=acc, acc_scale=scale if last else 1.0,
                         peers=None if last else self._peers(l % 2))
                if not last:
                    self.hdls[l % 2].barrier(channel=0)
            else:
                y = None if last else self._block(nxt)
                spmm_raw(self.handle, cur, y=y, acc_in=acc_in, acc_out=acc, acc_scale=scale if last else 1.0)
                if not last:
                    dist.all_gather_into_tensor(nxt, nxt[self.rank * plan.n_pad:(self.rank + 1) * plan.n_pad],
                                                group=self.group)
        return acc

    # ------------------------------------------------------------------ autograd
    def propagate(self, xu_loc: Tensor, xi_loc: Tensor, n_layers: int) -> Tuple[Tensor, Tensor]:
        """Differentiable ``LightGCN.forward`` of this rank's rows: ``(user_rows, item_rows)``.  All ranks must
        call forward AND backward collectively (the backward is the same exchange on the gradients)."""
        return _ShardedLayerMean.apply(xu_loc, xi_loc, self, int(n_layers))


class _ShardedLayerMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xu_loc, xi_loc, prop: ShardedPropagator, n_layers: int):
        out = prop.forward(xu_loc.detach(), xi_loc.detach(), n_layers)
        ctx.prop, ctx.n_layers, ctx.uc = prop, n_layers, xu_loc.size(0)
        return out[: ctx.uc], out[ctx.uc:]

    @staticmethod
    def backward(ctx, gu, gi):
        prop, uc = ctx.prop, ctx.uc
        D, dev = prop.dim, prop.device
        gu = torch.zeros(uc, D, device=dev) if gu is None else gu.contiguous()
        gi = torch.zeros(prop.n_loc - uc, D, device=dev) if gi is None else gi.contiguous()
        g = prop.forward(gu, gi, ctx.n_layers)           # M is symmetric: dL/dx0 = M dL/dout, row-sharded alike
        return g[:uc], g[uc:], None, None


# ---------------------------------------------------------------------------------------------------- bench
def _pin_to_gpu_numa_node(local: int) -> Optional[str]:
    """Best effort: run this rank's host thread (and thus its pinned-buffer first touches) on the CPUs next to
    its GPU, so that the PCIe copies of the end-to-end leg stay NUMA-local."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus near gpu {local}"
    except Exception as e:  # pragma: no cover
        return f"unpinned ({type(e).__name__})"
    return "unpinned"


class HostPipeline:
    """Per-rank three-stream pipeline (copy-in / compute / copy-out, ``depth`` slots) around
    ``ShardedPropagator.forward`` for callers that keep the tables in pinned HOST memory: the PCIe transfers of
    step k+1 and step k-1 ride under the kernel of step k.  Every rank must submit the same number of steps."""

    def __init__(self, prop: ShardedPropagator, n_layers: int, depth: int = 2):
        self.prop, self.L, self.depth = prop, int(n_layers), int(depth)
        dev, uc, n, D = prop.device, prop.plan.u_cnt[prop.rank], prop.n_loc, prop.dim
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.slots = [{"du": torch.empty(uc, D, device=dev), "di": torch.empty(n - uc, D, device=dev),
                       "in_ready": torch.cuda.Event(), "cmp_done": torch.cuda.Event(),
                       "out_done": torch.cuda.Event()} for _ in range(self.depth)]
        self.k = 0

    def submit(self, hu: Tensor, hi: Tensor, h_out: Tensor) -> None:
        s = self.slots[self.k % self.depth]
        first = self.k < self.depth
        self.k += 1
        with torch.cuda.stream(self.s_in):
            if not first:
                self.s_in.wait_event(s["cmp_done"])
            s["du"].copy_(hu, non_blocking=True)
            s["di"].copy_(hi, non_blocking=True)
            s["in_ready"].record(self.s_in)
        with torch.cuda.stream(self.s_cmp):
            self.s_cmp.wait_event(s["in_ready"])
            with torch.no_grad():
                out = self.prop.forward(s["du"], s["di"], self.L)
            s["cmp_done"].record(self.s_cmp)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(s["cmp_done"])
            h_out.copy_(out, non_blocking=True)
            out.record_stream(self.s_out)
            s["out_done"].record(self.s_out)

    def synchronize(self) -> None:
        for st in (self.s_in, self.s_cmp, self.s_out):
            st.synchronize()
