"""Row-sharded multi-GPU propagation: one process per GPU, embedding tables and CSR rows split 1-D over the
ranks of one NVLink/NVSwitch box, one exchange per layer.  The reference has no multi-GPU
path at all (single process, `quick_start.py:41`); this is the engine's extension of
`LightGCN.forward` (lightgcn.py:70-81) to P ranks with identical numerics.

Partition (``ShardPlan``): rank p owns users [pU/P, (p+1)U/P) and items [pI/P, (p+1)I/P) — both halves of
the bipartite graph stay balanced.  Node ids are relabelled rank-major: ``new = p * n_pad + local`` with the
rank's users first, then its items, so the gathered table of a layer is the plain concatenation of the
ranks' output blocks and every rank's block is one contiguous slab.

Exchange per layer, two implementations with the same result:

* ``"fused"`` (CUDA default): the SpMM epilogue itself stores every finished row into the next-layer table
  of EVERY rank over NVLink — peer-mapped symmetric memory (``st.global`` to P pointers) or one
  ``multimem.st`` through the NVSwitch multicast address — so the transfer overlaps the gathers row by row
  inside the one kernel; a symmetric-memory barrier orders the layers.
* ``"allgather"``: local SpMM into the own block, then ``all_gather_into_tensor`` (NCCL on GPUs; gloo in the
  CPU tests of the host logic, where the local product is injected by the test).
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


class ShardedPropagator:
    """K-layer LightGCN propagation of one rank.  ``forward(xu_loc, xi_loc, L)`` returns the rank's rows
    ``[n_loc, D]`` (own users then own items) of ``mean(x_0 .. x_L)``."""

    def __init__(self, plan: ShardPlan, rank: int, dst_local: Tensor, src_new: Tensor, w: Tensor, dim: int,
                 device, group=None, exchange: Optional[str] = None,
                 local_spmm: Optional[Callable[[Tensor, Tensor], Tensor]] = None):
        self.plan, self.rank, self.dim = plan, rank, int(dim)
        self.device = torch.device(device)
        self.group = group if group is not None else dist.group.WORLD
        self.n_loc = plan.n_loc[rank]
        is_cuda = self.device.type == "cuda"
        if exchange is None:
            exchange = os.environ.get("B200GCN_EXCHANGE", "fused" if is_cuda else "allgather")
        if exchange not in ("fused", "fused-split", "allgather"):
            raise ValueError("exchange must be 'fused', 'fused-split' or 'allgather'")
        self.split = exchange == "fused-split"
        if self.split:
            exchange = "fused"
        if exchange == "fused" and not is_cuda:
            raise RuntimeError("the fused exchange needs CUDA peer memory")
        self.exchange = exchange
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
        shape = (plan.n_full, self.dim)
        self._peer = [None, None]
        if exchange == "fused":
            import torch.distributed._symmetric_memory as symm_mem
            self.bufs, self.hdls = [], []
            for _ in range(2):
                t = symm_mem.empty(shape, dtype=torch.float32, device=self.device)
                hdl = symm_mem.rendezvous(t, self.group)
                self.bufs.append(t)
                self.hdls.append(hdl)
            self.use_multicast = bool(int(os.environ.get("B200GCN_MULTICAST", "1"))) and all(
                getattr(h, "has_multicast_support", False) and h.multicast_ptr for h in self.hdls)
        else:
            self.bufs = [torch.zeros(shape, dtype=torch.float32, device=self.device) for _ in range(2)]
        for b in self.bufs:
            b.zero_()
        self.acc = torch.empty(self.n_loc, self.dim, dtype=torch.float32, device=self.device)
        if self.split and (self.handle is None or self.handle._n_hubs > 0):
            self.split = False          # row-split launches need a hub-free graph
        self.side = torch.cuda.Stream(self.device) if self.split else None
        self._barrier()

    # ------------------------------------------------------------------ pieces
    def _barrier(self):
        if self.exchange == "fused":
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
            if os.environ.get("B200GCN_PUBLISH", "kernel") == "copy":
                # copy engines: one peer copy per rank, spread over a few side streams (A/B against the kernel)
                hdl = self.hdls[buf_idx]
                r0 = self.rank * plan.n_pad
                cur = torch.cuda.current_stream(self.device)
                if not hasattr(self, "_cstreams"):
                    self._cstreams = [torch.cuda.Stream(self.device) for _ in range(4)]
                ev = torch.cuda.Event()
                ev.record(cur)
                for q in range(plan.P):
                    st = self._cstreams[q % len(self._cstreams)]
                    st.wait_event(ev)
                    with torch.cuda.stream(st):
                        peer = hdl.get_buffer((self.rank + q) % plan.P, (plan.n_full, self.dim), torch.float32)
                        peer[r0:r0 + self.n_loc].copy_(x0_loc, non_blocking=True)
                for st in self._cstreams:
                    e2 = torch.cuda.Event()
                    e2.record(st)
                    cur.wait_event(e2)
            else:
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
        if self._peer[buf_idx] is None:
            hdl = self.hdls[buf_idx]
            mc = int(hdl.multicast_ptr) if self.use_multicast else None
            self._peer[buf_idx] = PeerTables(int(hdl.buffer_ptrs_dev), self.plan.P, self.rank * self.plan.n_pad,
                                             self.dim, mc_ptr=mc)
        return self._peer[buf_idx]

    # ------------------------------------------------------------------ forward
    def _forward_split(self, xu_loc: Tensor, xi_loc: Tensor, n_layers: int) -> Tensor:
        """Fused exchange with the bipartite dependency structure exploited: user rows gather only ITEM rows and
        vice versa, so layer 1 of the user rows needs only the item half of the ego tables.  The item half is
        published first; the user half is published on a side stream WHILE layer 1 of the user rows runs."""
        from .functional import spmm_raw
        plan, uc, n = self.plan, self.plan.u_cnt[self.rank], self.n_loc
        cur_stream = torch.cuda.current_stream(self.device)
        scale = 1.0 / (n_layers + 1)
        last1 = n_layers == 1
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
        spmm_raw(self.handle, self.bufs[0], rows=(0, uc), acc_in=xu_loc, acc_out=self.acc[:uc],
                 acc_scale=scale if last1 else 1.0, peers=p1)                       # layer 1, user rows
        cur_stream.wait_event(ev2)
        self.hdls[0].barrier(channel=0)
        spmm_raw(self.handle, self.bufs[0], rows=(uc, n), acc_in=xi_loc, acc_out=self.acc[uc:],
                 acc_scale=scale if last1 else 1.0, peers=p1)                       # layer 1, item rows
        for l in range(2, n_layers + 1):
            self.hdls[(l - 1) % 2].barrier(channel=0)
            last = l == n_layers
            spmm_raw(self.handle, self.bufs[(l - 1) % 2], acc_in=self.acc, acc_out=self.acc,
                     acc_scale=scale if last else 1.0, peers=None if last else self._peers(l % 2))
        return self.acc

    def forward(self, xu_loc: Tensor, xi_loc: Tensor, n_layers: int) -> Tensor:
        plan = self.plan
        if self.split and n_layers >= 1:
            return self._forward_split(xu_loc.contiguous(), xi_loc.contiguous(), n_layers)
        x0_loc = torch.cat([xu_loc, xi_loc], 0)
        if n_layers == 0:
            return x0_loc
        if self.exchange == "fused":
            self.hdls[0].barrier(channel=0)      # peers finished reading the tables of the previous call
        self._publish_ego(x0_loc, 0)
        scale = 1.0 / (n_layers + 1)
        for l in range(1, n_layers + 1):
            cur, nxt = self.bufs[(l - 1) % 2], self.bufs[l % 2]
            last = l == n_layers
            acc_in = x0_loc if l == 1 else self.acc
            if self._local_spmm is not None:            # CPU tests of the host logic
                y = self._local_spmm(self._edges, cur)
                self.acc = (acc_in + y) * (scale if last else 1.0)
                if not last:
                    pad = torch.zeros(plan.n_pad, self.dim, dtype=torch.float32, device=self.device)
                    pad[: self.n_loc] = y
                    dist.all_gather_into_tensor(nxt, pad, group=self.group)
                continue
            from .functional import spmm_raw
            if self.exchange == "fused":
                spmm_raw(self.handle, cur, acc_in=acc_in, acc_out=self.acc, acc_scale=scale if last else 1.0,
                         peers=None if last else self._peers(l % 2))
                if not last:
                    self.hdls[l % 2].barrier(channel=0)
            else:
                y = None if last else self._block(nxt)
                spmm_raw(self.handle, cur, y=y, acc_in=acc_in, acc_out=self.acc, acc_scale=scale if last else 1.0)
                if not last:
                    dist.all_gather_into_tensor(nxt, nxt[self.rank * plan.n_pad:(self.rank + 1) * plan.n_pad],
                                                group=self.group)
        return self.acc


# ---------------------------------------------------------------------------------------------------- bench
def bench_entry(args, rank: int, world: int, local: int) -> None:
    """`bench.py --gpus N` under torchrun: STRONG scaling — the same cfg2 graph (BASELINE.json configs[1])
    split over N ranks; value = total directed edges x L / max-over-ranks device time."""
    import time

    import bench as B

    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    U, I, E, D, L = B.WORKLOADS[args.workload]
    N, nnz = U + I, 2 * E
    plan = ShardPlan(U, I, world)
    t0 = time.perf_counter()
    uid, iid = B.synth_graph_device(U, I, E, dev)              # same seed on every rank -> same graph
    w = interaction_weights_device(uid, iid, U, I)
    dst, src, wl = plan.local_edges(rank, uid, iid, w)
    del uid, iid, w
    prop = ShardedPropagator(plan, rank, dst, src, wl, D, dev)
    del dst, src, wl
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    xu, xi = B.xavier_tables_device(U, I, D, dev)              # same seed: every rank slices its rows
    xu_loc, xi_loc = (t.contiguous() for t in plan.scatter_tables(rank, xu, xi))
    del xu, xi
    torch.cuda.empty_cache()

    from .functional import LaunchTimer

    with torch.no_grad():
        for _ in range(args.warmup):
            out = prop.forward(xu_loc, xi_loc, L)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        timer = LaunchTimer()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with B.ClockSampler(local) as clocks:
            with timer:
                start.record()
                for _ in range(args.steps):
                    out = prop.forward(xu_loc, xi_loc, L)
                end.record()
            torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([start.elapsed_time(end)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_step = ms.item() / args.steps
        launch_ms = timer.durations_ms()
        per_step = len(launch_ms) // args.steps
        by_pos = [sum(launch_ms[i::per_step]) / args.steps for i in range(per_step)] if per_step else []
        k_ms = torch.tensor([sum(launch_ms) / len(launch_ms)], device=dev)
        dist.all_reduce(k_ms, op=dist.ReduceOp.MAX)

        # e2e: pinned host slices in, pinned host result out, inside the timed region
        hu, hi = xu_loc.cpu().pin_memory(), xi_loc.cpu().pin_memory()
        ho = torch.empty(prop.n_loc, D).pin_memory()
        du, di = torch.empty_like(xu_loc), torch.empty_like(xi_loc)

        def e2e_step():
            du.copy_(hu, non_blocking=True)
            di.copy_(hi, non_blocking=True)
            ho.copy_(prop.forward(du, di, L), non_blocking=True)

        for _ in range(2):
            e2e_step()
        torch.cuda.synchronize()
        dist.barrier()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2e_steps = max(3, args.steps // 2)
        s2.record()
        for _ in range(e2e_steps):
            e2e_step()
        e2.record()
        torch.cuda.synchronize()
        e2e_ms = torch.tensor([s2.elapsed_time(e2) / e2e_steps], device=dev)
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)

    if rank == 0:
        peak, peak_src = B.measured_peak_gbs()
        n_loc = plan.n_loc[0]
        b_layer_rank = (nnz // world) * (4 * D + 8) + n_loc * (4 * D + 4)     # per-rank algorithmic bytes/launch
        achieved = b_layer_rank / (k_ms.item() * 1e-3) / 1e9
        line = {
            "metric": B.METRIC, "value": nnz * L / (ms_step * 1e-3), "unit": B.UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: LightGCN propagation U={U} I={I} E={E} (nnz={nnz}) D={D} L={L}",
                       "parallelism": f"row-sharded x{world}, exchange={prop.exchange}"
                                      + ("-split" if getattr(prop, "split", False) else "")
                                      + (" (multimem.st multicast)" if getattr(prop, "use_multicast", False) else
                                         " (peer st.global)" if prop.exchange == "fused" else " (NCCL all-gather)"),
                       "l2": "inputs larger than L2; no flush", "csr_build_s": round(build_s, 3)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "spmm_warp_kernel (per rank; includes the fused NVLink stores)",
                         "algorithmic_bytes_per_launch": b_layer_rank, "launch_ms_mean": k_ms.item(),
                         "launch_ms_by_position_in_step_rank0": [round(v, 4) for v in by_pos],
                         "peak_source": peak_src,
                         "nvlink_bytes_out_per_launch": n_loc * D * 4 * (1 if getattr(prop, "use_multicast", False)
                                                                          else world - 1)},
            "cpu_baseline": None,
            "e2e": {"value": nnz * L / (e2e_ms.item() * 1e-3), "unit": B.UNIT, "ms_per_step": e2e_ms.item(),
                    "h2d_bytes_per_step": N * D * 4, "d2h_bytes_per_step": N * D * 4,
                    "api": "ShardedPropagator.forward on pinned host slices -> pinned host rows (all ranks)"},
            "gpu_launches": timer.count,
            "clocks": clocks.summary(),
        }
        B.emit(line)
    dist.barrier()
    dist.destroy_process_group()
