#!/usr/bin/env python
"""Benchmark of the hot path: LightGCN K-layer normalised-adjacency propagation (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|small|...]

A *step* is one full ``LightGCN.forward`` (lightgcn.py:70-81): L SpMM layers with the layer mean fused,
over one synthetic user-item graph.  Metric: directed edge traversals per second = nnz * L / t_step
(SURVEY §8d), whole job.  Prints ONE JSON line on rank 0.

Workload (N = 1): BASELINE.json configs[1] — U = I = 1 M (incl. [PAD]), E = 100 M interactions
(nnz = 200 M), D = 64, L = 3, fp32, generated on the device with a seeded torch CUDA generator.
Inputs (512 MB table + 1.6 GB index stream) are far larger than the 126 MB L2, so no flush between
iterations is needed.  N > 1: see recbole_gnn_b200/sharded.py (row-sharded, per-layer all-gather).

`--impl reference` times the reference's CPU path restated by the oracle (torch.sparse.mm on the
normalised adjacency, oracle/oracle.py) on the full graph of the same workload (a row slice only if the host
cannot fit it in the time budget), on the host cores, thread count swept.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# keep stdout to the one JSON line: NCCL's own "NCCL version ..." banner (printed whenever NCCL_DEBUG is set
# on the box) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import torch  # noqa: E402

WORKLOADS = {
    # name: (user_num, item_num, n_inter, dim, n_layers)
    "cfg2": (1_000_000, 1_000_000, 100_000_000, 64, 3),      # BASELINE.json configs[1] (the metric's config)
    "cfg5": (10_000_000, 10_000_000, 1_000_000_000, 128, 3), # BASELINE.json configs[4] (row-sharded x8)
    "cfg5_1gpu": (10_000_000, 10_000_000, 1_000_000_000, 128, 3),
    "medium": (200_000, 200_000, 20_000_000, 64, 3),
    "small": (20_000, 20_000, 1_000_000, 64, 3),
}
METRIC = "lightgcn_3layer_propagation_edges_per_sec"
UNIT = "edges/s"


_REAL_STDOUT = None


def capture_stdout():
    """Keep the process's real stdout for the ONE JSON line: everything libraries print on fd 1 (NCCL's
    "NCCL version ..." banner ignores NCCL_DEBUG_FILE on this box) is sent to stderr instead."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def algorithmic_bytes_per_layer(nnz: int, n: int, d: int) -> int:
    """SURVEY §8d no-reuse gather model: per directed edge one neighbour row + int32 col + fp32 val,
    per node one output row + one rowptr entry."""
    return nnz * (4 * d + 4 + 4) + n * (4 * d + 4)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nme in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- ours
def synth_graph_device(U, I, E, dev, seed=0, chunk=50_000_000):
    gen = torch.Generator(device=dev).manual_seed(seed)
    uid = torch.empty(E, dtype=torch.int64, device=dev)
    iid = torch.empty(E, dtype=torch.int64, device=dev)
    for s in range(0, E, chunk):
        n = min(chunk, E - s)
        uid[s:s + n] = torch.randint(1, U, (n,), generator=gen, device=dev)
        iid[s:s + n] = torch.randint(1, I, (n,), generator=gen, device=dev)
    return uid, iid


def xavier_tables_device(U, I, D, dev, seed=1):
    gen = torch.Generator(device=dev).manual_seed(seed)
    bu, bi = (6.0 / (U + D)) ** 0.5, (6.0 / (I + D)) ** 0.5      # xavier_uniform_ on [rows, D] (lightgcn.py:57)
    xu = (torch.rand(U, D, generator=gen, device=dev) * 2 - 1) * bu
    xi = (torch.rand(I, D, generator=gen, device=dev) * 2 - 1) * bi
    return xu, xi


def cpu_slice_baseline(h, xu, xi, n_rows_sample, threads, reps=3):
    """The oracle (torch.sparse.mm, CSR layout = what torch_sparse.matmul runs on CPU) on rows
    [0, n_rows_sample) of the SAME graph, gathering from the full table; also a live parity check."""
    from oracle import oracle as O

    rowptr, col, val = h.csr()
    N = h.size(0)
    e_end = int(rowptr[n_rows_sample].item())
    crow = rowptr[: n_rows_sample + 1].cpu()
    ccol = col[:e_end].cpu().to(torch.int64)
    cval = val[:e_end].cpu()
    x = torch.cat([xu, xi]).cpu()
    torch.set_num_threads(threads)
    if x.numel() >= 2 ** 31:
        # MKL's 32-bit indexing segfaults on a dense operand with >= 2^31 elements (config 5: 20 M x 128):
        # gather only the referenced rows into a compact table (same arithmetic, same row order per entry)
        uniq, ccol = torch.unique(ccol, return_inverse=True)
        x, N = x[uniq].contiguous(), uniq.numel()
    a_csr = torch.sparse_csr_tensor(crow, ccol, cval, size=(n_rows_sample, N))
    O.propagate_sparse(a_csr, x)                      # warm-up
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        y = O.propagate_sparse(a_csr, x)
        ts.append(time.perf_counter() - t0)
    t = statistics.median(ts)
    return e_end / t, e_end, y, t


def run_ours(args):
    import recbole_gnn_b200 as rg
    from recbole_gnn_b200 import functional as F_

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        return run_sharded(args, rank, world, local)

    U, I, E, D, L = WORKLOADS[args.workload]
    N, nnz = U + I, 2 * E
    t0 = time.perf_counter()
    uid, iid = synth_graph_device(U, I, E, dev)
    h = rg.GraphHandle.from_interactions(uid, iid, U, I).gcn_norm().to(dev)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    del uid, iid
    torch.cuda.empty_cache()
    xu, xi = xavier_tables_device(U, I, D, dev)

    def step():
        return F_.lightgcn_propagate(h, xu, xi, L)

    with torch.no_grad():
        for _ in range(args.warmup):
            out = step()
        torch.cuda.synchronize()
        # ---- timed region: device-resident inputs, per-launch events on the launching stream ----------
        timer = F_.LaunchTimer()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clocks:
            torch.cuda.synchronize()
            with timer:
                start.record()
                for _ in range(args.steps):
                    out = step()
                end.record()
            torch.cuda.synchronize()
        ms_total = start.elapsed_time(end)
        ms_step = ms_total / args.steps
        launch_ms = timer.durations_ms()
        n_launches = timer.count

        # ---- e2e: the public host-buffer API (recbole_gnn_b200.host.HostPropagator): every step copies its
        #      two tables from pinned host memory, propagates, and copies its result back to pinned host
        #      memory, all inside the timed region; consecutive steps overlap on copy-in/compute/copy-out
        #      streams.  Timed on the host clock around submit..synchronize (three streams are involved).
        from recbole_gnn_b200.host import HostPropagator
        hu, hi = xu.cpu().pin_memory(), xi.cpu().pin_memory()
        outs = [(torch.empty(U, D).pin_memory(), torch.empty(I, D).pin_memory()) for _ in range(2)]
        hp = HostPropagator(h, U, I, D, L, depth=2)
        for k in range(3):
            hp.submit(hu, hi, *outs[k % 2])
        hp.synchronize()
        torch.cuda.synchronize()
        e2e_steps = args.steps
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s2.record()
        for k in range(e2e_steps):
            hp.submit(hu, hi, *outs[k % 2])
        hp.synchronize()
        e2.record()
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        assert torch.equal(outs[(e2e_steps - 1) % 2][0], out[0].cpu())   # the e2e route returns the same numbers
        # un-overlapped single call (copy-in -> propagate -> copy-out on one stream), for reference
        du, di = torch.empty_like(xu), torch.empty_like(xi)
        s3, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s3.record()
        for _ in range(3):
            du.copy_(hu, non_blocking=True); di.copy_(hi, non_blocking=True)
            u_, i_ = F_.lightgcn_propagate(h, du, di, L)
            outs[0][0].copy_(u_, non_blocking=True); outs[0][1].copy_(i_, non_blocking=True)
        e3.record()
        torch.cuda.synchronize()
        e2e_serial_ms = s3.elapsed_time(e3) / 3

    edges_per_step = nnz * L
    value = edges_per_step / (ms_step * 1e-3)
    peak, peak_src = measured_peak_gbs()
    b_layer = algorithmic_bytes_per_layer(nnz, N, D)
    k_ms = statistics.mean(launch_ms)
    achieved = b_layer / (k_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    # ---- CPU baseline (oracle port) on a bounded row slice of the same graph + live parity check --------
    threads = os.cpu_count() or 1
    cpu = None
    if not args.no_cpu_baseline:
        n_sample = min(U - 1, max(1000, int(args.cpu_sample_rows)))
        x_full = torch.cat([xu, xi])
        y_dev = torch.empty(N, D, device=dev)
        F_.spmm_raw(h, x_full, y=y_dev)
        eps_cpu, e_cnt, y_cpu, t_cpu = cpu_slice_baseline(h, xu, xi, n_sample, threads)
        d = (y_dev[:n_sample].cpu() - y_cpu).abs().max().item()
        scale = y_cpu.abs().max().item()
        assert d < 1e-4 and d / scale < 1e-5, ("full-size parity failed", d, d / scale)
        cpu = {"value": eps_cpu, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"rows [0,{n_sample}) of the {args.workload} graph = {e_cnt} of {nnz} directed edges, one "
                         f"layer, gathers from the full {N}x{D} table; torch.sparse.mm CSR (oracle.propagate_sparse), "
                         f"median of 3, {t_cpu:.3f} s; GPU rows match to {d:.2e} abs / {d / scale:.2e} scaled"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: LightGCN propagation U={U} I={I} E={E} (nnz={nnz}) D={D} L={L}",
                   "graph": "uniform random bipartite, ids in [1, n), seeded torch CUDA generator",
                   "l2": "inputs (table %d MB + index stream %d MB) larger than the 126 MB L2; no flush" %
                         (N * D * 4 // 2 ** 20, nnz * 8 // 2 ** 20),
                   "csr_build_s": round(build_s, 3), "parallelism": "1 gpu"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "spmm_warp_kernel<16,8,true,*> (v2: warp per row)",
                     "algorithmic_bytes_per_launch": b_layer, "launch_ms_mean": k_ms,
                     "launch_ms_min": min(launch_ms), "launch_ms_max": max(launch_ms),
                     "launch_ms_by_layer": [round(sum(launch_ms[i::L]) / args.steps, 4) for i in range(L)],
                     "peak_source": peak_src},
        "cpu_baseline": cpu,
        "e2e": {"value": edges_per_step / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": N * D * 4, "d2h_bytes_per_step": N * D * 4,
                "serial_ms_per_step": e2e_serial_ms,
                "api": "host.HostPropagator.submit(pinned tables) -> pinned result; 3 streams, depth 2; "
                       "serial_ms_per_step = the same without overlap"},
        "gpu_launches": n_launches,
        "clocks": clocks.summary(),
    }
    emit(line)


# ------------------------------------------------------------------------------------------- N > 1 arm
def sampled_row_parity(prop, xu_loc, xi_loc, out, n_layers: int,
                       n_sample: int, seed: int = 0):
    """Inductive parity check of a chain-mode forward against the CPU oracle on ``n_sample`` of this rank's rows:
    layer l of the sampled rows is recomputed by ``oracle.propagate_sparse`` (torch.sparse.mm, CSR) from the
    gathered layer-(l-1) table the GPUs produced, and compared with what the GPU wrote for layer l; the last step
    compares the final mean.  Returns (max abs, max scaled) over all layers.  The checker of the N > 1 arm (the only place besides the
    cpu_baseline leg where bench.py touches oracle/)."""
    from oracle import oracle as O
    plan, rank, n, D = prop.plan, prop.rank, prop.n_loc, prop.dim
    g = torch.Generator().manual_seed(seed + rank)
    rows = torch.randperm(n, generator=g)[: min(n_sample, n)].sort().values.to(prop.device)
    rowptr, col, val = prop.handle.csr()
    beg, end = rowptr[rows], rowptr[rows + 1]
    cnt = end - beg
    ptr = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=prop.device)
    ptr[1:] = torch.cumsum(cnt, 0)
    idx = torch.arange(int(ptr[-1]), device=prop.device) - torch.repeat_interleave(ptr[:-1], cnt) + \
        torch.repeat_interleave(beg, cnt)
    c, v = col[idx].to(torch.int64), val[idx]
    uniq, inv = torch.unique(c, return_inverse=True)
    a = torch.sparse_csr_tensor(ptr.cpu(), inv.cpu(), v.cpu(), size=(rows.numel(), uniq.numel()))
    r0 = rank * plan.n_pad
    x0_loc = torch.cat([xu_loc, xi_loc])
    max_abs = max_scaled = 0.0
    acc = x0_loc[rows].cpu().clone()
    for l in range(1, n_layers + 1):
        ref = O.propagate_sparse(a, prop.bufs[l - 1][uniq].cpu())
        acc += ref
        if l < n_layers:
            got = prop.bufs[l][r0 + rows].cpu()
        else:
            got, ref = out[rows].cpu(), acc / (n_layers + 1)
        d = (got - ref).abs().max().item()
        max_abs, max_scaled = max(max_abs, d), max(max_scaled, d / max(ref.abs().max().item(), 1e-30))
        if l < n_layers:
            acc += got - ref          # continue from the GPU's own layer: the check is per layer (inductive)
    return max_abs, max_scaled



def run_sharded(args, rank: int, world: int, local: int) -> None:
    """`bench.py --gpus N` under torchrun: STRONG scaling — the same graph (BASELINE.json configs[1] by default,
    configs[4] with --workload cfg5) split over N ranks; value = total directed edges x L / max-over-ranks device
    time of one step (= ONE chain-kernel launch per rank)."""
    import torch.distributed as dist
    from recbole_gnn_b200 import sharded as S
    from recbole_gnn_b200.functional import LaunchTimer

    numa = S._pin_to_gpu_numa_node(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    U, I, E, D, L = WORKLOADS[args.workload]
    N, nnz = U + I, 2 * E
    plan = S.ShardPlan(U, I, world)
    t0 = time.perf_counter()
    dst, src, wl = S.synth_local_edges(plan, rank, E, dev)       # same seed on every rank -> same graph, own rows only
    prop = S.ShardedPropagator(plan, rank, dst, src, wl, D, dev)
    nnz_loc = prop.handle.nnz()
    del dst, src, wl
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    xu, xi = xavier_tables_device(U, I, D, dev)              # same seed: every rank slices its rows
    xu_loc, xi_loc = (t.contiguous() for t in plan.scatter_tables(rank, xu, xi))
    del xu, xi
    torch.cuda.empty_cache()


    with torch.no_grad():
        for _ in range(args.warmup):
            out = prop.forward(xu_loc, xi_loc, L)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        timer = LaunchTimer()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clocks:
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
        k_ms = torch.tensor([sum(launch_ms) / args.steps], device=dev)        # kernel time per STEP on this rank
        dist.all_reduce(k_ms, op=dist.ReduceOp.MAX)

        # ---- parity, every rank, inside the bench run: sampled rows of every layer against the CPU oracle
        par = torch.zeros(2, device=dev)
        if prop.exchange == "chain" and not args.no_cpu_baseline:
            out = prop.forward(xu_loc, xi_loc, L)
            torch.cuda.synchronize()
            dist.barrier()
            a_, s_ = sampled_row_parity(prop, xu_loc, xi_loc, out, L, args.parity_rows)
            assert a_ < 1e-4 and s_ < 1e-5, ("sharded parity failed", rank, a_, s_)
            par = torch.tensor([a_, s_], device=dev)
        dist.all_reduce(par, op=dist.ReduceOp.MAX)

        # ---- e2e: pinned host slices in, pinned host result out, inside the timed region, three streams per rank
        hu, hi = xu_loc.cpu().pin_memory(), xi_loc.cpu().pin_memory()
        hos = [torch.empty(prop.n_loc, D).pin_memory() for _ in range(2)]
        pipe = S.HostPipeline(prop, L, depth=2)
        for k in range(3):
            pipe.submit(hu, hi, hos[k % 2])
        pipe.synchronize()
        torch.cuda.synchronize()
        dist.barrier()
        e2e_steps = args.steps
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            pipe.submit(hu, hi, hos[k % 2])
        pipe.synchronize()
        torch.cuda.synchronize()
        e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / e2e_steps], device=dev)
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        e2e_same = torch.equal(hos[(e2e_steps - 1) % 2], out.cpu()) if prop.exchange == "chain" else True
        assert e2e_same, "the host-buffer route must return the same numbers"

    phase_us = None
    copy_only_ms = None
    with torch.no_grad():
        if prop.exchange == "chain":
            # stamps of the LAST of six back-to-back launches (steady state; an isolated launch shows launch skew)
            for _ in range(6):
                prop.forward(xu_loc, xi_loc, L)
            phase_us = [round(v, 1) for v in prop.phase_times_us()]
        # the host<->device copies of the e2e leg alone, all ranks at once: the ceiling the pipeline can approach
        dist.barrier()
        du, di, dout = torch.empty_like(xu_loc), torch.empty_like(xi_loc), torch.empty(prop.n_loc, D, device=dev)
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            with torch.cuda.stream(s_in):
                du.copy_(hu, non_blocking=True)
                di.copy_(hi, non_blocking=True)
            with torch.cuda.stream(s_out):
                hos[k % 2].copy_(dout, non_blocking=True)
        torch.cuda.synchronize()
        cm = torch.tensor([(time.perf_counter() - t0) * 1e3 / e2e_steps], device=dev)
        dist.all_reduce(cm, op=dist.ReduceOp.MAX)
        copy_only_ms = cm.item()
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        n_loc = plan.n_loc[0]
        b_step_rank = L * (nnz_loc * (4 * D + 8) + n_loc * (4 * D + 4))     # per-rank algorithmic bytes per step
        achieved = b_step_rank / (k_ms.item() * 1e-3) / 1e9
        mc = bool(getattr(prop, "use_multicast", False))
        nv_in = L * (world - 1) * plan.n_pad * D * 4
        line = {
            "metric": METRIC, "value": nnz * L / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: LightGCN propagation U={U} I={I} E={E} (nnz={nnz}) D={D} L={L}",
                       "parallelism": f"row-sharded x{world}, exchange={prop.exchange}"
                                      + ("-split" if getattr(prop, "split", False) else "")
                                      + (" (multimem.st multicast)" if mc else
                                         " (peer st.global)" if prop.exchange != "allgather" else " (NCCL all-gather)"),
                       "l2": "inputs larger than L2; no flush", "csr_build_s": round(build_s, 3),
                       "host_numa": numa},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None,
                         "kernel": ("spmm_chain_kernel: ONE persistent launch per step per rank = 2 ego publishes + "
                                    f"{2 * L} half-layer SpMM phases with the NVLink stores fused"
                                    if prop.exchange == "chain" else "spmm_warp_kernel (per rank, fused NVLink stores)"),
                         "algorithmic_bytes_per_launch": b_step_rank, "launch_ms_mean": k_ms.item(),
                         "launch_ms_by_position_in_step_rank0": [round(v, 4) for v in by_pos],
                         "peak_source": peak_src,
                         "phase_end_us_rank0": phase_us,
                         "nvlink_bytes_in_per_step": nv_in,
                         "nvlink_in_gbs": nv_in / (ms_step * 1e-3) / 1e9},
            "parity": {"rows_per_rank": args.parity_rows, "max_abs": par[0].item(), "max_scaled": par[1].item(),
                       "what": "every rank: sampled rows of EVERY layer + the final mean vs oracle.propagate_sparse "
                               "(torch.sparse.mm CSR, CPU) fed with the previous layer's gathered table; asserted "
                               "< 1e-4 abs and < 1e-5 scaled"} if prop.exchange == "chain" and not args.no_cpu_baseline
            else None,
            "parity_max_scaled": par[1].item(),
            "cpu_baseline": None,
            "e2e": {"value": nnz * L / (e2e_ms.item() * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms.item(),
                    "h2d_bytes_per_step": N * D * 4, "d2h_bytes_per_step": N * D * 4,
                    "copies_only_ms_per_step": copy_only_ms,
                    "host_copy_gbs_per_direction_all_ranks": N * D * 4 / (copy_only_ms * 1e-3) / 1e9,
                    "api": "sharded.HostPipeline.submit(pinned host slices) -> pinned host rows on every rank; "
                           "3 streams, depth 2; host clock, max over ranks; copies_only = the same H2D + D2H traffic with no kernel (the ceiling of this host)"},
            "gpu_launches": timer.count,
            "clocks": clocks.summary(),
        }
        emit(line)
    dist.barrier()
    dist.destroy_process_group()


# -------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's CPU propagation, restated by the oracle, on the SAME workload configuration: the adjacency built
    as dataset.py:60-79 does (COO of [[0,R],[R^T,0]] + gcn_norm), then ``torch.sparse.mm`` per layer in the CSR layout
    (what torch_sparse.matmul runs on the CPU, layers.py:19-20) over the FULL graph, L layers per step, on all host
    threads that help (thread count swept, best kept).  The COO and dense-edge variants of the reference
    (lightgcl.py:130 style / layers.py:13-17) are timed once on bounded row slices and reported beside it.  If the full
    graph cannot fit the time budget on this host, the largest row slice that does is used and named.  No CUDA here."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O

    U, I, E, D, L = WORKLOADS[args.workload]
    N, nnz = U + I, 2 * E
    cores = os.cpu_count() or 1
    budget_s = float(os.environ.get("B200GCN_REF_BUDGET_S", "150"))
    t_all = time.perf_counter()
    torch.set_num_threads(cores)
    uid, iid = O.synth_interactions(U, I, E, seed=0)
    t0 = time.perf_counter()
    ei, ew = O.build_norm_adj(uid, iid, U, I)                 # dataset.py:60-79 (COO + gcn_norm)
    adj_build_s = time.perf_counter() - t0
    del uid, iid
    t0 = time.perf_counter()
    key, perm = torch.sort(ei[1] * N + ei[0])                   # SparseTensor(...).t(): CSR keyed by destination
    crow = torch.zeros(N + 1, dtype=torch.int64)
    crow[1:] = torch.cumsum(torch.bincount(key // N, minlength=N), 0)
    col, val = ei[0][perm].contiguous(), ew[perm].contiguous()
    del key, perm
    csr_build_s = time.perf_counter() - t0
    x = torch.cat([O.xavier_uniform_table(U, D, 1), O.xavier_uniform_table(I, D, 2)])

    def csr_rows(n_rows):
        e = int(crow[n_rows])
        return torch.sparse_csr_tensor(crow[: n_rows + 1], col[:e], val[:e], size=(n_rows, N)), e

    # ---- thread sweep on a 5 % row slice; oversubscribed hosts are faster with fewer threads
    a_s, e_s = csr_rows(max(1000, N // 20))
    best_t, best_thr = None, cores
    sweep = {}
    for thr in sorted({cores, max(1, cores // 2), max(1, cores // 4)}, reverse=True):
        torch.set_num_threads(thr)
        O.propagate_sparse(a_s, x)
        t0 = time.perf_counter()
        for _ in range(2):
            O.propagate_sparse(a_s, x)
        dt = (time.perf_counter() - t0) / 2
        sweep[thr] = round(e_s / dt / 1e6, 1)
        if best_t is None or dt < best_t:
            best_t, best_thr = dt, thr
    torch.set_num_threads(best_thr)
    # ---- the reference's other two forms, once, on bounded slices (reported, not the headline)
    n_coo = max(1000, N // 20)
    e_coo = int(crow[n_coo])
    rows_coo = torch.repeat_interleave(torch.arange(n_coo), crow[1:n_coo + 1] - crow[:n_coo])
    t0 = time.perf_counter()
    a_coo = torch.sparse_coo_tensor(torch.stack([rows_coo, col[:e_coo]]), val[:e_coo], (n_coo, N)).coalesce()
    coalesce_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.propagate_sparse(a_coo, x)
    coo_eps = e_coo / (time.perf_counter() - t0)
    n_de = max(200, N // 100)
    e_de = int(crow[n_de])
    rows_de = torch.repeat_interleave(torch.arange(n_de), crow[1:n_de + 1] - crow[:n_de])
    t0 = time.perf_counter()
    O.propagate_scatter(x, torch.stack([col[:e_de], rows_de]), val[:e_de], n_dst=n_de)
    dense_eps = e_de / (time.perf_counter() - t0)
    del a_coo, rows_coo, rows_de

    # ---- headline: the full graph when it fits the budget, else the largest row slice that does
    est_layer = best_t * (nnz / e_s)
    spent = time.perf_counter() - t_all
    n_steps = args.warmup + args.steps
    frac = min(1.0, max(0.0, (budget_s - min(spent, budget_s * 0.5)) / (est_layer * L * n_steps)))
    n_rows = N if frac >= 1.0 else max(1000, int(N * frac))
    a, e_cnt = csr_rows(n_rows)

    def step():
        y = x
        for _ in range(L):
            y = O.propagate_sparse(a, y if n_rows == N else x)
        return y

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    t = (time.perf_counter() - t0) / args.steps
    value = e_cnt * L / t
    full = n_rows == N
    sample = ((f"the full {args.workload} graph: {e_cnt} directed edges per layer, {L} chained layers per step, "
               if full else
               f"rows [0,{n_rows}) of the {args.workload} graph = {e_cnt} of {nnz} directed edges per layer, {L} layers per "
               f"step (each from the full {N}x{D} table; the full graph would not fit the {budget_s:.0f} s budget here), ")
              + f"torch.sparse.mm CSR, {best_thr} of {cores} threads")
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: LightGCN propagation U={U} I={I} E={E} (nnz={nnz}) D={D} L={L}",
                   "sample": sample, "same_config": full},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": best_thr, "kind": "port", "sample": sample,
                         "host_cores": cores, "thread_sweep_Medges_per_s": sweep,
                         "adjacency_build_s": round(adj_build_s, 2), "csr_build_s": round(csr_build_s, 2),
                         "variants": {"csr_full": value, "coo_5pct_rows": coo_eps, "dense_edge_1pct_rows": dense_eps,
                                      "coo_coalesce_s_5pct_rows": round(coalesce_s, 2)}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def main():
    sys.modules.setdefault("bench", sys.modules[__name__])   # `import bench` elsewhere must see THIS module's state
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-rows", type=int, default=100_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parity-rows", type=int, default=10_000,
                    help="N > 1: rows per rank checked against the CPU oracle inside the bench run")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    capture_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
