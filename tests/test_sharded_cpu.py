"""CPU: host logic of the row-sharded path — partition arithmetic, and the full K-layer exchange protocol on
2 (and 3) gloo ranks with the local product injected from the oracle (test-only), against the unsharded oracle."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from recbole_gnn_b200.sharded import ShardPlan, ShardedPropagator


def test_plan_is_a_balanced_bijection():
    for U, I, P in [(10, 7, 2), (101, 64, 3), (1000, 999, 8), (5, 5, 4)]:
        plan = ShardPlan(U, I, P)
        assert sum(plan.u_cnt) == U and sum(plan.i_cnt) == I
        assert max(plan.u_cnt) - min(plan.u_cnt) <= 1 and max(plan.i_cnt) - min(plan.i_cnt) <= 1
        nu = plan.relabel_users(torch.arange(U))
        ni = plan.relabel_items(torch.arange(I))
        allids = torch.cat([nu, ni])
        assert allids.unique().numel() == U + I and int(allids.max()) < plan.n_full
        for p in range(P):   # a rank's block holds its users first, then its items, contiguously
            blk = torch.cat([nu[plan.ub[p]:plan.ub[p + 1]], ni[plan.ib[p]:plan.ib[p + 1]]])
            assert torch.equal(blk, p * plan.n_pad + torch.arange(plan.n_loc[p]))


def test_local_edges_cover_every_directed_edge_once():
    U, I, E, P = 60, 45, 2000, 4
    uid, iid = O.synth_interactions(U, I, E, seed=1)
    w = torch.rand(E)
    plan = ShardPlan(U, I, P)
    tot = 0
    for p in range(P):
        d, s, wl = plan.local_edges(p, uid, iid, w)
        assert d.numel() == s.numel() == wl.numel()
        assert int(d.max()) < plan.n_loc[p] and int(s.max()) < plan.n_full
        tot += d.numel()
    assert tot == 2 * E


def _local_spmm(edges, x_full):
    d, s, w = edges
    n_loc = int(d.max()) + 1 if d.numel() else 0
    return None, d, s, w


def _worker(rank, world, port, U, I, E, D, L, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid, iid = O.synth_interactions(U, I, E, seed=3, zipf_alpha=1.1)
        _, w = O.build_bipartite_inter_mat(uid, iid, U, I, row_norm=False)    # = gcn_norm weight per interaction
        plan = ShardPlan(U, I, world)
        d, s, wl = plan.local_edges(rank, uid, iid, w)
        n_loc = plan.n_loc[rank]

        def local(edges, x_full):
            dd, ss, ww = edges
            return torch.zeros(n_loc, x_full.size(1)).index_add_(0, dd, ww[:, None] * x_full[ss])

        prop = ShardedPropagator(plan, rank, d, s, wl, D, "cpu", exchange="allgather", local_spmm=local)
        xu, xi = O.xavier_uniform_table(U, D, 5), O.xavier_uniform_table(I, D, 6)
        xu_l, xi_l = plan.scatter_tables(rank, xu, xi)
        out = prop.forward(xu_l, xi_l, L).clone()
        out2 = prop.forward(xu_l, xi_l, L)                 # second call re-uses the tables
        ei, ew = O.build_norm_adj(uid, iid, U, I)
        u_ref, i_ref = O.lightgcn_forward(xu, xi, ei, ew, L)
        ref = torch.cat([u_ref[plan.ub[rank]:plan.ub[rank + 1]], i_ref[plan.ib[rank]:plan.ib[rank + 1]]])
        err = (out - ref).abs().max().item() / ref.abs().max().item()
        q.put((rank, err, torch.equal(out, out2)))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_protocol_matches_unsharded_oracle(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 53, 41, 1500, 8, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, same in res:
        assert err < 1e-5, (rank, err)
        assert same


def test_sharded_refuses_cpu_without_injection():
    plan = ShardPlan(4, 4, 1)
    e = torch.zeros(0, dtype=torch.int64)
    with pytest.raises(RuntimeError):
        ShardedPropagator(plan, 0, e, e, torch.zeros(0), 8, "cpu", exchange="allgather")


@pytest.mark.parametrize("L", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("P", [1, 2, 8])
def test_chain_schedule_dependencies_are_sufficient(L, P):
    """Replay the chain kernel's phase list under adversarial timing: a rank may START phase p as soon as it has
    started p-1 (tiles are handed out in order, CTAs of one rank overlap adjacent phases) and the flags named by
    ``wait`` (complete on every rank) / ``wait_local`` (complete on this rank) are in.  Everything a phase reads —
    the gathered half-table of the previous layer from ALL ranks, and its own rows of the running layer sum — must
    be complete at that moment."""
    import random
    from recbole_gnn_b200.sharded import ShardPlan
    sched = ShardPlan(100, 80, P).chain_schedule(L)
    assert [s["kind"] for s in sched[:2]] == ["publish", "publish"] and len(sched) == 2 + 2 * L
    other = {"U": "I", "I": "U"}
    for trial in range(200):
        rnd = random.Random(trial)
        started = [[None] * len(sched) for _ in range(P)]
        done = [[None] * len(sched) for _ in range(P)]
        t_rank = [rnd.random() for _ in range(P)]            # ranks launch at different times
        for p, ph in enumerate(sched):
            for r in rnd.sample(range(P), P):
                t = t_rank[r] if p == 0 else started[r][p - 1]
                if ph["wait"] >= 0:
                    t = max(t, max(done[q][ph["wait"]] for q in range(P)))
                if ph["wait_local"] >= 0:
                    t = max(t, done[r][ph["wait_local"]])
                started[r][p] = t
                # reads
                if ph["kind"] == "spmm":
                    src = next(k for k, s in enumerate(sched) if s["layer"] == ph["layer"] - 1 and s["half"] == other[ph["half"]])
                    assert all(done[q][src] <= t for q in range(P)), (trial, p, "gather table incomplete")
                    if ph["layer"] > 1:
                        acc_src = next(k for k, s in enumerate(sched) if s["layer"] == ph["layer"] - 1 and s["half"] == ph["half"])
                        assert done[r][acc_src] <= t, (trial, p, "running sum rows incomplete")
                # a phase ends after all earlier phases of the rank ended (CTAs leave phases in order)
                dur = rnd.random() * (3.0 if rnd.random() < 0.2 else 0.3)
                done[r][p] = max([t + dur] + [done[r][k] for k in range(p)])
    # no phase waits on its direct predecessor across GPUs (that would expose the NVLink flight time)
    assert all(s["wait"] <= max(p - 2, -1) for p, s in enumerate(sched))


def test_merged_phase_tile_mapping_is_a_bijection():
    """Mirror of the chain kernel's interleaving rule for a merged phase pair (csrc/spmm.cu, `merge_next`): positions
    0, period, 2 period, .. of the shared tile range belong to the first phase (period = total // n_first), every
    other position to the second; each tile of either phase must appear exactly once for any pair of tile counts."""
    import random

    def mapping(n_first, n_second):
        total, per = n_first + n_second, (n_first + n_second) // n_first
        first, second = [], []
        for j in range(total):
            if j % per == 0 and j // per < n_first:
                first.append(j // per)
            else:
                second.append(j - min((j + per - 1) // per, n_first))
        return first, second

    rnd = random.Random(0)
    cases = [(489, 3907), (1, 1), (5, 1), (600, 3), (1, 4000)] + [(rnd.randint(1, 700), rnd.randint(1, 6000)) for _ in range(300)]
    for nf, ns in cases:
        first, second = mapping(nf, ns)
        assert sorted(first) == list(range(nf)) and sorted(second) == list(range(ns)), (nf, ns)
        assert first == sorted(first) and second == sorted(second)          # both phases are handed out in order
    # the publish tiles are spread over the whole range, not bunched in front (that is the point of merging)
    first, _ = mapping(489, 3907)
    total, per = 489 + 3907, (489 + 3907) // 489
    assert (len(first) - 1) * per > 0.85 * total
