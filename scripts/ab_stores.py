"""A/B of store cache policy on the 3-layer cfg2 forward: per-layer times."""
import sys, torch
sys.path.insert(0, '.')
import recbole_gnn_b200 as rg
from recbole_gnn_b200 import functional as F_
from bench import synth_graph_device, xavier_tables_device
dev = torch.device('cuda:0')
U = I = 1_000_000; E = 100_000_000; D = 64; L = 3
uid, iid = synth_graph_device(U, I, E, dev)
h = rg.GraphHandle.from_interactions(uid, iid, U, I).gcn_norm().to(dev)
del uid, iid
xu, xi = xavier_tables_device(U, I, D, dev)
for rep in range(2):
    for name, flags in (("default", 0), ("cs-stores", 1 << 25)):
        F_.DEFAULT_FLAGS = flags
        with torch.no_grad():
            for _ in range(3): F_.lightgcn_propagate(h, xu, xi, L)
            torch.cuda.synchronize()
            t = F_.LaunchTimer()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with t:
                s.record()
                for _ in range(10): F_.lightgcn_propagate(h, xu, xi, L)
                e.record()
            torch.cuda.synchronize()
            d = t.durations_ms()
            print(name, "step ms", round(s.elapsed_time(e) / 10, 3), "by layer", [round(sum(d[i::3]) / 10, 3) for i in range(3)], flush=True)
# what a layer costs without any [N, D] write: y=None and acc_out to a tiny dummy is not possible; compare y-only vs acc-only
x = torch.cat([xu, xi]); y = torch.empty_like(x); acc = torch.empty_like(x)
def tm(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return round(s.elapsed_time(e) / n, 3)
for name, flags in (("default", 0), ("cs-stores", 1 << 25)):
    F_.DEFAULT_FLAGS = flags
    print(name, "y only", tm(lambda: F_.spmm_raw(h, x, y=y)), "| y+acc", tm(lambda: F_.spmm_raw(h, x, y=y, acc_in=x, acc_out=acc)),
          "| acc only (in-place)", tm(lambda: F_.spmm_raw(h, x, acc_in=acc, acc_out=acc)), flush=True)
