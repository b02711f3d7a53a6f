"""Skewed graph at config-2 size: items drawn from a Zipf(1.1)-like law (hub rows with millions of entries).
Times one layer and the 3-layer forward, and re-checks sampled rows (incl. the largest hubs) in float64."""
import json, sys, torch
sys.path.insert(0, '.')
import recbole_gnn_b200 as rg
from recbole_gnn_b200 import functional as F_
dev = torch.device('cuda:0')
U = I = 1_000_000; E = 100_000_000; D = 64
gen = torch.Generator(device=dev).manual_seed(0)
uid = torch.randint(1, U, (E,), generator=gen, device=dev)
# inverse-CDF sampling of p(k) ~ k^-1.1 on k = 1..I-1, then a random permutation of item ids
ranks = torch.arange(1, I, device=dev, dtype=torch.float64)
cdf = torch.cumsum(ranks.pow(-1.1), 0); cdf /= cdf[-1].clone()
iid = torch.searchsorted(cdf, torch.rand(E, generator=gen, device=dev, dtype=torch.float64)).clamp_(max=I - 2)
perm = torch.randperm(I - 1, generator=gen, device=dev) + 1
iid = perm[iid]
h = rg.GraphHandle.from_interactions(uid, iid, U, I).gcn_norm().to(dev)
rowptr, col, val = h.csr()
deg = rowptr[1:] - rowptr[:-1]
x = torch.rand(U + I, D, device=dev, generator=gen) * 2 - 1
y = torch.empty_like(x)
def tm(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
ms = tm(lambda: F_.spmm_raw(h, x, y=y))
xu, xi = x[:U].contiguous(), x[U:].contiguous()
ms3 = tm(lambda: F_.lightgcn_propagate(h, xu, xi, 3))
rows = torch.cat([deg.topk(8).indices.cpu(), torch.randint(0, U + I, (200,))])
err = 0.0
for r in rows.tolist():
    b, e = int(rowptr[r]), int(rowptr[r + 1])
    ref = (val[b:e].double()[:, None] * x[col[b:e].long()].double()).sum(0)
    err = max(err, float((y[r].double() - ref).abs().max() / max(1e-30, float(ref.abs().max()))))
print(json.dumps({"config": "Zipf(1.1) items, U=I=1M, E=100M, D=64", "hubs": h._n_hubs, "chunks": h._n_chunks,
                  "max_degree": int(deg.max()), "ms_layer": ms, "ms_3layer": ms3, "edges_per_s_3layer": 6e8 / ms3 * 1e3,
                  "max_rel_err_sampled_rows_vs_f64": err}))
