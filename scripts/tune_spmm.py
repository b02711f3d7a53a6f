"""A/B the SpMM kernel variants on the cfg2 graph (one GPU call): flags word -> ms per layer."""
import sys, itertools, torch
sys.path.insert(0, '.')
import recbole_gnn_b200 as rg
from recbole_gnn_b200 import functional as F_
dev = 'cuda:0'
U = I = 1_000_000; E = 100_000_000; D = int(sys.argv[1]) if len(sys.argv) > 1 else 64
gen = torch.Generator(device=dev).manual_seed(0)
uid = torch.randint(1, U, (E,), generator=gen, device=dev); iid = torch.randint(1, I, (E,), generator=gen, device=dev)
h = rg.GraphHandle.from_interactions(uid, iid, U, I).gcn_norm().to(dev)
del uid, iid
x = (torch.rand(U + I, D, device=dev) * 2 - 1)
y0 = torch.empty_like(x); y = torch.empty_like(x)
F_.DEFAULT_FLAGS = 1
F_.spmm_raw(h, x, y=y0)
algo = (200e6 * (4 * D + 8) + 2e6 * (4 * D + 4))
def run(flags, reps=10):
    F_.DEFAULT_FLAGS = flags
    for _ in range(3): F_.spmm_raw(h, x, y=y)
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): F_.spmm_raw(h, x, y=y)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    err = (y - y0).abs().max().item()
    return ms, err
print("v1", run(1))
res = []
for u, pf, rpw, bulk in itertools.product((4, 8), (0, 16, 32, 64, 128, 256), (4, 8, 16, 31), (0, 1)):
    flags = 2 | (u << 4) | ((pf // 8) << 8) | (rpw << 16) | (bulk << 24)
    if pf == 0: flags = 2 | (u << 4) | (1 << 8) | (rpw << 16) | (bulk << 24)   # pf=8: effectively none
    ms, err = run(flags, reps=5)
    res.append((ms, u, pf, rpw, bulk, err))
    print(f"U={u} pf={pf} rpw={rpw} bulk={bulk}: {ms:.3f} ms  {algo/ms/1e6:.0f} GB/s algo  err={err:.2e}", flush=True)
res.sort()
print("best", res[:5])
