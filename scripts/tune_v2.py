"""Sweep of the v2 kernel (gathers in flight, rows per warp) against v1 on cfg2."""
import sys, torch
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
    return s.elapsed_time(e) / reps, (y - y0).abs().max().item()
for rep in range(2):
    print("v1", run(1))
    for u, code in ((8, 8), (4, 4)):
        for rpw in (1, 4, 8):
            for pf in (8, 16, 32):
                ms, err = run(3 | (code << 4) | (rpw << 16) | ((pf // 8) << 8))
                print(f"A D={D} U={u} rpw={rpw} pf={pf}: {ms:.3f} ms {algo/ms/1e6:.0f} GB/s algo err={err:.1e}", flush=True)
    for u, code in ((8, 8), (4, 4), (16, 1)):
        for rpw in (2, 4, 8):
            for pf in (0, 8, 16, 32):
                ms, err = run(2 | (code << 4) | (rpw << 16) | ((pf // 8 if pf else 255) << 8))
                print(f"B D={D} U={u} rpw={rpw} pf={pf}: {ms:.3f} ms {algo/ms/1e6:.0f} GB/s algo err={err:.1e}", flush=True)
