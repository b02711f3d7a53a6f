"""One launch of the NGCF tail at config-3 size (for ncu):  python scripts/tail_once.py [n_rows] [keep:0|1]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from recbole_gnn_b200 import functional as F_
n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000, 64
use_keep = len(sys.argv) > 2 and sys.argv[2] == "1"
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
p, x = torch.randn(n, d, generator=g, device=dev) * 0.1, torch.randn(n, d, generator=g, device=dev)
w1, w2 = torch.randn(d, d, generator=g, device=dev) * 0.125, torch.randn(d, d, generator=g, device=dev) * 0.125
b1, b2 = torch.randn(d, generator=g, device=dev) * 0.1, torch.randn(d, generator=g, device=dev) * 0.1
keep = (torch.rand(n, d, generator=g, device=dev) > 0.1).to(torch.uint8) if use_keep else None
cat = torch.empty(n, 4 * d, device=dev)
out2 = torch.empty(n, d, device=dev)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(4):
    s.record()
    F_.bignn_tail(p, x, w1, b1, w2, b2, keep=keep, drop_p=0.1 if use_keep else 0.0, out=cat[:, d:2 * d], out2=out2)
    e.record()
    torch.cuda.synchronize()
    print("launch", it, "ms", round(s.elapsed_time(e), 4))

# backward (training): fused tcgen05 dgrad + library wgrad GEMM
t = torch.empty(n, d, device=dev)
F_.bignn_tail(p, x, w1, b1, w2, b2, keep=keep, drop_p=0.1 if use_keep else 0.0, out=out2, pre_out=t)
g = torch.randn(n, d, generator=g, device=dev) if False else torch.ones(n, d, device=dev)
for it in range(3):
    s.record()
    F_.bignn_tail_backward_fused(p, x, w1, w2, t, keep, 0.1 if use_keep else 0.0, 0.2, True, g)
    e.record()
    torch.cuda.synchronize()
    print("backward (fused dgrad kernel + wgrad GEMM + bias sum)", it, "ms", round(s.elapsed_time(e), 4))
ks = keep.float() / 0.9 if use_keep else None
for it in range(2):
    s.record()
    F_.bignn_tail_backward(p, x, w1, w2, t, out2, ks, 0.2, True, g)
    e.record()
    torch.cuda.synchronize()
    print("backward (round-1 torch algebra + cuBLAS)", it, "ms", round(s.elapsed_time(e), 4))
