import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recbole_gnn_b200 import functional as F_
dev = "cuda:0"
for n, normalize, drop in ((40001, True, 0.1), (40001, False, 0.0)):
    g0 = torch.Generator().manual_seed(1)
    d = 64
    p, x = torch.randn(n, d, generator=g0) * 0.5, torch.randn(n, d, generator=g0)
    w1, w2 = torch.randn(d, d, generator=g0) * 0.1, torch.randn(d, d, generator=g0) * 0.1
    b1, b2 = torch.randn(d, generator=g0) * 0.1, torch.randn(d, generator=g0) * 0.1
    keep = (torch.rand(n, d, generator=g0) >= drop) if drop > 0 else None
    g = torch.randn(n, d, generator=g0)
    t = torch.nn.functional.linear(p + x, w1, b1) + torch.nn.functional.linear(p * x, w2, b2)
    z = torch.nn.functional.leaky_relu(t, 0.2)
    if keep is not None:
        z = z * keep / (1 - drop)
    out = torch.nn.functional.normalize(z, p=2, dim=1) if normalize else z
    ks = keep.float() / (1 - drop) if keep is not None else None
    ref = F_.bignn_tail_backward(p.double(), x.double(), w1.double(), w2.double(), t.double(), out.double(),
                                 None if ks is None else ks.double(), 0.2, normalize, g.double())
    kd = None if keep is None else keep.to(torch.uint8).to(dev)
    got = F_.bignn_tail_backward_fused(p.to(dev), x.to(dev), w1.to(dev), w2.to(dev), t.to(dev), kd, drop, 0.2, normalize, g.to(dev))
    for name, a, b in list(zip(("g_p", "g_x", "g_w1", "g_b", "g_w2"), got, ref))[:2]:
        e = (a.cpu().double() - b).abs()
        line = f"n={n} norm={normalize} drop={drop} {name}: max {e.max().item():.3e} scaled {e.max().item() / b.abs().max().item():.3e}"
        if e.dim() == 2 and e.size(0) == n:
            bad = torch.nonzero(e.max(1).values > 1e-4 * b.abs().max()).flatten()
            line += f" bad rows {bad.numel()}"
            if bad.numel():
                tiles = torch.unique(bad // 128)
                line += f" tiles {tiles[:12].tolist()} (n_tiles {(n + 127) // 128}) rows%128 {torch.unique(bad % 128)[:16].tolist()}"
        print(line)
