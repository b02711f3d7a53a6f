"""Small-shape run of the round-2 kernels for compute-sanitizer: tcgen05 NGCF tail forward / backward (several tiles
per CTA), full-sort top-k / dense scores, BPR loss + Adam, the training step.  `python scripts/sanitize_r2.py [quick]`"""
import sys, torch
sys.path.insert(0, '.')
import recbole_gnn_b200 as rg
from recbole_gnn_b200 import functional as F_, train as TR
from oracle import oracle as O
dev = "cuda:0"
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
g = torch.Generator(device=dev).manual_seed(0)
d = 64
n = 128 * 148 * 3 + 37 if not quick else 128 * 148 * 2 + 5          # >= 3 tiles on some CTAs
p, x = torch.randn(n, d, generator=g, device=dev), torch.randn(n, d, generator=g, device=dev)
w1, w2 = torch.randn(d, d, generator=g, device=dev) * 0.1, torch.randn(d, d, generator=g, device=dev) * 0.1
b1, b2 = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
keep = (torch.rand(n, d, generator=g, device=dev) > 0.1).to(torch.uint8)
cat = torch.empty(n, 2 * d, device=dev); out2 = torch.empty(n, d, device=dev); t = torch.empty(n, d, device=dev)
F_.bignn_tail(p, x, w1, b1, w2, b2, keep=keep, drop_p=0.1, out=cat[:, d:], out2=out2, pre_out=t)
F_.bignn_tail_backward_fused(p, x, w1, w2, t, keep, 0.1, 0.2, True, torch.ones(n, d, device=dev))
B, I = (300, 20011) if not quick else (130, 3000)
for D in (64, 128):
    u, items = torch.randn(B, D, generator=g, device=dev), torch.randn(I, D, generator=g, device=dev)
    rows = torch.arange(B, device=dev).repeat_interleave(5)
    its = torch.randint(1, I, (B * 5,), generator=g, device=dev)
    F_.full_sort_topk(u, items, 20, history=(rows, its))
    F_.full_sort_scores(u, items)
U, I2, E = 600, 400, 12000
uid, iid = O.synth_interactions(U, I2, E, seed=1)
ds = rg.InteractionDataset(uid, iid, U, I2, device=dev)
m = rg.LightGCN({"device": dev, "enable_sparse": True, "embedding_size": 64, "n_layers": 2}, ds).to(dev)
step = TR.LightGCNTrainStep(m, lr=1e-2)
inter = {"user_id": uid[:777].to(dev), "item_id": iid[:777].to(dev), "neg_item_id": torch.randint(1, I2, (777,), device=dev)}
step.step(inter); step.step(inter)
torch.cuda.synchronize()
print("sanitize r2 run ok")
