"""Small-shape run of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, torch
sys.path.insert(0, '.')
import recbole_gnn_b200 as rg
from recbole_gnn_b200 import functional as F_
from oracle import oracle as O
dev = "cuda:0"
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"      # racecheck is 10-100x slower: one skewed config
for D in ((64,) if quick else (8, 64, 128, 200)):
    for alpha in ((1.3,) if quick else (None, 1.3)):
        U, I, E = (600, 100, 12000) if quick else (3000, 400, 60000)
        uid, iid = O.synth_interactions(U, I, E, seed=1, zipf_alpha=alpha)
        ds = rg.InteractionDataset(uid, iid, U, I, device=dev)
        h, _ = ds.get_norm_adj_mat(enable_sparse=True)
        h = h.to(dev)
        xu, xi = torch.randn(U, D, device=dev), torch.randn(I, D, device=dev)
        u, i = F_.lightgcn_propagate(h, xu, xi, 3)
        F_.simgcl_propagate(h, xu, xi, 2, 0.1, perturbed=True, seed=3)
        F_.simgcl_views(h, xu, xi, 2, 0.1, seeds=(1, 2))
        x = torch.cat([xu, xi]).requires_grad_(True)
        y = rg.LightGCNConv(D)(x, h, None); y.sum().backward()
        ei, ew = ds.get_norm_adj_mat(enable_sparse=False)
        rg.LightGCNConv(D)(x.detach(), ei, ew)
        h.t(); h.coo(); h.masked(torch.rand(h.nnz(), device=dev) > 0.5)
        bi, bw = ds.get_bipartite_inter_mat("user", False)
        rg.BipartiteGCNConv(D)((xi, xu), bi.flip([0]), bw, size=(I, U))
        if D <= 128:
            W = [(torch.randn(D, D, device=dev), torch.zeros(D, device=dev), torch.randn(D, D, device=dev), torch.zeros(D, device=dev)) for _ in range(2)]
            F_.ngcf_forward(h, xu, xi, W)
torch.cuda.synchronize()
print("sanitize run ok")
