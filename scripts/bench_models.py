"""Timings of the other two BASELINE.json single-GPU configs (parity-test cases, not bench lines):
configs[2] NGCF 3 layers (W-transform + LeakyReLU + L2-normalise) and configs[3] SimGCL 3 layers x 3 views per
step, on the config-2 graph.  Writes one JSON object per config to stdout (kept under profiles/)."""
import json, sys, torch
sys.path.insert(0, '.')
import recbole_gnn_b200 as rg
from recbole_gnn_b200 import functional as F_
from bench import synth_graph_device, xavier_tables_device, algorithmic_bytes_per_layer, measured_peak_gbs

dev = torch.device('cuda:0')
U = I = 1_000_000; E = 100_000_000; D = 64; L = 3
N, nnz = U + I, 2 * E
uid, iid = synth_graph_device(U, I, E, dev)
h = rg.GraphHandle.from_interactions(uid, iid, U, I).gcn_norm().to(dev)
del uid, iid
xu, xi = xavier_tables_device(U, I, D, dev)
peak, _ = measured_peak_gbs()

def timeit(fn, warm=3, reps=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps

with torch.no_grad():
    # configs[1] for reference
    ms = timeit(lambda: F_.lightgcn_propagate(h, xu, xi, L))
    print(json.dumps({"config": "LightGCN cfg2 L=3", "ms_per_step": ms, "edges_per_s": nnz * L / ms * 1e3,
                      "algo_GBps": algorithmic_bytes_per_layer(nnz, N, D) * L / ms / 1e6}))
    # configs[2]: NGCF
    g = torch.Generator(device=dev).manual_seed(5)
    std = (2.0 / (D + D)) ** 0.5
    W = [(torch.randn(D, D, generator=g, device=dev) * std, torch.zeros(D, device=dev),
          torch.randn(D, D, generator=g, device=dev) * std, torch.zeros(D, device=dev)) for _ in range(L)]
    ms = timeit(lambda: F_.ngcf_forward(h, xu, xi, W))
    p = torch.empty(N, D, device=dev); x0 = torch.cat([xu, xi]); o = torch.empty(N, D, device=dev)
    ms_tail = timeit(lambda: F_.bignn_tail(p, x0, *W[0], out=o))
    ms_spmm = timeit(lambda: F_.spmm_raw(h, x0, y=p))
    wide = torch.zeros(N, 4 * D, device=dev); wide[:, :D] = x0
    ms_spmm_strided = timeit(lambda: F_.spmm_raw(h, wide[:, :D], y=p))
    del wide
    print(json.dumps({"config": "NGCF cfg3 L=3 (hidden 64,64,64; dropout 0)", "ms_per_step": ms,
                      "edges_per_s": nnz * L / ms * 1e3, "ms_spmm_layer": ms_spmm, "ms_spmm_layer_from_1KB_strided_slice": ms_spmm_strided,
                      "ms_tail_layer": ms_tail,
                      "tail_GFLOPs": 4 * N * D * D / ms_tail / 1e6,
                      "algo_GBps": (algorithmic_bytes_per_layer(nnz, N, D) + N * D * 4) * L / ms / 1e6}))
    # configs[3]: SimGCL, 1 clean + 2 perturbed forwards per step, fused Philox noise
    def simgcl_step():
        F_.simgcl_propagate(h, xu, xi, L, 0.1, perturbed=False)
        F_.simgcl_propagate(h, xu, xi, L, 0.1, perturbed=True, seed=1)
        F_.simgcl_propagate(h, xu, xi, L, 0.1, perturbed=True, seed=2)
    ms = timeit(simgcl_step, reps=5)
    ms7 = timeit(lambda: F_.simgcl_views(h, xu, xi, L, 0.1, seeds=(1, 2)), reps=5)
    print(json.dumps({"config": "SimGCL cfg4 L=3 x 3 views, shared first layer (7 SpMM + 2 identity passes)",
                      "ms_per_step": ms7, "edges_per_s_9spmm_equiv": nnz * L * 3 / ms7 * 1e3}))
    print(json.dumps({"config": "SimGCL cfg4 L=3 x 3 views (9 SpMM, in-kernel Philox noise)", "ms_per_step": ms,
                      "edges_per_s": nnz * L * 3 / ms * 1e3,
                      "algo_GBps": algorithmic_bytes_per_layer(nnz, N, D) * L * 3 / ms / 1e6}))

# ---- training steps at config-2 size: fused LightGCN step; NGCF forward + backward with the tcgen05 tails
from recbole_gnn_b200 import train as TR

ds = rg.InteractionDataset(torch.zeros(1, dtype=torch.int64), torch.zeros(1, dtype=torch.int64), U, I, device=dev)


class _Shim:
    """model-shaped holder of the tables / graph the fused step reads"""
    USER_ID, ITEM_ID, NEG_ITEM_ID = "user_id", "item_id", "neg_item_id"
    n_layers, reg_weight, require_pow = L, 1e-4, False

    def __init__(self):
        self.user_embedding = torch.nn.Embedding(U, D, device=dev)
        self.item_embedding = torch.nn.Embedding(I, D, device=dev)

    def _graph(self):
        return h

    def _clear_restore(self):
        pass


m = _Shim()
step = TR.LightGCNTrainStep(m, lr=1e-3)
gb = torch.Generator(device=dev).manual_seed(3)
B = 4096
inter = {"user_id": torch.randint(1, U, (B,), generator=gb, device=dev),
         "item_id": torch.randint(1, I, (B,), generator=gb, device=dev),
         "neg_item_id": torch.randint(1, I, (B,), generator=gb, device=dev)}
ms_fused = timeit(lambda: step.step(inter), warm=2, reps=5)
# the same step through autograd + torch.optim.Adam (engine kernels for the propagation, torch for the rest)
pu, pi = m.user_embedding.weight, m.item_embedding.weight
opt = torch.optim.Adam([pu, pi], lr=1e-3)


def autograd_step():
    opt.zero_grad(set_to_none=True)
    ua, ia = F_.lightgcn_propagate(h, pu, pi, L)
    u, pos, neg = ua[inter["user_id"]], ia[inter["item_id"]], ia[inter["neg_item_id"]]
    loss = -torch.log(1e-10 + torch.sigmoid((u * pos).sum(1) - (u * neg).sum(1))).mean()
    loss = loss + 1e-4 * (pu[inter["user_id"]].norm() + pi[inter["item_id"]].norm() + pi[inter["neg_item_id"]].norm()) / B
    loss.backward()
    opt.step()


ms_auto = timeit(autograd_step, warm=2, reps=5)
print(json.dumps({"config": "LightGCN cfg2 TRAINING step (forward + BPR/EmbLoss + backward + Adam), batch 4096",
                  "ms_fused_step": ms_fused, "ms_autograd_plus_torch_adam": ms_auto,
                  "edges_per_s_fwd_plus_bwd": 2 * nnz * L / ms_fused * 1e3}))
del step, opt
Wg = [tuple(t.clone().requires_grad_(True) for t in w) for w in W]
xg = [xu.clone().requires_grad_(True), xi.clone().requires_grad_(True)]
keeps = [torch.rand(N, D, device=dev) >= 0.1 for _ in range(L)]


def ngcf_train():
    x = torch.cat(xg)
    outs = [x]
    for l, (w1, b1, w2, b2) in enumerate(Wg):
        x = F_.bignn_tail_autograd(F_.spmm(h, x), x, w1, b1, w2, b2, slope=0.2, keep=keeps[l], drop_p=0.1, normalize=True)
        outs.append(x)
    out = torch.cat(outs, 1)
    out.square().sum().backward()


ms_ngcf = timeit(ngcf_train, warm=2, reps=3)
print(json.dumps({"config": "NGCF cfg3 forward + backward (3 layers, message dropout 0.1; tcgen05 tails both ways)",
                  "ms_fwd_bwd": ms_ngcf}))
