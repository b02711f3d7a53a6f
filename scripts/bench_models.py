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
