"""Full-sort evaluation at BASELINE config-2 size: 4096 users x 1 M items, D = 64, top-50 with a seen-item mask.
Engine (tcgen05 contraction + fused top-k, no score matrix) vs the reference's route restated in torch on the same GPU
(torch.matmul -> mask -> torch.topk, lightgcn.py:123-133 + RecBole's evaluator).  Run under gpurun."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from recbole_gnn_b200 import functional as F_

B, I, D, k, H = 4096, 1_000_000, int(sys.argv[1]) if len(sys.argv) > 1 else 64, 50, 100
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
u = torch.randn(B, D, generator=g, device=dev) * 0.1
items = torch.randn(I, D, generator=g, device=dev) * 0.1
rows = torch.arange(B, device=dev).repeat_interleave(H)
its = torch.randint(1, I, (B * H,), generator=g, device=dev)


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        out = fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n, out


def ref():
    s = torch.matmul(u, items.t())
    s[rows, its] = float("-inf")
    s[:, 0] = float("-inf")
    return torch.topk(s, k, dim=1)


ms_ours, (sc, ids) = timed(lambda: F_.full_sort_topk(u, items, k, history=(rows, its)))
ms_dense, _ = timed(lambda: F_.full_sort_scores(u, items), n=3)
ms_ref, (v_ref, i_ref) = timed(ref, n=3)
agree = (ids == i_ref).float().mean().item()
err = ((sc - v_ref).abs().max() / v_ref.abs().max()).item()
flops = 2.0 * B * I * D
print(json.dumps({"B": B, "I": I, "D": D, "k": k, "history_per_user": H,
                  "fullsort_topk_ms": round(ms_ours, 3), "fullsort_dense_scores_ms": round(ms_dense, 3),
                  "torch_matmul_mask_topk_ms": round(ms_ref, 3), "speedup": round(ms_ref / ms_ours, 2),
                  "fp32_equivalent_tflops": round(flops / ms_ours / 1e9, 1),
                  "tf32_tensor_tflops_issued": round(4 * flops / ms_ours / 1e9, 1),
                  "ids_equal_to_torch": agree, "score_scaled_err": err}))
