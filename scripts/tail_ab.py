"""A/B of the NGCF layer tail: tcgen05 3xTF32 kernel vs the fp32 CUDA-core kernel (B200GCN_TAIL=cuda), accuracy against
float64 and time per launch at config-3 size.  Run under gpurun:  python scripts/tail_ab.py [n_rows]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run():
    import torch
    from recbole_gnn_b200 import functional as F_
    n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000, 64
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    p, x = torch.randn(n, d, generator=g, device=dev) * 0.1, torch.randn(n, d, generator=g, device=dev)
    w1, w2 = torch.randn(d, d, generator=g, device=dev) * 0.125, torch.randn(d, d, generator=g, device=dev) * 0.125
    b1, b2 = torch.randn(d, generator=g, device=dev) * 0.1, torch.randn(d, generator=g, device=dev) * 0.1
    keep = torch.rand(n, d, generator=g, device=dev) > 0.1
    cat = torch.empty(n, 4 * d, device=dev)
    out2 = torch.empty(n, d, device=dev)
    m = min(n, 200_000)
    t = (torch.nn.functional.linear((p[:m] + x[:m]).double(), w1.double(), b1.double()) +
         torch.nn.functional.linear((p[:m] * x[:m]).double(), w2.double(), b2.double()))
    ref = torch.nn.functional.normalize(torch.nn.functional.leaky_relu(t, 0.2) * keep[:m] / 0.9, p=2, dim=1)
    F_.bignn_tail(p, x, w1, b1, w2, b2, keep=keep, drop_p=0.1, out=cat[:, d:2 * d], out2=out2)
    torch.cuda.synchronize()
    err = (out2[:m].double() - ref).abs().max().item() / ref.abs().max().item()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        F_.bignn_tail(p, x, w1, b1, w2, b2, keep=keep, drop_p=0.1, out=cat[:, d:2 * d], out2=out2)
    s.record()
    for _ in range(20):
        F_.bignn_tail(p, x, w1, b1, w2, b2, keep=keep, drop_p=0.1, out=cat[:, d:2 * d], out2=out2)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    bytes_ = n * d * 4 * 4 + n * d                  # p, x in; out, out2 back; keep mask
    print({"variant": os.environ.get("B200GCN_TAIL", "tcgen05"), "n": n, "ms": round(ms, 4), "scaled_err_vs_f64": err,
           "GB/s": round(bytes_ / ms / 1e6, 1), "hbm_floor_ms_at_6552GBs": round(bytes_ / 6552.6e6, 4)})


if __name__ == "__main__":
    if os.environ.get("_TAIL_AB_CHILD"):
        run()
    else:
        for v in ("tcgen05", "cuda"):
            env = dict(os.environ, _TAIL_AB_CHILD="1", B200GCN_TAIL=v)
            subprocess.run([sys.executable, __file__] + sys.argv[1:], env=env, check=False)
