"""Shared test helpers: golden decoding and the error report the parity contract asks for."""
import numpy as np
import torch

ABS_TOL = 1e-4          # BASELINE.json north_star: per-element |delta| < 1e-4 (fp32)
REL_TOL = 1e-5          # self-imposed: max|delta| / max|ref| (fp32 re-association noise is ~1e-7*sqrt(deg))


def T(a):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.long() if t.dtype == torch.int32 else t


def report(got: torch.Tensor, ref: torch.Tensor):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    d = (got - ref).abs()
    scale = ref.abs().max().item()
    return {"max_abs": d.max().item() if d.numel() else 0.0,
            "scaled": (d.max().item() / scale) if d.numel() and scale > 0 else 0.0,
            "max_rel_floor": (d / (ref.abs() + 1e-8 * scale)).max().item() if d.numel() and scale > 0 else 0.0}


def assert_parity(got, ref, abs_tol=ABS_TOL, rel_tol=REL_TOL, what=""):
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    r = report(got, ref)
    assert r["max_abs"] < abs_tol, (what, r)
    assert r["scaled"] < rel_tol, (what, r)
    return r


def golden_graph(g):
    uid, iid = T(g["uid"]), T(g["iid"])
    return uid, iid, int(g["U"]), int(g["I"])


def ngcf_weights(g, L=3):
    return [(T(g[f"ngcf_w1_{l}"]), T(g[f"ngcf_b1_{l}"]), T(g[f"ngcf_w2_{l}"]), T(g[f"ngcf_b2_{l}"])) for l in range(L)]


def ngcf_masks(g, N, D, L=3):
    bits = np.unpackbits(g["ngcf_masks_packed"])[: L * N * D].reshape(L, N, D).astype(bool)
    return [torch.from_numpy(bits[l]) for l in range(L)]
