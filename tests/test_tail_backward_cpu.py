"""CPU: the hand-written backward algebra of the fused NGCF layer tail against torch autograd of the reference
composite (layers.py:56-58 + ngcf.py:96-98).  The forward kernel is CUDA-only; its outputs (t, out) are produced
here by the same composite so that only the gradient formulas are under test."""
import pytest
import torch
import torch.nn.functional as F

from recbole_gnn_b200.functional import bignn_tail_backward


@pytest.mark.parametrize("drop,normalize", [(0.0, True), (0.1, True), (0.0, False), (0.3, False)])
def test_tail_backward_matches_autograd(drop, normalize):
    g = torch.Generator().manual_seed(0)
    n, d_in, d_out, slope = 57, 24, 16, 0.2
    p = torch.randn(n, d_in, generator=g, dtype=torch.float64, requires_grad=True)
    x = torch.randn(n, d_in, generator=g, dtype=torch.float64, requires_grad=True)
    w1 = torch.randn(d_out, d_in, generator=g, dtype=torch.float64, requires_grad=True)
    w2 = torch.randn(d_out, d_in, generator=g, dtype=torch.float64, requires_grad=True)
    b1 = torch.randn(d_out, generator=g, dtype=torch.float64, requires_grad=True)
    b2 = torch.randn(d_out, generator=g, dtype=torch.float64, requires_grad=True)
    keep = (torch.rand(n, d_out, generator=g) >= drop) if drop > 0 else None
    ks = keep.double() / (1 - drop) if keep is not None else None
    t = F.linear(p + x, w1, b1) + F.linear(p * x, w2, b2)
    z = F.leaky_relu(t, slope)
    if ks is not None:
        z = z * ks
    out = F.normalize(z, p=2, dim=1) if normalize else z
    g_out = torch.randn(n, d_out, generator=g, dtype=torch.float64)
    out.backward(g_out)
    got = bignn_tail_backward(p.detach(), x.detach(), w1.detach(), w2.detach(), t.detach(), out.detach(), ks, slope,
                              normalize, g_out)
    for name, a, b in zip(("p", "x", "w1", "b1", "w2", "b2"), got, (p.grad, x.grad, w1.grad, b1.grad, w2.grad, b2.grad)):
        assert torch.allclose(a, b, rtol=1e-9, atol=1e-11), name


def test_tail_backward_zero_rows():
    """[PAD]/isolated rows: t = b (non-zero) in general, but an all-zero z must not produce NaN."""
    n, d = 5, 8
    p = torch.zeros(n, d, dtype=torch.float64); x = torch.zeros(n, d, dtype=torch.float64)
    w = torch.zeros(d, d, dtype=torch.float64)
    t = torch.zeros(n, d, dtype=torch.float64); out = torch.zeros(n, d, dtype=torch.float64)
    got = bignn_tail_backward(p, x, w, w, t, out, None, 0.2, True, torch.ones(n, d, dtype=torch.float64))
    assert all(torch.isfinite(v).all() for v in got)
