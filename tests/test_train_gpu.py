"""GPU: the fused training step — loss kernel, Adam kernel and the whole step — against the oracle's
restatement of ``LightGCN.calculate_loss`` (pinned to the reference's own output in tests/test_oracle.py) driven by
torch autograd + ``torch.optim.Adam`` on the CPU."""
import pytest
import torch

import recbole_gnn_b200 as rg
from recbole_gnn_b200 import train as TR
from oracle import oracle as O
from tests.helpers import T, assert_parity, golden_graph

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("require_pow", [False, True])
@pytest.mark.parametrize("D", [8, 64, 128])
def test_bpr_loss_kernel_matches_autograd(require_pow, D):
    gen = torch.Generator().manual_seed(D)
    U, I, B = 300, 200, 777
    ua, ia = torch.randn(U, D, generator=gen) * 0.3, torch.randn(I, D, generator=gen) * 0.3
    ru, ri = torch.randn(U, D, generator=gen) * 0.1, torch.randn(I, D, generator=gen) * 0.1
    user, pos, neg = (torch.randint(0, n, (B,), generator=gen) for n in (U, I, I))       # duplicates on purpose
    leaves = [t.clone().requires_grad_(True) for t in (ua, ia, ru, ri)]
    a, b, c, d = leaves
    mf = O.bpr_loss((a[user] * b[pos]).sum(1), (a[user] * b[neg]).sum(1))
    loss = mf + 1e-3 * O.emb_loss(c[user], d[pos], d[neg], require_pow=require_pow)
    loss.backward()
    dev = [t.to(DEV) for t in (ua, ia, ru, ri)]
    grads = [torch.zeros_like(t) for t in dev]
    stats = TR.bpr_loss_fused(*dev, user.to(DEV), pos.to(DEV), neg.to(DEV), reg_weight=1e-3, require_pow=require_pow,
                              g_u_all=grads[0], g_i_all=grads[1], g_reg_u=grads[2], g_reg_i=grads[3])
    assert_parity(stats[:2].cpu(), torch.stack([loss.detach().reshape(()), mf.detach()]), abs_tol=1e-5, rel_tol=2e-6)
    for got, leaf in zip(grads, leaves):
        assert_parity(got, leaf.grad, abs_tol=1e-6, rel_tol=1e-5)
    # loss-only call (no gradient tables) gives the same value
    s2 = TR.bpr_loss_fused(*dev, user.to(DEV), pos.to(DEV), neg.to(DEV), reg_weight=1e-3, require_pow=require_pow)
    assert torch.equal(s2[0], stats[0])


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_adam_kernel_matches_torch(wd):
    gen = torch.Generator().manual_seed(1)
    p0 = torch.randn(1000, 64, generator=gen)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-2, weight_decay=wd)
    p = p0.to(DEV)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for t in range(1, 8):
        g = torch.randn(1000, 64, generator=gen) * (0.1 if t % 2 else 3.0)
        ref.grad = g.clone()
        opt.step()
        TR.adam_step(p, g.to(DEV), m, v, lr=1e-2, step=t, weight_decay=wd)
    assert_parity(p, ref.detach(), abs_tol=1e-5, rel_tol=2e-6)


@pytest.mark.parametrize("require_pow", [False, True])
def test_fused_training_step_follows_the_reference_objective(g1, require_pow):
    uid, iid, U, I = golden_graph(g1)
    L, D, lr = 3, 64, 5e-3
    ds = rg.InteractionDataset(uid, iid, U, I, device=DEV)
    m = rg.LightGCN({"device": DEV, "enable_sparse": True, "embedding_size": D, "n_layers": L, "reg_weight": 1e-4,
                     "require_pow": require_pow}, ds).to(DEV)
    with torch.no_grad():
        m.user_embedding.weight.copy_(T(g1["xu"])); m.item_embedding.weight.copy_(T(g1["xi"]))
    step = TR.LightGCNTrainStep(m, lr=lr)
    xu, xi = torch.nn.Parameter(T(g1["xu"]).clone()), torch.nn.Parameter(T(g1["xi"]).clone())
    opt = torch.optim.Adam([xu, xi], lr=lr)
    ei, ew = O.build_norm_adj(uid, iid, U, I)
    gen = torch.Generator().manual_seed(5)
    got, ref = [], []
    for it in range(5):
        k = torch.randperm(uid.numel(), generator=gen)[:512]
        user, pos, neg = uid[k], iid[k], torch.randint(1, I, (512,), generator=gen)
        opt.zero_grad()
        loss = O.lightgcn_loss(xu, xi, ei, ew, L, user, pos, neg, 1e-4, require_pow)
        loss.backward()
        opt.step()
        ref.append(float(loss))
        got.append(float(step.step({"user_id": user.to(DEV), "item_id": pos.to(DEV), "neg_item_id": neg.to(DEV)})))
    assert ref[-1] < ref[0]
    for a, b in zip(got, ref):
        assert abs(a - b) <= 1e-6 * abs(b) + 1e-7, (got, ref)
    assert_parity(m.user_embedding.weight.data, xu.detach(), abs_tol=1e-4, rel_tol=1e-4)
    assert_parity(m.item_embedding.weight.data, xi.detach(), abs_tol=1e-4, rel_tol=1e-4)
