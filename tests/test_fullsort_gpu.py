"""GPU: full-sort evaluation — the tcgen05 score contraction and the fused top-k — against
``torch.topk(torch.matmul(...))`` in float64 with RecBole's masking ([PAD] item, seen items)."""
import pytest
import torch

import recbole_gnn_b200 as rg
from recbole_gnn_b200 import functional as F_
from tests.helpers import T, assert_parity, golden_graph

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _reference(u, items, k, history, first_item):
    s = u.double() @ items.double().t()
    if history is not None:
        s[history[0], history[1]] = float("-inf")
    s[:, :first_item] = float("-inf")
    return s, torch.topk(s, k, dim=1)


@pytest.mark.parametrize("B,I,D,k", [(1, 5, 8, 3), (77, 1000, 64, 10), (300, 20_011, 64, 50), (129, 3000, 32, 64),
                                      (64, 5000, 128, 20), (40, 4097, 256, 10), (16, 2000, 52, 5)])
def test_fullsort_topk_and_scores(B, I, D, k):
    gen = torch.Generator().manual_seed(B * 31 + D)
    u, items = torch.randn(B, D, generator=gen) * 0.3, torch.randn(I, D, generator=gen) * 0.3
    H = min(I // 3, 40)
    rows = torch.arange(B).repeat_interleave(H)
    its = torch.stack([torch.randperm(I, generator=gen)[:H] for _ in range(B)]).flatten()
    # plant the best item of every user into its history: the mask must remove it
    best = (u.double() @ items.double().t())[:, 1:].argmax(1) + 1
    rows, its = torch.cat([rows, torch.arange(B)]), torch.cat([its, best])
    key = torch.unique(rows * I + its)
    rows, its = key // I, key % I
    ud, itd = u.to(DEV), items.to(DEV)
    dense = F_.full_sort_scores(ud, itd)
    assert_parity(dense, (u.double() @ items.double().t()).float(), rel_tol=5e-6, what="dense scores")
    for history in (None, (rows, its)):
        s_ref, (v_ref, i_ref) = _reference(u, items, k, history, 1)
        kk = min(k, int(torch.isfinite(s_ref).sum(1).min()))
        scores, ids = F_.full_sort_topk(ud, itd, k, history=None if history is None else (rows.to(DEV), its.to(DEV)))
        scores, ids = scores.cpu(), ids.cpu()
        assert_parity(scores[:, :kk], v_ref[:, :kk].float(), rel_tol=5e-6, what="top-k scores")
        # the ids are the items that HAVE those scores (robust to the order of exact ties) and none is masked
        got = torch.gather(s_ref, 1, ids[:, :kk].clamp_min(0))
        assert_parity(got.float(), v_ref[:, :kk].float(), rel_tol=5e-6, what="scores at the returned ids")
        assert (ids[:, :kk] >= 1).all() and (ids[:, :kk] < I).all()
        assert all(len(set(r.tolist())) == kk for r in ids[:, :kk])
        if kk < k:
            assert (ids[:, kk:] == -1).all() and torch.isinf(scores[:, kk:]).all()
        assert (scores[:, :-1] >= scores[:, 1:]).all()
        if k > 1:
            assert (ids[:, :kk] == i_ref[:, :kk]).float().mean() > 0.99     # exact ties aside, the very same ids


def test_fullsort_topk_beyond_the_fused_list():
    """k > 64: the dense-scores kernel + library selection, same masking."""
    gen = torch.Generator().manual_seed(7)
    B, I, D, k = 33, 1500, 64, 100
    u, items = torch.randn(B, D, generator=gen) * 0.3, torch.randn(I, D, generator=gen) * 0.3
    rows = torch.arange(B).repeat_interleave(10)
    its = torch.randint(1, I, (B * 10,), generator=gen)
    s_ref, (v_ref, i_ref) = _reference(u, items, k, (rows, its), 1)
    scores, ids = F_.full_sort_topk(u.to(DEV), items.to(DEV), k, history=(rows.to(DEV), its.to(DEV)))
    assert_parity(scores, v_ref.float(), rel_tol=5e-6)
    assert_parity(torch.gather(s_ref, 1, ids.cpu()).float(), v_ref.float(), rel_tol=5e-6)


def test_model_full_sort_routes(g1):
    uid, iid, U, I = golden_graph(g1)
    ds = rg.InteractionDataset(uid, iid, U, I, device=DEV)
    m = rg.LightGCN({"device": DEV, "enable_sparse": True, "embedding_size": 64, "n_layers": 3}, ds).to(DEV)
    with torch.no_grad():
        m.user_embedding.weight.copy_(T(g1["xu"])); m.item_embedding.weight.copy_(T(g1["xi"]))
        users = torch.arange(1, 200, device=DEV)
        full = m.full_sort_predict({"user_id": users}).view(users.numel(), I)
        ref = T(g1["lightgcn_L3"])
        assert_parity(full, ref[:U][users.cpu()] @ ref[U:].t(), rel_tol=5e-6)
        # RecBole's evaluator: mask the training interactions of the batch users, then top-k
        sel = (uid >= 1) & (uid < 200)
        hist = (uid[sel] - 1, iid[sel])
        scores, ids = m.full_sort_topk({"user_id": users}, 20, history=(hist[0].to(DEV), hist[1].to(DEV)))
        s = full.clone()
        s[hist[0].to(DEV), hist[1].to(DEV)] = float("-inf")
        s[:, 0] = float("-inf")
        v_ref, i_ref = torch.topk(s, 20, dim=1)
        assert_parity(scores, v_ref, rel_tol=5e-6)
        assert (ids == i_ref).float().mean() > 0.99
