"""GPU: the per-epoch graph re-sampling of SGL (ND / ED / RW), SEPT and the per-layer graph lists they feed to the conv
layer, on the device, against the oracle's restatement with the same sampled indices."""
import pytest
import torch

import recbole_gnn_b200 as rg
from recbole_gnn_b200 import augment as A
from oracle import oracle as O
from tests.helpers import T, assert_parity, golden_graph

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _same_graph(h, ei, ew, n):
    """Resident handle == oracle COO: same entry multiset (bit-exact weights) in the (dst, src)-sorted order."""
    row, col, val = h.coo()
    key = ei[1] * n + ei[0]
    order = torch.argsort(key, stable=True)
    assert torch.equal(row.cpu(), ei[1][order]) and torch.equal(col.cpu(), ei[0][order])
    assert torch.equal(val.cpu(), ew[order])


@pytest.mark.parametrize("name", ["g1", "g2"])
def test_sgl_node_and_edge_dropout_match_oracle(name, request):
    g = request.getfixturevalue(name)
    uid, iid, U, I = golden_graph(g)
    n, E = U + I, uid.numel()
    aug = A.SGLAugmenter(uid, iid, U, I, DEV)
    gen = torch.Generator().manual_seed(4)
    x = torch.cat([T(g["xu"]), T(g["xi"])])
    conv = rg.LightGCNConv(x.size(1))
    # ND: dropped ids handed to both sides (sgl.py:95-104)
    du = torch.randperm(U, generator=gen)[: int(U * 0.3)]
    di = torch.randperm(I, generator=gen)[: int(I * 0.3)]
    h = aug.node_dropout(du, di)
    ei, ew = O.sgl_augmented_adj(uid, iid, U, I, "ND", drop_user=du, drop_item=di)
    _same_graph(h, ei, ew, n)
    assert h.is_symmetric
    assert_parity(conv(x.to(DEV), h, None), O.propagate_scatter(x, ei, ew), rel_tol=2e-6)
    # ED: kept interaction indices handed to both sides (sgl.py:106-109)
    keep = torch.randperm(E, generator=gen)[: int(E * 0.8)]
    h = aug.edge_dropout(keep)
    ei, ew = O.sgl_augmented_adj(uid, iid, U, I, "ED", keep_idx=keep)
    _same_graph(h, ei, ew, n)
    assert_parity(conv(x.to(DEV), h, None), O.propagate_scatter(x, ei, ew), rel_tol=2e-6)
    # own draws: exact counts, like np.random.choice(replace=False)
    h, w = aug.random_graph_augment("ED", 0.25)
    assert w is None and h.nnz() == 2 * int(E * 0.75)
    h, _ = aug.random_graph_augment("ND", 0.25)
    assert h.nnz() <= 2 * E and h.nnz() % 2 == 0


def test_sgl_random_walk_forward_per_layer_graphs(g1):
    """aug_type RW: a different graph per layer (sgl.py:88-90,136-139)."""
    uid, iid, U, I = golden_graph(g1)
    E, L = uid.numel(), 3
    aug = A.SGLAugmenter(uid, iid, U, I, DEV)
    gen = torch.Generator().manual_seed(8)
    keeps = [torch.randperm(E, generator=gen)[: int(E * 0.9)] for _ in range(L)]
    graphs = [(aug.edge_dropout(k), None) for k in keeps]
    xu, xi = T(g1["xu"]), T(g1["xi"])
    u, i = A.sgl_forward(rg.LightGCNConv(64), xu.to(DEV), xi.to(DEV), graphs)
    x = torch.cat([xu, xi])
    outs = [x]
    for k in keeps:
        ei, ew = O.sgl_augmented_adj(uid, iid, U, I, "RW", keep_idx=k)
        x = O.propagate_scatter(x, ei, ew)
        outs.append(x)
    ref = torch.stack(outs, 1).mean(1)
    assert_parity(torch.cat([u, i]), ref, rel_tol=2e-6)
    g1_, g2_ = aug.graph_construction("RW", 0.1, L)
    assert len(g1_) == L and g1_[0][0] is not g1_[1][0]
    g1_, g2_ = aug.graph_construction("ED", 0.1, L)
    assert g1_[0][0] is g1_[2][0] and g1_[0][0] is not g2_[0][0]


def test_sept_subgraph_matches_oracle(g1):
    uid, iid, U, I = golden_graph(g1)
    gen = torch.Generator().manual_seed(2)
    S = 600
    src, tgt = torch.randint(1, U, (S,), generator=gen), torch.randint(1, U, (S,), generator=gen)
    keep = torch.randperm(uid.numel(), generator=gen)[: int(uid.numel() * 0.8)]
    net_keep = torch.randperm(S, generator=gen)[: int(S * 0.8)]
    ei, ew = A.sept_subgraph_construction(uid, iid, src, tgt, U, I, 0.2, DEV, keep=keep, net_keep=net_keep)
    ei_ref, ew_ref = O.sept_subgraph(uid, iid, src, tgt, U, I, keep, net_keep)
    assert torch.equal(ei.cpu(), ei_ref)
    assert_parity(ew, ew_ref, abs_tol=1e-7, rel_tol=2e-7)
    x = torch.cat([T(g1["xu"]), T(g1["xi"])])
    y = rg.LightGCNConv(64)(x.to(DEV), ei, ew)                    # sept.py:173,177: the joint (non-symmetric) graph
    assert_parity(y, O.propagate_scatter(x, ei_ref, ew_ref), rel_tol=2e-6)
    ei2, ew2 = A.sept_subgraph_construction(uid, iid, src, tgt, U, I, 0.2, DEV)
    assert ei2.size(1) == 2 * int(uid.numel() * 0.8) + int(S * 0.8) and ew2.numel() == ei2.size(1)
