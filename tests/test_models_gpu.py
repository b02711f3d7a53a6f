"""GPU parity of the model-level routes: ``calculate_loss`` of the three drop-in models against the goldens produced
by the reference's own ``calculate_loss`` (tests/golden/make_golden.py executes lightgcn.py / simgcl.py / ngcf.py
unmodified), the call shapes of the callers that inherit the layers (SGL, NGCF node dropout), the `.inter` ingest
route, and NGCF / SimGCL at a size (200 k x 200 k, 20 M interactions) where the fixture-sized cases say nothing
about grids, hub plans or 32-bit offsets."""
import os

import pytest
import torch

import recbole_gnn_b200 as rg
from recbole_gnn_b200 import functional as F_
from oracle import oracle as O
from tests.helpers import T, assert_parity, golden_graph, ngcf_masks, ngcf_weights

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _batch(g):
    return {"user_id": T(g["batch_user"]).to(DEV), "item_id": T(g["batch_pos"]).to(DEV),
            "neg_item_id": T(g["batch_neg"]).to(DEV)}


def _load_tables(m, xu, xi):
    with torch.no_grad():
        m.user_embedding.weight.copy_(xu)
        m.item_embedding.weight.copy_(xi)


# ------------------------------------------------------------------------------------------ losses vs goldens
@pytest.mark.parametrize("name", ["g1", "g2"])
@pytest.mark.parametrize("fused", [True, False])
def test_lightgcn_calculate_loss_matches_reference(name, fused, request):
    g = request.getfixturevalue(name)
    uid, iid, U, I = golden_graph(g)
    ds = rg.InteractionDataset(uid, iid, U, I, device=DEV)
    for rp, tag in ((False, "nopow"), (True, "pow")):
        m = rg.LightGCN({"device": DEV, "enable_sparse": True, "embedding_size": g["xu"].shape[1], "n_layers": 3,
                         "reg_weight": 1e-5, "require_pow": rp, "fused_propagation": fused}, ds).to(DEV)
        _load_tables(m, T(g["xu"]), T(g["xi"]))
        loss = m.calculate_loss(_batch(g))
        loss.backward()
        assert_parity(loss.detach().reshape(1), T(g[f"lightgcn_loss_{tag}"]), abs_tol=1e-6, rel_tol=2e-6)
        assert_parity(m.user_embedding.weight.grad, T(g[f"lightgcn_loss_{tag}_gu"]), abs_tol=1e-7, rel_tol=1e-5)
        assert_parity(m.item_embedding.weight.grad, T(g[f"lightgcn_loss_{tag}_gi"]), abs_tol=1e-7, rel_tol=1e-5)


@pytest.mark.parametrize("name", ["g1", "g2"])
@pytest.mark.parametrize("fused", [True, False])
def test_simgcl_calculate_loss_matches_reference(name, fused, request):
    """simgcl.py:48-60 incl. the InfoNCE term over the two perturbed views (it was silently dropped)."""
    g = request.getfixturevalue(name)
    uid, iid, U, I = golden_graph(g)
    ds = rg.InteractionDataset(uid, iid, U, I, device=DEV)
    m = rg.SimGCL({"device": DEV, "enable_sparse": True, "embedding_size": g["xu"].shape[1], "n_layers": 3, "reg_weight": 1e-5,
                   "lambda": 0.1, "eps": 0.1, "temperature": 0.2, "fused_propagation": fused}, ds).to(DEV)
    _load_tables(m, T(g["xu"]), T(g["xi"]))
    nz = (T(g["simgcl_loss_noise_u8"]).float() / 256.0).to(DEV)
    loss = m.calculate_loss(_batch(g), noises1=[nz[l] for l in range(3)], noises2=[nz[l] for l in range(3, 6)])
    loss.backward()
    assert_parity(loss.detach().reshape(1), T(g["simgcl_loss"]), abs_tol=1e-4, rel_tol=3e-6)
    assert_parity(m.user_embedding.weight.grad, T(g["simgcl_loss_gu"]), abs_tol=1e-4, rel_tol=1e-5)
    assert_parity(m.item_embedding.weight.grad, T(g["simgcl_loss_gi"]), abs_tol=1e-4, rel_tol=1e-5)
    # default route: in-kernel Philox noise, finite and different from the clean-only objective
    m.zero_grad()
    l2 = m.calculate_loss(_batch(g))
    l2.backward()
    assert torch.isfinite(l2) and torch.isfinite(m.user_embedding.weight.grad).all()
    assert abs(float(l2) - float(T(g["lightgcn_loss_nopow"]))) > 1e-3


def _ngcf(g, ds, **kw):
    D = g["ngcf_xu"].shape[1]
    cfg = {"device": DEV, "enable_sparse": True, "embedding_size": D, "hidden_size_list": [D, D, D],
           "node_dropout": 0.0, "message_dropout": 0.0, "reg_weight": 1e-5}
    cfg.update(kw)
    m = rg.NGCF(cfg, ds).to(DEV)
    W = ngcf_weights(g)
    _load_tables(m, T(g["ngcf_xu"]), T(g["ngcf_xi"]))
    with torch.no_grad():
        for l, layer in enumerate(m.GNNlayers):
            layer.lin1.weight.copy_(W[l][0]); layer.lin1.bias.copy_(W[l][1])
            layer.lin2.weight.copy_(W[l][2]); layer.lin2.bias.copy_(W[l][3])
    return m


@pytest.mark.parametrize("name", ["g1", "g2"])
def test_ngcf_calculate_loss_predict_and_full_sort_match_reference(name, request):
    g = request.getfixturevalue(name)
    uid, iid, U, I = golden_graph(g)
    ds = rg.InteractionDataset(uid, iid, U, I, device=DEV)
    m = _ngcf(g, ds)
    loss = m.calculate_loss(_batch(g))
    loss.backward()
    assert_parity(loss.detach().reshape(1), T(g["ngcf_loss_p0"]), abs_tol=1e-6, rel_tol=3e-6)
    assert_parity(m.user_embedding.weight.grad, T(g["ngcf_loss_p0_gu"]), abs_tol=1e-6, rel_tol=1e-4)
    assert_parity(m.GNNlayers[0].lin1.weight.grad, T(g["ngcf_loss_p0_gw1_0"]), abs_tol=1e-6, rel_tol=1e-4)
    assert_parity(m.GNNlayers[2].lin2.weight.grad, T(g["ngcf_loss_p0_gw2_2"]), abs_tol=1e-6, rel_tol=1e-4)
    assert_parity(m.GNNlayers[1].lin1.bias.grad, T(g["ngcf_loss_p0_gb1_1"]), abs_tol=1e-6, rel_tol=1e-4)
    # predict / full_sort_predict (ngcf.py:125-150) on the golden forward
    ref = T(g["ngcf_p0"])
    ru, ri = ref[:U], ref[U:]
    b = _batch(g)
    with torch.no_grad():
        s = m.predict(b)
        assert_parity(s, (ru[b["user_id"].cpu()] * ri[b["item_id"].cpu()]).sum(1), rel_tol=1e-5)
        users = b["user_id"][:7]
        full = m.full_sort_predict({"user_id": users})
        assert_parity(full, (ru[users.cpu()] @ ri.t()).reshape(-1), rel_tol=1e-5)
        assert m.restore_user_e is not None
    m.calculate_loss(b)
    assert m.restore_user_e is None                                        # ngcf.py:108-109


def test_ngcf_hidden_sizes_outside_the_fused_tail_train_and_evaluate(g1):
    """hidden sizes the tail kernel does not take (> 256 or not a multiple of 4) must work on BOTH routes."""
    uid, iid, U, I = golden_graph(g1)
    ds = rg.InteractionDataset(uid, iid, U, I, device=DEV)
    m = rg.NGCF({"device": DEV, "enable_sparse": True, "embedding_size": 64, "hidden_size_list": [320, 64],
                 "message_dropout": 0.0}, ds).to(DEV)
    u, i = m.forward()
    with torch.no_grad():
        u2, i2 = m.forward()
    assert u.shape == (U, 64 + 320 + 64)
    assert_parity(u2, u.detach(), rel_tol=1e-5)
    assert_parity(i2, i.detach(), rel_tol=1e-5)


def test_ngcf_node_dropout_route_matches_oracle(g1):
    """ngcf.py:74-90: dropout_adj on the edges (Bernoulli keep, no rescale), then the layer stack on the thinned
    graph — with the keep flags handed to both sides."""
    uid, iid, U, I = golden_graph(g1)
    ds = rg.InteractionDataset(uid, iid, U, I, device=DEV)
    m = _ngcf(g1, ds, node_dropout=0.3, message_dropout=0.1)
    m.train()
    N, D = U + I, 64
    masks = ngcf_masks(g1, N, D)
    m.eval()
    row, col, val = m._graph().coo()           # (dst, src, w) in CSR order == the order keep_edges indexes
    m.train()
    gen = torch.Generator().manual_seed(11)
    keep = torch.rand(row.numel(), generator=gen) >= 0.3
    ei = torch.stack([col.cpu(), row.cpu()])
    ei_k, ew_k = O.dropout_adj(ei, val.cpu(), keep)
    ref_u, ref_i = O.ngcf_forward(T(g1["ngcf_xu"]), T(g1["ngcf_xi"]), ei_k, ew_k, ngcf_weights(g1), 0.1, masks)
    for grad in (True, False):
        with torch.set_grad_enabled(grad):
            u, i = m.forward(keep_masks=[k.to(DEV) for k in masks], keep_edges=keep.to(DEV))
        assert_parity(u, ref_u, rel_tol=5e-6, what=f"grad={grad}")
        assert_parity(i, ref_i, rel_tol=5e-6, what=f"grad={grad}")
    u, i = m.forward()                           # own draws
    assert torch.isfinite(u).all() and u.shape == (U, 256)


# ------------------------------------------------------------------------------------------ inherited call shapes
@pytest.mark.parametrize("name", ["g1", "g2"])
def test_sgl_call_shape_adj_t_gcn_norm_to_device(name, request):
    """sgl.py:120-122: ``dataset.edge_index_to_adj_t(edge_index, edge_weight, N, N)`` -> ``gcn_norm(adj_t, None, N,
    add_self_loops=False)`` -> ``.to(device)``, then LightGCNConv on the result; and the dense-edge twin
    sgl.py:124 ``gcn_norm(edge_index, edge_weight, N, add_self_loops=False)``."""
    g = request.getfixturevalue(name)
    uid, iid, U, I = golden_graph(g)
    N = U + I
    row, col = uid, iid + U
    edge_index = torch.cat([torch.stack([row, col]), torch.stack([col, row])], dim=1)     # sgl.py:113-115
    edge_weight = torch.ones(edge_index.size(1))
    adj_t = rg.GeneralGraphDataset.edge_index_to_adj_t(edge_index, edge_weight, N, N)
    adj_t = rg.gcn_norm(adj_t, None, N, add_self_loops=False)
    adj_t = adj_t.to(DEV)
    r, c, v = adj_t.coo()
    assert torch.equal(r.cpu(), T(g["adj_row"])) and torch.equal(c.cpu(), T(g["adj_col"]))
    assert torch.equal(v.cpu(), T(g["adj_val"]))
    x = torch.cat([T(g["xu"]), T(g["xi"])]).to(DEV)
    conv = rg.LightGCNConv(x.size(1))
    assert_parity(conv(x, adj_t, None), T(g["prop_sparse"]), rel_tol=2e-6)
    ei2, ew2 = rg.gcn_norm(edge_index.to(DEV), edge_weight.to(DEV), N, add_self_loops=False)
    assert torch.equal(ew2.cpu(), T(g["edge_weight"]))
    assert_parity(conv(x, ei2, ew2), T(g["prop_dense"]), rel_tol=2e-6)
    # the MessagePassing surface the reference's layers define (layers.py:14-20)
    assert_parity(conv.propagate(adj_t, x=x, edge_weight=None), T(g["prop_sparse"]), rel_tol=2e-6)
    assert_parity(conv.message_and_aggregate(adj_t, x), T(g["prop_sparse"]), rel_tol=2e-6)
    xj = x[ei2[0]]
    assert torch.equal(conv.message(xj, ew2), ew2.view(-1, 1) * xj)


def test_inter_file_to_propagation_on_device(tmp_path, g1):
    """§8f-4: `.inter` atomic file -> ids -> device CSR build -> propagation, against the goldens of the same fixture
    (tokens written back from the golden id columns; the reference file itself does not travel to the GPU box)."""
    p = tmp_path / "fixture.inter"
    uid, iid, U, I = golden_graph(g1)
    with open(p, "w") as f:
        f.write("user_id:token\titem_id:token\trating:float\ttimestamp:float\n")
        for u, i in zip(uid.tolist(), iid.tolist()):
            f.write(f"u{u * 7 + 3}\ti{i * 5 + 1}\t3\t881250949\n")
    ds = rg.InteractionDataset.from_inter_file(str(p), device=DEV)
    assert (ds.user_num, ds.item_num) == (U, I)
    assert torch.equal(ds.inter_feat["user_id"], uid) and torch.equal(ds.inter_feat["item_id"], iid)
    m = rg.LightGCN({"device": DEV, "enable_sparse": True, "embedding_size": 64, "n_layers": 3}, ds).to(DEV)
    _load_tables(m, T(g1["xu"]), T(g1["xi"]))
    with torch.no_grad():
        u, i = m.forward()
    assert_parity(torch.cat([u, i]), T(g1["lightgcn_L3"]), rel_tol=2e-6)
    r, c, v = m._graph().coo()
    assert torch.equal(v.cpu(), T(g1["adj_val"]))


def test_layer_cache_drops_dead_graphs():
    """resident CSRs built from raw edge tensors die with the tensor, not after 16 newer graphs."""
    import gc
    from recbole_gnn_b200 import layers
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(500, 16, generator=gen).to(DEV)
    conv = rg.LightGCNConv(16)
    layers._CACHE.clear()
    for k in range(5):
        ei = torch.randint(0, 500, (2, 4000), generator=gen).to(DEV)
        ew = torch.rand(4000, generator=gen).to(DEV)
        y = conv(x, ei, ew)
        assert_parity(y, O.propagate_scatter(x.cpu(), ei.cpu(), ew.cpu()), rel_tol=5e-6)
        assert len(layers._CACHE) == 1
        del ei, ew
        gc.collect()
        assert len(layers._CACHE) == 0
    with pytest.raises(NotImplementedError):
        conv(x, torch.randint(0, 500, (2, 10)).to(DEV), torch.rand(10, device=DEV, requires_grad=True))


# ------------------------------------------------------------------------------------------ medium size
@pytest.fixture(scope="module")
def medium():
    U = I = 200_000
    E = 20_000_000
    if os.environ.get("B200GCN_TEST_SMALL"):
        U = I = 20_000
        E = 1_000_000
    uid, iid = O.synth_interactions(U, I, E, seed=0)
    ei, ew = O.build_norm_adj(uid, iid, U, I)
    a = O.adj_sparse(ei, ew, U + I, U + I, layout="csr")
    ds = rg.InteractionDataset(uid, iid, U, I, device=DEV)
    h, _ = ds.get_norm_adj_mat(enable_sparse=True)
    return {"U": U, "I": I, "a": a, "h": h.to(DEV), "ds": ds}


def test_ngcf_medium_matches_oracle(medium):
    """BASELINE config 3 shape at 1/5 scale: 3 NGCF layers with the always-on message dropout (mask supplied)."""
    U, I, a = medium["U"], medium["I"], medium["a"]
    N, D, p = U + I, 64, 0.1
    xu, xi = O.xavier_normal_((U, D), 1), O.xavier_normal_((I, D), 2)
    W = [(O.xavier_normal_((D, D), 10 + l), torch.zeros(D).uniform_(-0.05, 0.05, generator=torch.Generator().manual_seed(20 + l)),
          O.xavier_normal_((D, D), 30 + l), torch.zeros(D).uniform_(-0.05, 0.05, generator=torch.Generator().manual_seed(40 + l)))
         for l in range(3)]
    gen = torch.Generator().manual_seed(5)
    masks = [torch.rand(N, D, generator=gen) >= p for _ in range(3)]
    # oracle with the CSR torch.sparse.mm form of the propagation (layers.py:19-20), same layer algebra
    x = torch.cat([xu, xi])
    ref = [x]
    for l, (w1, b1, w2, b2) in enumerate(W):
        pr = O.propagate_sparse(a, x)
        x = torch.nn.functional.linear(pr + x, w1, b1) + torch.nn.functional.linear(pr * x, w2, b2)
        x = torch.nn.functional.leaky_relu(x, 0.2) * masks[l] / (1 - p)
        x = torch.nn.functional.normalize(x, p=2, dim=1)
        ref.append(x)
    ref = torch.cat(ref, 1)
    Wd = [tuple(t.to(DEV) for t in w) for w in W]
    u, i = F_.ngcf_forward(medium["h"], xu.to(DEV), xi.to(DEV), Wd, message_dropout=p,
                           keep_masks=[m.to(DEV) for m in masks])
    assert_parity(torch.cat([u, i]), ref, rel_tol=1e-5)


def test_simgcl_three_views_medium_match_oracle(medium):
    """BASELINE config 4 shape at 1/5 scale: clean + two perturbed views (noise supplied), shared first layer."""
    U, I, a = medium["U"], medium["I"], medium["a"]
    N, D, L, eps = U + I, 64, 3, 0.1
    xu, xi = O.xavier_uniform_table(U, D, 1), O.xavier_uniform_table(I, D, 2)
    gen = torch.Generator().manual_seed(9)
    nz = [[torch.rand(N, D, generator=gen) for _ in range(L)] for _ in range(2)]

    def ref_view(noises):
        """Oracle view + the mask of output elements that depend on a sign(e) decision the two sides may legitimately
        take differently: sign() is discontinuous at 0, |e| below fp32 re-association noise occurs ~1e-7 of the time,
        and a flipped element of layer l reaches column d of its neighbours' rows in every later layer."""
        e, acc = torch.cat([xu, xi]), 0
        tainted = torch.zeros(N, D)
        for l in range(L):
            e = O.propagate_sparse(a, e)
            tainted = (O.propagate_sparse(a_bool, tainted) > 0).float()
            if noises is not None:
                tainted = torch.maximum(tainted, (e.abs() < 2e-7 * e.abs().max()).float() * (e != 0).float())
                e = e + torch.sign(e) * torch.nn.functional.normalize(noises[l], dim=-1) * eps
            acc = acc + e
            tainted_out = tainted if l == 0 else torch.maximum(tainted_out, tainted)
        return acc / L, tainted_out.bool()

    crow, ccol = a.crow_indices(), a.col_indices()
    a_bool = torch.sparse_csr_tensor(crow, ccol, torch.ones(ccol.numel()), size=(N, N))
    views = F_.simgcl_views(medium["h"], xu.to(DEV), xi.to(DEV), L, eps,
                            noises1=[t.to(DEV) for t in nz[0]], noises2=[t.to(DEV) for t in nz[1]])
    for (u, i), noises in zip(views, (None, nz[0], nz[1])):
        ref, skip = ref_view(noises)
        assert skip.float().mean().item() < 0.05 and (noises is not None or not skip.any())
        got = torch.cat([u, i]).cpu()
        assert_parity(torch.where(skip, ref, got), ref, rel_tol=1e-5)


def test_dispatcher_ops_match_and_compile(g1):
    """`torch.ops.b200gcn.*` over the CSR of a resident handle: same numbers as the layer API, autograd, and usable
    inside torch.compile(fullgraph=True) (the op is opaque to the tracer; its fake kernel supplies the shapes)."""
    import recbole_gnn_b200.ops  # noqa: F401
    uid, iid, U, I = golden_graph(g1)
    ds = rg.InteractionDataset(uid, iid, U, I, device=DEV)
    h, _ = ds.get_norm_adj_mat(enable_sparse=True)
    rowptr, col, val = h.to(DEV).csr()
    xu, xi = T(g1["xu"]).to(DEV), T(g1["xi"]).to(DEV)
    x = torch.cat([xu, xi]).requires_grad_(True)
    y = torch.ops.b200gcn.spmm(rowptr, col, val, x, U + I, True)
    assert_parity(y, T(g1["prop_sparse"]), rel_tol=2e-6)
    gen = torch.Generator().manual_seed(3)
    gy = torch.randn(U + I, 64, generator=gen)
    y.backward(gy.to(DEV))
    ei, ew = O.build_norm_adj(uid, iid, U, I)
    assert_parity(x.grad, O.propagate_scatter(gy, ei.flip([0]), ew), rel_tol=5e-6)
    u, i = torch.ops.b200gcn.lightgcn_propagate(rowptr, col, val, xu, xi, 3)
    assert_parity(torch.cat([u, i]), T(g1["lightgcn_L3"]), rel_tol=2e-6)

    def f(a, b):
        u, i = torch.ops.b200gcn.lightgcn_propagate(rowptr, col, val, a, b, 3)
        return (u * 2).sum() + (i * 3).sum()

    a, b = xu.clone().requires_grad_(True), xi.clone().requires_grad_(True)
    fc = torch.compile(f, fullgraph=True, backend="aot_eager")   # tracing + autograd capture; no code generation
    out = fc(a, b)
    out.backward()
    a2, b2 = xu.clone().requires_grad_(True), xi.clone().requires_grad_(True)
    ref = f(a2, b2)
    ref.backward()
    assert torch.allclose(out, ref, rtol=1e-6)
    assert_parity(a.grad, a2.grad, rel_tol=1e-6)
    assert_parity(b.grad, b2.grad, rel_tol=1e-6)
