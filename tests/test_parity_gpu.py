"""GPU parity tests: the CUDA path (through the C ABI in libb200gcn.so) against the CPU oracle on the
same seeded inputs and against the committed golden fixtures.  Tolerance: per-element |delta| < 1e-4
absolute (BASELINE.json north_star) AND max|delta| / max|ref| < 1e-5 (the absolute bound is
nearly vacuous for xavier-initialised tables, so the scaled bound is the one that bites)."""
import pytest
import torch

import recbole_gnn_b200 as rg
from recbole_gnn_b200 import functional as F_
from oracle import oracle as O
from tests.helpers import T, assert_parity, golden_graph, ngcf_masks, ngcf_weights

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _handle(uid, iid, U, I):
    ds = rg.InteractionDataset(uid, iid, U, I, device=DEV)
    h, w = ds.get_norm_adj_mat(enable_sparse=True)
    assert w is None
    return h.to(DEV)


# ------------------------------------------------------------------------------------------ graph build
@pytest.mark.parametrize("name", ["g1", "g2"])
def test_norm_adj_build_matches_golden(name, request):
    g = request.getfixturevalue(name)
    uid, iid, U, I = golden_graph(g)
    h = _handle(uid, iid, U, I)
    row, col, val = h.coo()
    # golden adj_t.coo() is sorted by (row, col) with duplicates adjacent — the engine's CSR order
    assert torch.equal(row.cpu(), T(g["adj_row"]))
    assert torch.equal(col.cpu(), T(g["adj_col"]))
    assert torch.equal(val.cpu(), T(g["adj_val"]))                 # bit-exact weights
    # dense-edge return: reference edge order, bit-exact weights
    ds = rg.InteractionDataset(uid, iid, U, I, device=DEV)
    ei, ew = ds.get_norm_adj_mat(enable_sparse=False)
    assert ei.is_cuda and torch.equal(ei.cpu(), T(g["edge_index"]))
    assert torch.equal(ew.cpu(), T(g["edge_weight"]))


@pytest.mark.parametrize("name", ["g1", "g2"])
def test_bipartite_inter_mat_and_conv_match_golden(name, request):
    g = request.getfixturevalue(name)
    uid, iid, U, I = golden_graph(g)
    ds = rg.InteractionDataset(uid, iid, U, I, device=DEV)
    xu, xi = T(g["xu"]).to(DEV), T(g["xi"]).to(DEV)
    conv = rg.BipartiteGCNConv(xu.size(1))
    for row in ("user", "item"):
        for rn in (True, False):
            tag = f"bip_{row}_{'rown' if rn else 'sym'}"
            ei, ew = ds.get_bipartite_inter_mat(row=row, row_norm=rn)
            assert torch.equal(ei.cpu(), T(g[tag + "_ei"]))
            assert torch.equal(ew.cpu(), T(g[tag + "_ew"]))        # bit-exact
            n_row, n_col = (U, I) if row == "user" else (I, U)
            x_row, x_col = (xu, xi) if row == "user" else (xi, xu)
            y = conv((x_col, x_row), ei.flip([0]), ew, size=(n_col, n_row))   # diffnet.py:97 call shape
            assert_parity(y, T(g[tag + "_y"]), what=tag)


def test_csr_from_coo_general_transpose_and_mask():
    gen = torch.Generator().manual_seed(11)
    n_dst, n_src, nnz = 700, 300, 20000
    src = torch.randint(0, n_src, (nnz,), generator=gen)
    dst = torch.randint(0, n_dst, (nnz,), generator=gen)
    dst[dst == 5] = 6                                               # an empty row
    w = torch.rand(nnz, generator=gen)
    x = torch.randn(n_src, 32, generator=gen)
    ei = torch.stack([src, dst])
    ref = O.propagate_scatter(x, ei, w, n_dst)
    h = rg.GraphHandle(row=dst, col=src, value=w, sparse_sizes=(n_dst, n_src)).to(DEV)
    y = F_.spmm(h, x.to(DEV))
    assert_parity(y, ref, rel_tol=2e-6)
    rowptr, col, val = h.csr()
    assert rowptr[0] == 0 and rowptr[-1] == nnz and bool((rowptr[1:] >= rowptr[:-1]).all())
    assert rowptr[5] == rowptr[6]
    # entries sorted by (row, col)
    r, c, v = h.coo()
    key = r * n_src + c
    assert bool((key[1:] >= key[:-1]).all())
    # transpose: A^T g
    gy = torch.randn(n_dst, 32, generator=gen)
    ref_t = O.propagate_scatter(gy, ei.flip([0]), w, n_src)
    ht = h.t()
    assert ht.sparse_sizes() == (n_src, n_dst) and ht.t() is h
    assert_parity(F_.spmm(ht, gy.to(DEV)), ref_t, rel_tol=2e-6)
    # edge masking == dropout_adj on the same entries
    keep = torch.rand(nnz, generator=gen) > 0.25
    hm = h.masked(keep.to(DEV))
    rc, cc, vc = r.cpu(), c.cpu(), v.cpu()
    e2, w2 = O.dropout_adj(torch.stack([cc, rc]), vc, keep)
    assert hm.nnz() == int(keep.sum())
    assert_parity(F_.spmm(hm, x.to(DEV)), O.propagate_scatter(x, e2, w2, n_dst), rel_tol=2e-6)
    # out-of-range ids are rejected loudly
    bad = dst.clone(); bad[0] = n_dst
    with pytest.raises(IndexError):
        rg.GraphHandle(row=bad, col=src, value=w, sparse_sizes=(n_dst, n_src)).to(DEV)


def test_gcn_norm_on_raw_cuda_edges_keeps_edge_order():
    u, i = O.synth_interactions(300, 200, 5000, seed=4)
    ei, ew = O.build_norm_adj(u, i, 300, 200)
    ei_d, ew_d = rg.gcn_norm(ei.to(DEV), None, 500, add_self_loops=False)
    assert torch.equal(ew_d.cpu(), ew)
    # weighted input
    w_in = torch.rand(ei.size(1), generator=torch.Generator().manual_seed(1)) + 0.5
    _, ref = O.gcn_norm(ei, w_in, 500)
    _, got = rg.gcn_norm(ei.to(DEV), w_in.to(DEV), 500, add_self_loops=False)
    assert_parity(got, ref, rel_tol=2e-6)


# ------------------------------------------------------------------------------------------ one layer
@pytest.mark.parametrize("name", ["g1", "g2"])
def test_lightgcnconv_both_edge_forms_match_golden(name, request):
    g = request.getfixturevalue(name)
    uid, iid, U, I = golden_graph(g)
    x0 = torch.cat([T(g["xu"]), T(g["xi"])]).to(DEV)
    conv = rg.LightGCNConv(x0.size(1))
    h = _handle(uid, iid, U, I)
    y_sparse = conv(x0, h, None)                                    # enable_sparse path (layers.py:19-20)
    ei, ew = T(g["edge_index"]).to(DEV), T(g["edge_weight"]).to(DEV)
    y_dense = conv(x0, ei, ew)                                      # dense-edge path (layers.py:13-17)
    for y in (y_sparse, y_dense):
        assert_parity(y, T(g["prop_dense"]), rel_tol=2e-6)
        assert_parity(y, T(g["prop_sparse"]), rel_tol=2e-6)
    assert torch.equal(y_sparse, y_dense)                           # same CSR order -> deterministic
    assert conv(x0, ei, ew) is not None and len(rg.layers._CACHE) >= 1


@pytest.mark.parametrize("D", [4, 8, 32, 48, 64, 100, 128, 200, 256, 512])
def test_spmm_all_dims_and_strides(D):
    u, i = O.synth_interactions(400, 300, 9000, seed=D)
    ei, ew = O.build_norm_adj(u, i, 400, 300)
    gen = torch.Generator().manual_seed(D)
    x = torch.rand(700, D, generator=gen) * 2 - 1                   # U(-1,1): the absolute bound bites
    ref = O.propagate_scatter(x, ei, ew)
    h = _handle(u, i, 400, 300)
    assert_parity(F_.spmm(h, x.to(DEV)), ref, rel_tol=2e-6, what=f"D={D}")
    # strided input / output views (column slices of a wider buffer, as NGCF's concat buffer)
    wide = torch.zeros(700, D + 8, device=DEV)
    wide[:, 4:4 + D] = x.to(DEV)
    out = torch.zeros(700, D + 12, device=DEV)
    F_.spmm_raw(h, wide[:, 4:4 + D], y=out[:, 8:8 + D])
    assert_parity(out[:, 8:8 + D], ref, rel_tol=2e-6)
    assert float(out[:, :8].abs().sum()) == 0.0 and float(out[:, 8 + D:].abs().sum()) == 0.0


def test_spmm_argument_errors():
    u, i = O.synth_interactions(40, 30, 200, seed=1)
    h = _handle(u, i, 40, 30)
    with pytest.raises(ValueError):
        F_.spmm(h, torch.zeros(70, 6, device=DEV))                  # dim % 4
    with pytest.raises(ValueError):
        F_.spmm(h, torch.zeros(71, 8, device=DEV))                  # row count
    with pytest.raises(TypeError):
        F_.spmm(h, torch.zeros(70, 8, device=DEV, dtype=torch.float64))
    with pytest.raises(RuntimeError):
        F_.spmm(h, torch.zeros(70, 8))                              # CPU tensor: no fallback
    with pytest.raises(ValueError):
        F_.spmm(h, torch.zeros(70, 516, device=DEV))                # dim > 512


def test_empty_and_degenerate_graphs():
    # no interactions at all
    h = _handle(torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.int64), 3, 4)
    y = F_.spmm(h, torch.ones(7, 8, device=DEV))
    assert y.shape == (7, 8) and float(y.abs().sum()) == 0.0
    # a single interaction, only [PAD]-adjacent rows otherwise
    h = _handle(torch.tensor([2]), torch.tensor([1]), 3, 2)
    x = torch.arange(5 * 4, dtype=torch.float32).view(5, 4).to(DEV)
    y = F_.spmm(h, x).cpu()
    assert torch.equal(y[2], x[4].cpu()) and torch.equal(y[4], x[2].cpu()) and float(y[[0, 1, 3]].abs().sum()) == 0.0


def test_hub_rows_power_law_graph():
    """Zipf item popularity: a few item rows hold tens of thousands of entries (> LONG_ROW) and take the
    hub kernel; results must match the oracle and be deterministic."""
    U, I, E = 30000, 2000, 400000
    u, i = O.synth_interactions(U, I, E, seed=2, zipf_alpha=1.3)
    ei, ew = O.build_norm_adj(u, i, U, I)
    h = _handle(u, i, U, I)
    assert h._n_hubs > 0, "fixture should exercise the hub path"
    x = torch.rand(U + I, 64, generator=torch.Generator().manual_seed(0)) * 2 - 1
    ref = O.propagate_f64(x, ei, ew).float()
    y1 = F_.spmm(h, x.to(DEV))
    y2 = F_.spmm(h, x.to(DEV))
    assert_parity(y1, ref, rel_tol=5e-6)
    assert torch.equal(y1, y2)


# ------------------------------------------------------------------------------------------ model loops
@pytest.mark.parametrize("name", ["g1", "g2"])
@pytest.mark.parametrize("L", [2, 3])
def test_lightgcn_forward_matches_golden(name, L, request):
    g = request.getfixturevalue(name)
    uid, iid, U, I = golden_graph(g)
    h = _handle(uid, iid, U, I)
    xu, xi = T(g["xu"]).to(DEV), T(g["xi"]).to(DEV)
    ref = T(g[f"lightgcn_L{L}"])
    u, i = F_.lightgcn_propagate(h, xu, xi, L)                      # fused
    assert_parity(torch.cat([u, i]), ref, rel_tol=2e-6)
    # reference-shaped loop over the drop-in layer
    conv = rg.LightGCNConv(xu.size(1))
    e = torch.cat([xu, xi]); embs = [e]
    for _ in range(L):
        e = conv(e, h, None); embs.append(e)
    assert_parity(torch.stack(embs, 1).mean(1), ref, rel_tol=2e-6)


def test_model_classes_on_fixture(g1):
    uid, iid, U, I = golden_graph(g1)
    ds = rg.InteractionDataset(uid, iid, U, I, device=DEV)
    for sparse in (True, None):
        cfg = {"device": DEV, "enable_sparse": sparse, "embedding_size": 64, "n_layers": 3}
        m = rg.LightGCN(cfg, ds).to(DEV)
        assert m.use_sparse == bool(sparse)
        with torch.no_grad():
            m.user_embedding.weight.copy_(T(g1["xu"])); m.item_embedding.weight.copy_(T(g1["xi"]))
        u, i = m.forward()
        assert_parity(torch.cat([u, i]), T(g1["lightgcn_L3"]), rel_tol=2e-6)
        m.fused = False
        u2, i2 = m.forward()
        assert_parity(torch.cat([u2, i2]), T(g1["lightgcn_L3"]), rel_tol=2e-6)
    with pytest.raises(ValueError):
        rg.LightGCN({"device": DEV, "enable_sparse": "yes", "embedding_size": 64, "n_layers": 3}, ds)


def test_simgcl_matches_golden(g1):
    uid, iid, U, I = golden_graph(g1)
    h = _handle(uid, iid, U, I)
    xu, xi = T(g1["xu"]).to(DEV), T(g1["xi"]).to(DEV)
    noise = (T(g1["simgcl_L3_noise_u8"]).float() / 256.0).to(DEV)
    u, i = F_.simgcl_propagate(h, xu, xi, 3, 0.1, perturbed=True, noises=[noise[l] for l in range(3)])
    assert_parity(torch.cat([u, i]), T(g1["simgcl_L3"]), rel_tol=2e-6)
    u, i = F_.simgcl_propagate(h, xu, xi, 3, 0.1, perturbed=False)
    assert_parity(torch.cat([u, i]), T(g1["simgcl_clean_L3"]), rel_tol=2e-6)


def test_simgcl_inkernel_noise_statistics(g1):
    """Fused Philox noise cannot be bit-matched to torch.rand_like; check what the perturbation must
    satisfy (simgcl.py:31-32): per row ||delta||_2 == eps, sign(delta) == sign(e), and the noise
    directions are uniform-positive (mean direction cosine with the all-ones vector as for U[0,1)^D)."""
    uid, iid, U, I = golden_graph(g1)
    h = _handle(uid, iid, U, I)
    x0 = torch.cat([T(g1["xu"]), T(g1["xi"])]).to(DEV)
    clean = torch.empty_like(x0); pert = torch.empty_like(x0); pert2 = torch.empty_like(x0)
    F_.spmm_raw(h, x0, y=clean)
    F_.spmm_raw(h, x0, y=pert, eps=0.1, seed=1234)
    F_.spmm_raw(h, x0, y=pert2, eps=0.1, seed=1235)
    delta = pert - clean
    deg = (h.csr()[0][1:] - h.csr()[0][:-1]) > 0
    nz = (clean != 0).all(dim=1) & deg
    assert nz.sum() > 1000
    assert torch.allclose(delta[nz].norm(dim=1), torch.full((int(nz.sum()),), 0.1, device=DEV), atol=2e-6)
    assert bool((torch.sign(delta[nz]) == torch.sign(clean[nz])).all())
    assert not torch.equal(pert, pert2)
    n_hat = (delta[nz].abs() / 0.1)
    cos = n_hat.sum(dim=1) / 8.0                                     # <n_hat, 1/sqrt(64)>
    assert 0.84 < float(cos.mean()) < 0.89                           # E = sqrt(3)/2 = 0.866 for U[0,1)^64
    # rows with no neighbours: sign(0) = 0 -> untouched
    assert float(pert[~deg].abs().sum()) == 0.0


def test_bignn_and_ngcf_match_golden(g1):
    uid, iid, U, I = golden_graph(g1)
    N = U + I
    h = _handle(uid, iid, U, I)
    W = [tuple(t.to(DEV) for t in w) for w in ngcf_weights(g1)]
    xu, xi = T(g1["ngcf_xu"]).to(DEV), T(g1["ngcf_xi"]).to(DEV)
    x0 = torch.cat([xu, xi])
    layer = rg.BiGNNConv(64, 64).to(DEV)
    with torch.no_grad():
        layer.lin1.weight.copy_(W[0][0]); layer.lin1.bias.copy_(W[0][1])
        layer.lin2.weight.copy_(W[0][2]); layer.lin2.bias.copy_(W[0][3])
        y_fused = layer(x0, h, None)                                # fused tail (no grad)
    y_train = layer(x0, h, None)                                    # autograd route
    assert_parity(y_fused, T(g1["bignn_layer0"]), rel_tol=5e-6)
    assert_parity(y_train, T(g1["bignn_layer0"]), rel_tol=5e-6)
    u, i = F_.ngcf_forward(h, xu, xi, W)
    assert_parity(torch.cat([u, i]), T(g1["ngcf_p0"]), rel_tol=5e-6)
    masks = [m.to(DEV) for m in ngcf_masks(g1, N, 64)]
    u, i = F_.ngcf_forward(h, xu, xi, W, message_dropout=0.1, keep_masks=masks)
    assert_parity(torch.cat([u, i])[:, -64:], T(g1["ngcf_p01"]), rel_tol=5e-6)


@pytest.mark.parametrize("n", [1, 127, 128, 333, 128 * 148 * 2 + 77])
def test_bignn_tail_tensor_core_path(n):
    """d_in = d_out = 64 runs on tcgen05 (3xTF32 split, TMEM accumulator): fp32-accurate against a float64 reference,
    every output (strided concat slice, contiguous copy, pre-activation), partial and multiple tiles per CTA."""
    gen = torch.Generator().manual_seed(n)
    d = 64
    p, x = torch.randn(n, d, generator=gen), torch.randn(n, d, generator=gen)
    w1, w2 = O.xavier_normal_((d, d), 1), O.xavier_normal_((d, d), 2)
    b1, b2 = torch.randn(d, generator=gen) * 0.1, torch.randn(d, generator=gen) * 0.1
    keep = torch.rand(n, d, generator=gen) > 0.2
    t = (torch.nn.functional.linear((p + x).double(), w1.double(), b1.double()) +
         torch.nn.functional.linear((p * x).double(), w2.double(), b2.double()))
    ref = torch.nn.functional.normalize(torch.nn.functional.leaky_relu(t, 0.2) * keep / 0.8, p=2, dim=1)
    a = [v.to(DEV) for v in (p, x, w1, b1, w2, b2)]
    cat = torch.zeros(n, 3 * d, device=DEV)
    out2, pre = torch.empty(n, d, device=DEV), torch.empty(n, d, device=DEV)
    F_.bignn_tail(*a, keep=keep.to(DEV), drop_p=0.2, out=cat[:, d:2 * d], out2=out2, pre_out=pre)
    r = assert_parity(cat[:, d:2 * d], ref.float(), rel_tol=2e-6, what="tc tail out")
    assert torch.equal(out2, cat[:, d:2 * d]) and not cat[:, :d].any() and not cat[:, 2 * d:].any()
    assert_parity(pre, t.float(), rel_tol=2e-6, what="tc tail pre-activation")
    print("tc tail", n, r)


@pytest.mark.parametrize("n,normalize,drop", [(1, True, 0.0), (333, True, 0.2), (333, False, 0.2), (40_001, True, 0.1),
                                              (128 * 148 + 5, False, 0.0)])
def test_bignn_tail_backward_tensor_core_path(n, normalize, drop):
    """The fused 64 x 64 tail backward (tcgen05 dgrad + row-local parts, library GEMM for the weight gradient) against
    float64 autograd of the reference algebra (layers.py:56-58, ngcf.py:96-98)."""
    gen = torch.Generator().manual_seed(n + int(normalize))
    d = 64
    p, x = torch.randn(n, d, generator=gen) * 0.5, torch.randn(n, d, generator=gen)
    w1, w2 = O.xavier_normal_((d, d), 1), O.xavier_normal_((d, d), 2)
    b1, b2 = torch.randn(d, generator=gen) * 0.1, torch.randn(d, generator=gen) * 0.1
    keep = (torch.rand(n, d, generator=gen) >= drop) if drop > 0 else None
    g = torch.randn(n, d, generator=gen)
    leaves = [v.double().requires_grad_(True) for v in (p, x, w1, b1, w2, b2)]
    P, X, W1, B1, W2, B2 = leaves
    t = torch.nn.functional.linear(P + X, W1, B1) + torch.nn.functional.linear(P * X, W2, B2)
    z = torch.nn.functional.leaky_relu(t, 0.2)
    if keep is not None:
        z = z * keep / (1 - drop)
    out = torch.nn.functional.normalize(z, p=2, dim=1) if normalize else z
    (out * g.double()).sum().backward()
    dl = [v.to(DEV).requires_grad_(True) for v in (p, x, w1, b1, w2, b2)]
    got = F_.bignn_tail_autograd(*dl, slope=0.2, keep=None if keep is None else keep.to(DEV), drop_p=drop,
                                 normalize=normalize)
    assert_parity(got, out.float(), rel_tol=2e-6)
    got.backward(g.to(DEV))
    for name, a, b in zip("p x w1 b1 w2 b2".split(), dl, leaves):
        assert_parity(a.grad, b.grad.float(), abs_tol=1e-3, rel_tol=5e-6, what=f"grad {name}")


@pytest.mark.parametrize("d_in,d_out", [(64, 32), (32, 128), (128, 64), (200, 256), (8, 4)])
def test_bignn_tail_shapes(d_in, d_out):
    gen = torch.Generator().manual_seed(d_in * 1000 + d_out)
    n = 333
    p, x = torch.randn(n, d_in, generator=gen), torch.randn(n, d_in, generator=gen)
    w1, w2 = O.xavier_normal_((d_out, d_in), 1), O.xavier_normal_((d_out, d_in), 2)
    b1, b2 = torch.randn(d_out, generator=gen) * 0.1, torch.randn(d_out, generator=gen) * 0.1
    keep = torch.rand(n, d_out, generator=gen) > 0.2
    t = torch.nn.functional.linear(p + x, w1, b1) + torch.nn.functional.linear(p * x, w2, b2)
    ref = torch.nn.functional.leaky_relu(t, 0.2) * keep / 0.8
    ref = torch.nn.functional.normalize(ref, p=2, dim=1)
    a = [v.to(DEV) for v in (p, x, w1, b1, w2, b2)]
    got = F_.bignn_tail(*a, keep=keep.to(DEV), drop_p=0.2)
    assert_parity(got, ref, rel_tol=1e-5, what=f"{d_in}x{d_out}")


def test_ngcf_model_class_and_node_dropout(g1):
    uid, iid, U, I = golden_graph(g1)
    ds = rg.InteractionDataset(uid, iid, U, I, device=DEV)
    cfg = {"device": DEV, "enable_sparse": True, "embedding_size": 64, "hidden_size_list": [64, 64, 64],
           "node_dropout": 0.0, "message_dropout": 0.0}
    m = rg.NGCF(cfg, ds).to(DEV)
    W = ngcf_weights(g1)
    with torch.no_grad():
        m.user_embedding.weight.copy_(T(g1["ngcf_xu"])); m.item_embedding.weight.copy_(T(g1["ngcf_xi"]))
        for l, layer in enumerate(m.GNNlayers):
            layer.lin1.weight.copy_(W[l][0]); layer.lin1.bias.copy_(W[l][1])
            layer.lin2.weight.copy_(W[l][2]); layer.lin2.bias.copy_(W[l][3])
        u, i = m.forward()
    assert_parity(torch.cat([u, i]), T(g1["ngcf_p0"]), rel_tol=5e-6)
    u, i = m.forward()                                               # training route, same numbers
    assert_parity(torch.cat([u, i]), T(g1["ngcf_p0"]), rel_tol=5e-6)
    m.node_dropout = 0.3; m.train()
    u, i = m.forward()
    assert torch.isfinite(u).all() and u.shape == (U, 256)


# ------------------------------------------------------------------------------------------ autograd
def test_backward_matches_oracle_autograd(g1):
    uid, iid, U, I = golden_graph(g1)
    h = _handle(uid, iid, U, I)
    ei, ew = O.build_norm_adj(uid, iid, U, I)
    xu_c = T(g1["xu"]).clone().requires_grad_(True); xi_c = T(g1["xi"]).clone().requires_grad_(True)
    gen = torch.Generator().manual_seed(3)
    gu, gi = torch.randn(U, 64, generator=gen), torch.randn(I, 64, generator=gen)
    u, i = O.lightgcn_forward(xu_c, xi_c, ei, ew, 3)
    (u * gu).sum().add((i * gi).sum()).backward()
    xu_d = T(g1["xu"]).to(DEV).requires_grad_(True); xi_d = T(g1["xi"]).to(DEV).requires_grad_(True)
    u2, i2 = F_.lightgcn_propagate(h, xu_d, xi_d, 3)
    ((u2 * gu.to(DEV)).sum() + (i2 * gi.to(DEV)).sum()).backward()
    assert_parity(xu_d.grad, xu_c.grad, rel_tol=5e-6)
    assert_parity(xi_d.grad, xi_c.grad, rel_tol=5e-6)
    # per-layer op
    x = torch.cat([T(g1["xu"]), T(g1["xi"])]).to(DEV).requires_grad_(True)
    y = rg.LightGCNConv(64)(x, h, None)
    gy = torch.randn(U + I, 64, generator=gen)
    y.backward(gy.to(DEV))
    assert_parity(x.grad, O.propagate_scatter(gy, ei.flip([0]), ew), rel_tol=5e-6)


def test_training_step_runs(g1):
    uid, iid, U, I = golden_graph(g1)
    ds = rg.InteractionDataset(uid, iid, U, I, device=DEV)
    m = rg.LightGCN({"device": DEV, "enable_sparse": True, "embedding_size": 64, "n_layers": 2}, ds).to(DEV)
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    inter = {"user_id": uid[:512].to(DEV), "item_id": iid[:512].to(DEV),
             "neg_item_id": torch.randint(1, I, (512,), device=DEV)}
    losses = []
    for _ in range(5):
        opt.zero_grad(); loss = m.calculate_loss(inter); loss.backward(); opt.step(); losses.append(float(loss))
    assert losses[-1] < losses[0]


# ------------------------------------------------------------------------------------------ larger / properties
def test_g3_synthetic_uniform_and_zipf():
    for alpha in (None, 1.1):
        U = I = 10000
        u, i = O.synth_interactions(U, I, 1_000_000, seed=1, zipf_alpha=alpha)
        ei, ew = O.build_norm_adj(u, i, U, I)
        a = O.adj_sparse(ei, ew, U + I, U + I, "csr")
        xu, xi = O.xavier_uniform_table(U, 64, 2), O.xavier_uniform_table(I, 64, 3)
        u_ref, i_ref = O.lightgcn_forward(xu, xi, ei, ew, 3, prop=lambda x: O.propagate_sparse(a, x))
        h = _handle(u, i, U, I)
        ug, ig = F_.lightgcn_propagate(h, xu.to(DEV), xi.to(DEV), 3)
        r = assert_parity(torch.cat([ug, ig]), torch.cat([u_ref, i_ref]), rel_tol=5e-6, what=f"alpha={alpha}")
        # U(-1,1) inputs so that the absolute bound bites.  Hub rows of the Zipf graph sum ~1e5 terms: the
        # fp32 oracle itself carries ~sqrt(k)*eps error there, so the float64 form is the tie-breaker
        #  and the fp32 oracle is held to the same bound.
        x = torch.rand(U + I, 64, generator=torch.Generator().manual_seed(9)) * 2 - 1
        ref64 = O.propagate_f64(x, ei, ew)
        got = F_.spmm(h, x.to(DEV))
        assert_parity(got, ref64.float())                      # 1e-4 abs, 1e-5 scaled (hub rows: ~6e-6)
        assert_parity(O.propagate_sparse(a, x), ref64.float())   # the fp32 oracle under the same bound


def test_full_size_properties_sampled_rows():
    """BASELINE config-2 scale is checked by size-independent properties: (1) rows sampled from the device
    CSR are recomputed in float64 on the host; (2) linearity A(ax+by) = aAx + bAy; (3) A·(D^1/2 1) = D^1/2 1
    on non-isolated nodes (row sums of the symmetric normalisation).  Scaled to 200k x 200k / 20 M edges to
    keep the test in seconds; bench.py runs the same sampled-row check at the full 2 M / 200 M size."""
    U = I = 200_000
    E = 20_000_000
    gen = torch.Generator(device=DEV).manual_seed(0)
    uid = torch.randint(1, U, (E,), generator=gen, device=DEV)
    iid = torch.randint(1, I, (E,), generator=gen, device=DEV)
    h = rg.GraphHandle.from_interactions(uid, iid, U, I).gcn_norm().to(DEV)
    N, D = U + I, 64
    rowptr, col, val = h.csr()
    x = torch.rand(N, D, generator=gen, device=DEV) * 2 - 1
    y = F_.spmm(h, x)
    rows = torch.randint(0, N, (512,), generator=torch.Generator().manual_seed(1))
    rp = rowptr.cpu()
    for r in rows.tolist():
        b, e = int(rp[r]), int(rp[r + 1])
        c, v = col[b:e].long(), val[b:e].double()
        ref = (v[:, None] * x[c].double()).sum(0)
        assert float((y[r].double() - ref).abs().max()) < 1e-5
    x2 = torch.rand(N, D, generator=gen, device=DEV)
    lhs = F_.spmm(h, 0.5 * x - 2.0 * x2)
    rhs = 0.5 * y - 2.0 * F_.spmm(h, x2)
    assert float((lhs - rhs).abs().max()) < 1e-5
    deg = (rowptr[1:] - rowptr[:-1]).float()
    s = deg.sqrt()[:, None].expand(N, 4).contiguous()
    ys = F_.spmm(h, s)
    assert float((ys - s).abs().max() / s.max()) < 1e-5


def test_host_propagator_pipeline_matches(g1):
    """Host-buffer entry point: pinned tables in, pinned result out, steps overlapped on three streams; every
    step must return exactly what the device-resident call returns."""
    from recbole_gnn_b200.host import HostPropagator
    uid, iid, U, I = golden_graph(g1)
    h = _handle(uid, iid, U, I)
    hp = HostPropagator(h, U, I, 64, 3, depth=2)
    outs = []
    for k in range(5):
        hu = (T(g1["xu"]) * (k + 1)).pin_memory()
        hi = (T(g1["xi"]) * (k + 1)).pin_memory()
        ou, oi = torch.empty(U, 64).pin_memory(), torch.empty(I, 64).pin_memory()
        hp.submit(hu, hi, ou, oi)
        outs.append((ou, oi, k + 1))
    hp.synchronize()
    ref = T(g1["lightgcn_L3"])
    for ou, oi, s in outs:
        assert_parity(torch.cat([ou, oi]), ref * s, abs_tol=1e-4 * s, rel_tol=2e-6)
    with pytest.raises(ValueError):
        hp.submit(T(g1["xu"]).to(DEV), T(g1["xi"]), outs[0][0], outs[0][1])


def test_identity_mode_and_simgcl_views(g1):
    uid, iid, U, I = golden_graph(g1)
    N = U + I
    h = _handle(uid, iid, U, I)
    xu, xi = T(g1["xu"]).to(DEV), T(g1["xi"]).to(DEV)
    x0 = torch.cat([xu, xi])
    y = torch.empty_like(x0)
    F_.spmm_raw(None, xu, x2=xi, y=y)                       # identity mode: p = x
    assert torch.equal(y, x0)
    noise = (T(g1["simgcl_L3_noise_u8"]).float() / 256.0)
    n1 = [noise[l].to(DEV) for l in range(3)]
    gen = torch.Generator().manual_seed(77)
    n2c = [torch.rand(N, 64, generator=gen) for _ in range(3)]
    n2 = [t.to(DEV) for t in n2c]
    (u0, i0), (u1, i1), (u2, i2) = F_.simgcl_views(h, xu, xi, 3, 0.1, noises1=n1, noises2=n2)
    assert_parity(torch.cat([u0, i0]), T(g1["simgcl_clean_L3"]), rel_tol=2e-6)
    assert_parity(torch.cat([u1, i1]), T(g1["simgcl_L3"]), rel_tol=2e-6)
    ei, ew = O.build_norm_adj(uid, iid, U, I)
    ur, ir = O.simgcl_forward(T(g1["xu"]), T(g1["xi"]), ei, ew, 3, 0.1, n2c)
    assert_parity(torch.cat([u2, i2]), torch.cat([ur, ir]), rel_tol=2e-6)
    # the same three results as three separate forwards
    ua, ia = F_.simgcl_propagate(h, xu, xi, 3, 0.1, perturbed=True, noises=n2)
    assert_parity(torch.cat([u2, i2]), torch.cat([ua, ia]), rel_tol=1e-6)
    # backward: one propagation of the summed grads == autograd through the three oracle forwards
    xu_d, xi_d = xu.clone().requires_grad_(True), xi.clone().requires_grad_(True)
    views = F_.simgcl_views(h, xu_d, xi_d, 3, 0.1, noises1=n1, noises2=n2)
    gs = [torch.randn(U, 64, generator=gen) for _ in range(3)] + [torch.randn(I, 64, generator=gen) for _ in range(3)]
    loss = sum((views[v][0] * gs[v].to(DEV)).sum() + (views[v][1] * gs[3 + v].to(DEV)).sum() for v in range(3))
    loss.backward()
    xu_c, xi_c = T(g1["xu"]).clone().requires_grad_(True), T(g1["xi"]).clone().requires_grad_(True)
    lc = 0
    for v, nz in enumerate((None, [noise[l] for l in range(3)], n2c)):
        uc, ic = O.simgcl_forward(xu_c, xi_c, ei, ew, 3, 0.1, nz)
        lc = lc + (uc * gs[v]).sum() + (ic * gs[3 + v]).sum()
    lc.backward()
    assert_parity(xu_d.grad, xu_c.grad, rel_tol=5e-6)
    assert_parity(xi_d.grad, xi_c.grad, rel_tol=5e-6)
    # L = 1 and L = 2 shapes
    for L in (1, 2):
        v = F_.simgcl_views(h, xu, xi, L, 0.1, noises1=n1[:L], noises2=n2[:L])
        ur, ir = O.simgcl_forward(T(g1["xu"]), T(g1["xi"]), ei, ew, L, 0.1, n2c[:L])
        assert_parity(torch.cat(v[2]), torch.cat([ur, ir]), rel_tol=2e-6, what=f"L={L}")
        ur, ir = O.simgcl_forward(T(g1["xu"]), T(g1["xi"]), ei, ew, L, 0.1, None)
        assert_parity(torch.cat(v[0]), torch.cat([ur, ir]), rel_tol=2e-6, what=f"L={L} clean")


@pytest.mark.parametrize("L", [1, 4, 5, 6])
def test_layer_combine_paths(L, g1):
    """L <= 4 (LightGCN) / L <= 5 (SimGCL) form the mean in the last layer's epilogue from the stored layer outputs;
    deeper stacks fall back to the running sum.  Both against the oracle."""
    uid, iid, U, I = golden_graph(g1)
    h = _handle(uid, iid, U, I)
    ei, ew = O.build_norm_adj(uid, iid, U, I)
    xu, xi = T(g1["xu"]), T(g1["xi"])
    u_ref, i_ref = O.lightgcn_forward(xu, xi, ei, ew, L)
    u, i = F_.lightgcn_propagate(h, xu.to(DEV), xi.to(DEV), L)
    assert_parity(torch.cat([u, i]), torch.cat([u_ref, i_ref]), rel_tol=2e-6, what=f"lightgcn L={L}")
    u_ref, i_ref = O.simgcl_forward(xu, xi, ei, ew, L, 0.1, None)
    u, i = F_.simgcl_propagate(h, xu.to(DEV), xi.to(DEV), L, 0.1, perturbed=False)
    assert_parity(torch.cat([u, i]), torch.cat([u_ref, i_ref]), rel_tol=2e-6, what=f"simgcl L={L}")


def test_ngcf_training_route_grads_match_oracle(g1):
    """Training route of NGCF: SpMM (own backward) + fused tail with the hand-written backward, against torch
    autograd through the CPU oracle, with a recorded dropout mask."""
    uid, iid, U, I = golden_graph(g1)
    N = U + I
    ds = rg.InteractionDataset(uid, iid, U, I, device=DEV)
    cfg = {"device": DEV, "enable_sparse": True, "embedding_size": 64, "hidden_size_list": [64, 64, 64],
           "node_dropout": 0.0, "message_dropout": 0.1}
    m = rg.NGCF(cfg, ds).to(DEV)
    W = ngcf_weights(g1)
    masks = ngcf_masks(g1, N, 64)
    with torch.no_grad():
        m.user_embedding.weight.copy_(T(g1["ngcf_xu"])); m.item_embedding.weight.copy_(T(g1["ngcf_xi"]))
        for l, layer in enumerate(m.GNNlayers):
            layer.lin1.weight.copy_(W[l][0]); layer.lin1.bias.copy_(W[l][1])
            layer.lin2.weight.copy_(W[l][2]); layer.lin2.bias.copy_(W[l][3])
    u, i = m.forward(keep_masks=[k.to(DEV) for k in masks])
    assert_parity(torch.cat([u, i])[:, -64:], T(g1["ngcf_p01"]), rel_tol=5e-6)
    gen = torch.Generator().manual_seed(4)
    gu, gi = torch.randn(U, 256, generator=gen), torch.randn(I, 256, generator=gen)
    ((u * gu.to(DEV)).sum() + (i * gi.to(DEV)).sum()).backward()
    # oracle with autograd
    ei, ew = O.build_norm_adj(uid, iid, U, I)
    xu = T(g1["ngcf_xu"]).clone().requires_grad_(True); xi = T(g1["ngcf_xi"]).clone().requires_grad_(True)
    Wc = [tuple(t.clone().requires_grad_(True) for t in w) for w in W]
    uo, io = O.ngcf_forward(xu, xi, ei, ew, Wc, message_dropout=0.1, drop_masks=masks)
    ((uo * gu).sum() + (io * gi).sum()).backward()
    assert_parity(m.user_embedding.weight.grad, xu.grad, abs_tol=1e-3, rel_tol=2e-5, what="grad xu")
    assert_parity(m.item_embedding.weight.grad, xi.grad, abs_tol=1e-3, rel_tol=2e-5, what="grad xi")
    for l, layer in enumerate(m.GNNlayers):
        for name, got, ref in (("w1", layer.lin1.weight.grad, Wc[l][0].grad), ("b1", layer.lin1.bias.grad, Wc[l][1].grad),
                               ("w2", layer.lin2.weight.grad, Wc[l][2].grad), ("b2", layer.lin2.bias.grad, Wc[l][3].grad)):
            assert_parity(got, ref, abs_tol=1e-2, rel_tol=5e-5, what=f"layer {l} {name}")
    # BiGNNConv alone (pre-activation output) with autograd
    layer = m.GNNlayers[0]
    x = torch.cat([T(g1["ngcf_xu"]), T(g1["ngcf_xi"])]).to(DEV).requires_grad_(True)
    y = layer(x, m.edge_index, None)
    assert_parity(y, T(g1["bignn_layer0"]), rel_tol=5e-6)
    y.sum().backward()
    assert torch.isfinite(x.grad).all()
