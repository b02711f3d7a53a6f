"""CPU: the oracle restatement against the golden vectors produced by executing the reference's own
layers.py / dataset.py (tests/golden/make_golden.py), plus hand-checked values on the toy graph."""
import math

import pytest
import torch

from oracle import oracle as O
from tests.helpers import T, assert_parity, golden_graph, ngcf_masks, ngcf_weights

TIGHT = dict(abs_tol=2e-6, rel_tol=2e-6)


@pytest.mark.parametrize("name", ["g1", "g2"])
def test_norm_adj_matches_reference(name, request):
    g = request.getfixturevalue(name)
    uid, iid, U, I = golden_graph(g)
    ei, ew = O.build_norm_adj(uid, iid, U, I)
    assert torch.equal(ei, T(g["edge_index"]))
    assert torch.equal(ew, T(g["edge_weight"]))          # bit-exact: same formula, same order
    # sparse path object: same entries sorted by (row, col)
    a = O.adj_sparse(ei, ew, U + I, U + I, "coo")
    # duplicates are coalesced by torch.sparse; compare through a product instead of entries
    x = torch.randn(U + I, 8, generator=torch.Generator().manual_seed(1))
    ref = torch.zeros(U + I, 8).index_add_(0, T(g["adj_row"]), T(g["adj_val"])[:, None] * x[T(g["adj_col"])])
    assert_parity(O.propagate_sparse(a, x), ref, **TIGHT)


@pytest.mark.parametrize("name", ["g1", "g2"])
def test_propagation_forms_agree_with_reference(name, request):
    g = request.getfixturevalue(name)
    uid, iid, U, I = golden_graph(g)
    ei, ew = O.build_norm_adj(uid, iid, U, I)
    x0 = torch.cat([T(g["xu"]), T(g["xi"])], 0)
    y_ref_dense, y_ref_sparse = T(g["prop_dense"]), T(g["prop_sparse"])
    assert torch.equal(O.propagate_scatter(x0, ei, ew), y_ref_dense)      # same op order as layers.py:13-17
    for layout in ("coo", "csr"):
        y = O.propagate_sparse(O.adj_sparse(ei, ew, U + I, U + I, layout), x0)
        assert_parity(y, y_ref_sparse, **TIGHT, what=layout)
        assert_parity(y, y_ref_dense, **TIGHT, what=layout)
    assert_parity(O.propagate_f64(x0, ei, ew).float(), y_ref_dense, **TIGHT)


@pytest.mark.parametrize("name", ["g1", "g2"])
def test_model_loops_match_reference(name, request):
    g = request.getfixturevalue(name)
    uid, iid, U, I = golden_graph(g)
    N = U + I
    ei, ew = O.build_norm_adj(uid, iid, U, I)
    xu, xi = T(g["xu"]), T(g["xi"])
    for L in (2, 3):
        u, i = O.lightgcn_forward(xu, xi, ei, ew, L)
        assert torch.equal(torch.cat([u, i]), T(g[f"lightgcn_L{L}"]))
    noise = T(g["simgcl_L3_noise_u8"]).float() / 256.0
    u, i = O.simgcl_forward(xu, xi, ei, ew, 3, 0.1, [noise[l] for l in range(3)])
    assert torch.equal(torch.cat([u, i]), T(g["simgcl_L3"]))
    u, i = O.simgcl_forward(xu, xi, ei, ew, 3, 0.1, None)
    assert torch.equal(torch.cat([u, i]), T(g["simgcl_clean_L3"]))
    # NGCF
    W = ngcf_weights(g)
    xun, xin = T(g["ngcf_xu"]), T(g["ngcf_xi"])
    x0n = torch.cat([xun, xin])
    assert_parity(O.bignn_layer(x0n, ei, ew, *W[0]), T(g["bignn_layer0"]), **TIGHT)
    u, i = O.ngcf_forward(xun, xin, ei, ew, W)
    assert_parity(torch.cat([u, i]), T(g["ngcf_p0"]), **TIGHT)
    D = xun.size(1)
    masks = ngcf_masks(g, N, D)
    u, i = O.ngcf_forward(xun, xin, ei, ew, W, message_dropout=0.1, drop_masks=masks)
    assert_parity(torch.cat([u, i])[:, -D:], T(g["ngcf_p01"]), **TIGHT)


@pytest.mark.parametrize("name", ["g1", "g2"])
def test_losses_match_reference_calculate_loss(name, request):
    """The oracle's loss restatements vs LightGCN/SimGCL/NGCF.calculate_loss executed from the reference's own
    files (values and embedding / weight gradients)."""
    g = request.getfixturevalue(name)
    uid, iid, U, I = golden_graph(g)
    ei, ew = O.build_norm_adj(uid, iid, U, I)
    user, pos, neg = T(g["batch_user"]), T(g["batch_pos"]), T(g["batch_neg"])
    for rp, tag in ((False, "nopow"), (True, "pow")):
        xu, xi = T(g["xu"]).requires_grad_(), T(g["xi"]).requires_grad_()
        loss = O.lightgcn_loss(xu, xi, ei, ew, 3, user, pos, neg, 1e-5, rp)
        loss.backward()
        assert_parity(loss.detach().reshape(1), T(g[f"lightgcn_loss_{tag}"]), abs_tol=1e-6, rel_tol=1e-6)
        assert_parity(xu.grad, T(g[f"lightgcn_loss_{tag}_gu"]), abs_tol=1e-7, rel_tol=1e-5)
        assert_parity(xi.grad, T(g[f"lightgcn_loss_{tag}_gi"]), abs_tol=1e-7, rel_tol=1e-5)
    # SimGCL: the stored lightgcn_L*_sparse come from the enable_sparse=True graph object
    for L in (2, 3):
        assert_parity(T(g[f"lightgcn_L{L}_sparse"]), T(g[f"lightgcn_L{L}"]), **TIGHT)
    nz = T(g["simgcl_loss_noise_u8"]).float() / 256.0
    xu, xi = T(g["xu"]).requires_grad_(), T(g["xi"]).requires_grad_()
    loss = O.simgcl_loss(xu, xi, ei, ew, 3, 0.1, [nz[l] for l in range(3)], [nz[l] for l in range(3, 6)],
                         user, pos, neg)
    loss.backward()
    assert_parity(loss.detach().reshape(1), T(g["simgcl_loss"]), abs_tol=1e-4, rel_tol=2e-6)
    assert_parity(xu.grad, T(g["simgcl_loss_gu"]), abs_tol=1e-4, rel_tol=1e-5)
    assert_parity(xi.grad, T(g["simgcl_loss_gi"]), abs_tol=1e-4, rel_tol=1e-5)
    # NGCF
    W = [tuple(t.clone().requires_grad_() for t in w) for w in ngcf_weights(g)]
    xun, xin = T(g["ngcf_xu"]).requires_grad_(), T(g["ngcf_xi"]).requires_grad_()
    loss = O.ngcf_loss(xun, xin, ei, ew, W, user, pos, neg)
    loss.backward()
    assert_parity(loss.detach().reshape(1), T(g["ngcf_loss_p0"]), abs_tol=1e-6, rel_tol=2e-6)
    assert_parity(xun.grad, T(g["ngcf_loss_p0_gu"]), abs_tol=1e-6, rel_tol=1e-4)
    assert_parity(W[0][0].grad, T(g["ngcf_loss_p0_gw1_0"]), abs_tol=1e-6, rel_tol=1e-4)
    assert_parity(W[2][2].grad, T(g["ngcf_loss_p0_gw2_2"]), abs_tol=1e-6, rel_tol=1e-4)
    assert_parity(W[1][1].grad, T(g["ngcf_loss_p0_gb1_1"]), abs_tol=1e-6, rel_tol=1e-4)


@pytest.mark.parametrize("name", ["g1", "g2"])
def test_bipartite_matches_reference(name, request):
    g = request.getfixturevalue(name)
    uid, iid, U, I = golden_graph(g)
    xu, xi = T(g["xu"]), T(g["xi"])
    for row in ("user", "item"):
        for rn in (True, False):
            tag = f"bip_{row}_{'rown' if rn else 'sym'}"
            r, c = (uid, iid) if row == "user" else (iid, uid)
            n_row, n_col = (U, I) if row == "user" else (I, U)
            ei, ew = O.build_bipartite_inter_mat(r, c, n_row, n_col, rn)
            assert torch.equal(ei, T(g[tag + "_ei"]))
            assert torch.equal(ew, T(g[tag + "_ew"]))
            x_col = xi if row == "user" else xu
            y = O.bipartite_forward(x_col, ei.flip([0]), ew, n_row)
            assert torch.equal(y, T(g[tag + "_y"]))


def test_toy_graph_by_hand(g2):
    """G2: users {1,2,3}, items {1,2}; interactions (1,1) x2, (1,2), (2,1), (3,2).  N = 8, item ids offset 4."""
    uid, iid, U, I = golden_graph(g2)
    ei, ew = O.build_norm_adj(uid, iid, U, I)
    deg = {1: 3, 2: 1, 3: 1, 5: 3, 6: 2}           # node -> degree (parallel edges count twice)
    for k in range(ei.size(1)):
        s, d = int(ei[0, k]), int(ei[1, k])
        assert math.isclose(float(ew[k]), 1 / math.sqrt(deg[s] * deg[d]), rel_tol=1e-6)
    x = torch.eye(8)
    y = O.propagate_scatter(x, ei, ew)              # y = A_hat itself
    assert math.isclose(float(y[1, 5]), 2 / 3, rel_tol=1e-6)      # duplicate edge accumulates
    assert torch.all(y[0] == 0) and torch.all(y[4] == 0) and torch.all(y[7] == 0)   # PADs + isolated item
    assert torch.allclose(y, y.t())


def test_dropout_adj_and_generators():
    u, i = O.synth_interactions(50, 60, 500, seed=3)
    assert u.min() >= 1 and i.min() >= 1 and u.max() < 50 and i.max() < 60
    u2, i2 = O.synth_interactions(50, 60, 500, seed=3, zipf_alpha=1.1)
    assert i2.min() >= 1 and i2.max() < 60
    ei, ew = O.build_norm_adj(u, i, 50, 60)
    keep = torch.rand(ei.size(1), generator=torch.Generator().manual_seed(0)) > 0.3
    e2, w2 = O.dropout_adj(ei, ew, keep)
    assert e2.size(1) == int(keep.sum()) and torch.equal(w2, ew[keep])


def test_sgl_and_sept_resampling_restatements(g2):
    """sgl.py:92-126 / sept.py:81-87,111-133 restated: ND == ED over the interactions that survive the dropped nodes;
    SEPT weights by hand on the toy graph."""
    uid, iid, U, I = golden_graph(g2)
    du, di = torch.tensor([1]), torch.tensor([2])
    ei_nd, ew_nd = O.sgl_augmented_adj(uid, iid, U, I, "ND", drop_user=du, drop_item=di)
    keep = torch.nonzero(~((uid == 1) | (iid == 2))).flatten()
    ei_ed, ew_ed = O.sgl_augmented_adj(uid, iid, U, I, "ED", keep_idx=keep)
    assert torch.equal(ei_nd, ei_ed) and torch.equal(ew_nd, ew_ed)
    assert not ((ei_nd == 1) | (ei_nd == U + 2)).any()
    ei = torch.tensor([[0, 0, 1, 3], [1, 2, 0, 0]])
    w = O.sept_norm_edge_weight(ei, 4)            # out-degrees 2, 1, 0 (read as 1), 1
    assert torch.allclose(w, torch.tensor([2 ** -0.5, 2 ** -0.5, 2 ** -0.5, 2 ** -0.5]))
    e2, w2 = O.sept_subgraph(uid, iid, torch.tensor([1, 2]), torch.tensor([2, 3]), U, I,
                             torch.arange(uid.numel()), torch.tensor([1]))
    assert e2.shape == (2, 2 * uid.numel() + 1) and e2[:, -1].tolist() == [2, 3] and w2.numel() == e2.size(1)
