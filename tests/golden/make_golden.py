"""Generate golden vectors by EXECUTING the reference's own source files in this container.

Run (only where /root/reference exists; the GPU box never runs this):

    python tests/golden/make_golden.py

The reference's hot-path files import three packages that are absent here (no network):
``torch_geometric``, ``torch_sparse`` and ``recbole``.  This script installs minimal stand-ins for
exactly the names those two files import, then imports

    /root/reference/recbole_gnn/model/layers.py   (LightGCNConv, BipartiteGCNConv, BiGNNConv)
    /root/reference/recbole_gnn/data/dataset.py   (GeneralGraphDataset.get_norm_adj_mat,
                                                   edge_index_to_adj_t, get_bipartite_inter_mat)

unmodified, and runs them.  The stand-ins restate the third-party semantics the reference relies on:

* ``MessagePassing.propagate`` (aggr='add', flow source_to_target): Tensor edge_index ->
  ``x_j = x_src[edge_index[0]]``, ``self.message(x_j=..., edge_weight=...)``, sum-scatter onto
  ``edge_index[1]`` with ``dim_size = size[1]``; SparseTensor -> ``self.message_and_aggregate(adj_t, x)``.
* ``torch_sparse.SparseTensor(row, col, value, sparse_sizes)`` / ``.t()`` / ``.coo()`` and
  ``matmul(adj_t, x, reduce='add')`` = CSR SpMM with per-row sequential accumulation.
* ``gcn_norm(·, add_self_loops=False)`` and ``degree`` as documented by PyG.

So the *reference-owned* lines (message(), forward() bodies, the COO assembly and the
normalisation formulas of get_bipartite_inter_mat) are the real thing; the third-party kernels are
restated.  The model-level loops (lightgcn.py:70-81, ngcf.py:92-104, simgcl.py:24-38) import recbole
model base classes and losses and are re-driven here from the reference conv layers directly.

Outputs: tests/golden/g1_fixture.npz (graph of tests/test_data/test/test.inter, D=64) and
tests/golden/g2_toy.npz (hand-checkable toy graph with a duplicate edge and an isolated node).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


# ----------------------------------------------------------------------------- stand-ins
class SparseTensor:
    def __init__(self, row, col, value=None, sparse_sizes=None):
        self.row, self.col, self.value, self.sizes = row, col, value, tuple(sparse_sizes)

    def t(self):
        return SparseTensor(self.col, self.row, self.value, (self.sizes[1], self.sizes[0]))

    def coo(self):
        # torch_sparse keeps entries sorted by (row, col)
        key = self.row * self.sizes[1] + self.col
        perm = torch.argsort(key, stable=True)
        v = None if self.value is None else self.value[perm]
        return self.row[perm], self.col[perm], v

    def to(self, device):
        return self


def sparse_matmul(adj_t, x, reduce="add"):
    assert reduce == "add"
    row, col, val = adj_t.coo()
    a = torch.sparse_coo_tensor(torch.stack([row, col]), val, adj_t.sizes).coalesce().to_sparse_csr()
    return a @ x


class MessagePassing(nn.Module):
    def __init__(self, aggr="add", **kw):
        super().__init__()
        self.aggr = aggr

    def propagate(self, edge_index, size=None, **kwargs):
        x = kwargs["x"]
        if isinstance(edge_index, SparseTensor):
            return self.message_and_aggregate(edge_index, x)
        x_src = x[0] if isinstance(x, (tuple, list)) else x
        dim_size = size[1] if size is not None else x_src.size(0)
        x_j = x_src.index_select(0, edge_index[0])
        msg = self.message(x_j=x_j, edge_weight=kwargs["edge_weight"])
        out = torch.zeros(dim_size, msg.size(1), dtype=msg.dtype)
        return out.index_add_(0, edge_index[1], msg)


def degree(index, num_nodes=None, dtype=None):
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    return torch.zeros(n, dtype=torch.float32).scatter_add_(0, index, torch.ones(index.numel()))


def gcn_norm(edge_index, edge_weight=None, num_nodes=None, improved=False, add_self_loops=True,
             flow="source_to_target", dtype=None):
    assert add_self_loops is False and flow == "source_to_target"
    if isinstance(edge_index, SparseTensor):
        adj_t = edge_index
        val = adj_t.value if adj_t.value is not None else torch.ones(adj_t.row.numel())
        deg = torch.zeros(adj_t.sizes[0]).scatter_add_(0, adj_t.row, val)      # sum(adj_t, dim=1)
        dis = deg.pow(-0.5)
        dis.masked_fill_(dis == float("inf"), 0.0)
        val = val * dis[adj_t.row]
        val = val * dis[adj_t.col]
        return SparseTensor(adj_t.row, adj_t.col, val, adj_t.sizes)
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1))
    row, col = edge_index[0], edge_index[1]
    deg = torch.zeros(num_nodes, dtype=edge_weight.dtype).scatter_add_(0, col, edge_weight)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    return edge_index, dis[row] * edge_weight * dis[col]


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_standins():
    _module("torch_geometric")
    _module("torch_geometric.nn", MessagePassing=MessagePassing)
    _module("torch_geometric.nn.conv")
    _module("torch_geometric.nn.conv.gcn_conv", gcn_norm=gcn_norm)
    _module("torch_geometric.utils", degree=degree)
    _module("torch_sparse", SparseTensor=SparseTensor, matmul=sparse_matmul)

    class _Dataset:
        def __init__(self, config=None):
            pass

    rb = _module("recbole", __version__="1.1.1")
    _module("recbole.data")
    _module("recbole.data.dataset", SequentialDataset=_Dataset, Dataset=_Dataset)
    _module("recbole.utils", set_color=lambda s, c: s, FeatureSource=object, ensure_dir=lambda d: None)
    return rb


def load_ref(relpath, modname):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_dataset(ds_cls, uid, iid, user_num, item_num):
    """An instance of the reference's GeneralGraphDataset carrying only the attributes that
    get_norm_adj_mat / get_bipartite_inter_mat read (RecBole's loader is stood in for)."""

    class FixtureDataset(ds_cls):
        def __init__(self):
            self.inter_feat = {"user_id": uid, "item_id": iid}
            self.uid_field, self.iid_field = "user_id", "item_id"
            self.user_num, self.item_num = user_num, item_num

        def num(self, field):
            return self.user_num if field == self.uid_field else self.item_num

    return FixtureDataset()


def model_loops(layers, ds_cls, uid, iid, U, I, D, seed):
    """Drive the reference conv layers through the three model loops."""
    from oracle import oracle as O

    out = {}
    ds = make_dataset(ds_cls, uid, iid, U, I)
    edge_index, edge_weight = ds.get_norm_adj_mat(enable_sparse=False)
    adj_t, none = ds.get_norm_adj_mat(enable_sparse=True)
    assert none is None
    out["edge_index"], out["edge_weight"] = edge_index, edge_weight
    r, c, v = adj_t.coo()
    out["adj_row"], out["adj_col"], out["adj_val"] = r, c, v

    N = U + I
    xu, xi = O.xavier_uniform_table(U, D, seed), O.xavier_uniform_table(I, D, seed + 1)
    x0 = torch.cat([xu, xi], 0)
    out["xu"], out["xi"] = xu, xi

    conv = layers.LightGCNConv(D)
    out["prop_dense"] = conv(x0, edge_index, edge_weight)            # layers.py:13-17
    out["prop_sparse"] = conv(x0, adj_t, None)                       # layers.py:19-20
    for L in (2, 3):                                                 # lightgcn.py:70-81
        embs, e = [x0], x0
        for _ in range(L):
            e = conv(e, edge_index, edge_weight)
            embs.append(e)
        out[f"lightgcn_L{L}"] = torch.stack(embs, dim=1).mean(dim=1)

    # SimGCL perturbed forward with recorded noise (simgcl.py:24-38), L=3.  rand_like draws U[0,1);
    # the recorded draw is quantised to k/256 so that it stores as uint8 (still a valid U[0,1) sample).
    g = torch.Generator().manual_seed(seed + 7)
    L = 3
    noise_u8 = torch.randint(0, 256, (L, N, D), generator=g, dtype=torch.uint8)
    noises = [noise_u8[l].float() / 256.0 for l in range(L)]
    embs, e = [], x0
    for l in range(L):
        e = conv(e, edge_index, edge_weight)
        e = e + torch.sign(e) * F.normalize(noises[l], dim=-1) * 0.1
        embs.append(e)
    out["simgcl_L3"] = torch.stack(embs, dim=1).mean(dim=1)
    out["simgcl_L3_noise_u8"] = noise_u8
    embs, e = [], x0
    for l in range(3):
        e = conv(e, edge_index, edge_weight)
        embs.append(e)
    out["simgcl_clean_L3"] = torch.stack(embs, dim=1).mean(dim=1)

    # NGCF 3 layers, hidden [64,64,64] (NGCF.yaml), xavier_normal weights, zero biases, dropout 0 and
    # dropout 0.1 with a recorded mask (ngcf.py:92-102)
    xun = O.xavier_normal_((U, D), seed + 2)
    xin = O.xavier_normal_((I, D), seed + 3)
    x0n = torch.cat([xun, xin], 0)
    out["ngcf_xu"], out["ngcf_xi"] = xun, xin
    gnn = []
    for l in range(3):
        m = layers.BiGNNConv(D, D)
        with torch.no_grad():
            m.lin1.weight.copy_(O.xavier_normal_((D, D), seed + 10 + 2 * l))
            m.lin2.weight.copy_(O.xavier_normal_((D, D), seed + 11 + 2 * l))
            # non-zero biases exercise the bias path (the reference initialises them to 0)
            m.lin1.bias.copy_(0.01 * O.xavier_normal_((1, D), seed + 30 + l)[0])
            m.lin2.bias.copy_(0.01 * O.xavier_normal_((1, D), seed + 40 + l)[0])
        gnn.append(m)
        out[f"ngcf_w1_{l}"], out[f"ngcf_b1_{l}"] = m.lin1.weight.detach(), m.lin1.bias.detach()
        out[f"ngcf_w2_{l}"], out[f"ngcf_b2_{l}"] = m.lin2.weight.detach(), m.lin2.bias.detach()
    gmask = torch.Generator().manual_seed(seed + 99)
    masks = [torch.rand(N, D, generator=gmask) >= 0.1 for _ in range(3)]
    out["ngcf_masks_packed"] = np.packbits(torch.stack(masks, 0).numpy().reshape(-1))
    with torch.no_grad():
        out["bignn_layer0"] = gnn[0](x0n, edge_index, edge_weight)       # layers.py:54-58
        for tag, p in (("p0", 0.0), ("p01", 0.1)):
            embs, e = [x0n], x0n
            for l, m in enumerate(gnn):
                e = m(e, edge_index, edge_weight)
                e = nn.LeakyReLU(negative_slope=0.2)(e)
                if p > 0:
                    e = e * masks[l].float() / (1 - p)          # nn.Dropout(p) with the recorded mask
                e = F.normalize(e, p=2, dim=1)
                embs.append(e)
            # dropout variant: only the last layer's slice is stored (it depends on all earlier masks)
            out[f"ngcf_{tag}"] = torch.cat(embs, dim=1) if p == 0 else embs[-1]

    # rectangular BipartiteGCNConv fed by get_bipartite_inter_mat (dataset.py:81-106, layers.py:31-35)
    bconv = layers.BipartiteGCNConv(D)
    for row in ("user", "item"):
        for rn in (True, False):
            ei, ew = ds.get_bipartite_inter_mat(row=row, row_norm=rn)
            n_row, n_col = (U, I) if row == "user" else (I, U)
            x_row = xu if row == "user" else xi
            x_col = xi if row == "user" else xu
            tag = f"bip_{row}_{'rown' if rn else 'sym'}"
            out[tag + "_ei"], out[tag + "_ew"] = ei, ew
            # callers flip so that sources are the `col` side and destinations the `row` side
            # (diffnet.py:97): out[row] = sum w * x_col[col]
            out[tag + "_y"] = bconv((x_col, x_row), ei.flip([0]), ew, size=(n_col, n_row))
    return out


def save(path, d):
    arrs = {}
    for k, v in d.items():
        a = v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
        if a.dtype == np.int64 and a.size and a.max() < 2 ** 31:
            a = a.astype(np.int32)
        arrs[k] = a
    np.savez_compressed(path, **arrs)
    print(path, "%.2f MB" % (os.path.getsize(path) / 1e6))


def main():
    from oracle import oracle as O

    torch.manual_seed(0)
    install_standins()
    layers = load_ref("recbole_gnn/model/layers.py", "ref_layers")
    dataset = load_ref("recbole_gnn/data/dataset.py", "ref_dataset")
    ds_cls = dataset.GeneralGraphDataset

    # G1: the reference's own fixture graph
    uid, iid, U, I = O.load_inter_file(os.path.join(REF, "tests/test_data/test/test.inter"))
    assert (U, I, uid.numel()) == (347, 1125, 5999), (U, I, uid.numel())
    g1 = model_loops(layers, ds_cls, uid, iid, U, I, 64, seed=0)
    g1.update(uid=uid, iid=iid, U=U, I=I)
    save(os.path.join(HERE, "g1_fixture.npz"), g1)

    # G2: 3 real users (+PAD) x 2 real items (+PAD, + one isolated item); duplicate (1,1) interaction
    uid = torch.tensor([1, 1, 2, 3, 1], dtype=torch.int64)
    iid = torch.tensor([1, 2, 1, 2, 1], dtype=torch.int64)
    g2 = model_loops(layers, ds_cls, uid, iid, 4, 4, 8, seed=5)
    g2.update(uid=uid, iid=iid, U=4, I=4)
    save(os.path.join(HERE, "g2_toy.npz"), g2)


if __name__ == "__main__":
    main()
