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
restated.

The model files are ALSO executed unmodified (round 2):

    /root/reference/recbole_gnn/model/abstract_recommender.py          (GeneralGraphRecommender.__init__)
    /root/reference/recbole_gnn/model/general_recommender/lightgcn.py  (LightGCN.forward / calculate_loss)
    /root/reference/recbole_gnn/model/general_recommender/ngcf.py      (NGCF.forward / calculate_loss)
    /root/reference/recbole_gnn/model/general_recommender/simgcl.py    (SimGCL.forward / calculate_loss)

over stand-ins for ``recbole.model.abstract_recommender.GeneralRecommender`` (the six attributes its
``__init__`` sets in recbole 1.1.1), ``recbole.model.init`` (xavier_*_initialization), ``recbole.model.loss``
(``BPRLoss``, ``EmbLoss`` restated from recbole 1.1.1), ``recbole.utils.InputType/ModelType`` and
``torch_geometric.utils.dropout_adj``.  Random draws inside the reference code are made reproducible from
OUTSIDE the files: ``torch.rand_like`` (simgcl.py:31) and ``nn.Dropout`` (ngcf.py:97) are patched for the
duration of the call to return / apply recorded draws, which are stored next to the outputs.

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
    import enum

    class InputType(enum.Enum):          # recbole.utils.enum_type.InputType
        POINTWISE = 1
        PAIRWISE = 2
        LISTWISE = 3

    class RecBoleModelType(enum.Enum):   # recbole.utils.enum_type.ModelType
        GENERAL = 1
        SEQUENTIAL = 2
        CONTEXT = 3
        KNOWLEDGE = 4
        TRADITIONAL = 5
        DECISIONTREE = 6

    class GnnModelType(enum.Enum):       # recbole_gnn/utils.py:159-165
        SOCIAL = 7

    _module("recbole.utils", set_color=lambda s, c: s, FeatureSource=object, ensure_dir=lambda d: None,
            InputType=InputType, ModelType=RecBoleModelType, Enum=enum.Enum)

    class GeneralRecommender(nn.Module):
        """recbole 1.1.1 ``GeneralRecommender.__init__``: field names, vocabulary sizes, device."""
        type = RecBoleModelType.GENERAL

        def __init__(self, config, dataset):
            super().__init__()
            self.USER_ID = config["USER_ID_FIELD"]
            self.ITEM_ID = config["ITEM_ID_FIELD"]
            self.NEG_ITEM_ID = config["NEG_PREFIX"] + self.ITEM_ID
            self.n_users = dataset.num(self.USER_ID)
            self.n_items = dataset.num(self.ITEM_ID)
            self.device = config["device"]

    def xavier_uniform_initialization(module):
        if isinstance(module, nn.Embedding):
            nn.init.xavier_uniform_(module.weight.data)
        elif isinstance(module, nn.Linear):
            nn.init.xavier_uniform_(module.weight.data)
            if module.bias is not None:
                nn.init.constant_(module.bias.data, 0)

    def xavier_normal_initialization(module):
        if isinstance(module, nn.Embedding):
            nn.init.xavier_normal_(module.weight.data)
        elif isinstance(module, nn.Linear):
            nn.init.xavier_normal_(module.weight.data)
            if module.bias is not None:
                nn.init.constant_(module.bias.data, 0)

    class BPRLoss(nn.Module):            # recbole 1.1.1 recbole/model/loss.py
        def __init__(self, gamma=1e-10):
            super().__init__()
            self.gamma = gamma

        def forward(self, pos_score, neg_score):
            return -torch.log(self.gamma + torch.sigmoid(pos_score - neg_score)).mean()

    class EmbLoss(nn.Module):            # recbole 1.1.1 recbole/model/loss.py
        def __init__(self, norm=2):
            super().__init__()
            self.norm = norm

        def forward(self, *embeddings, require_pow=False):
            emb_loss = torch.zeros(1).to(embeddings[-1].device)
            if require_pow:
                for embedding in embeddings:
                    emb_loss += torch.pow(input=torch.norm(embedding, p=self.norm), exponent=self.norm)
                emb_loss /= embeddings[-1].shape[0]
                emb_loss /= self.norm
                return emb_loss
            for embedding in embeddings:
                emb_loss += torch.norm(embedding, p=self.norm)
            emb_loss /= embeddings[-1].shape[0]
            return emb_loss

    _module("recbole.model")
    _module("recbole.model.abstract_recommender", GeneralRecommender=GeneralRecommender)
    _module("recbole.model.init", xavier_uniform_initialization=xavier_uniform_initialization,
            xavier_normal_initialization=xavier_normal_initialization)
    _module("recbole.model.loss", BPRLoss=BPRLoss, EmbLoss=EmbLoss)

    def dropout_adj(edge_index, edge_attr=None, p=0.5, force_undirected=False, num_nodes=None, training=True):
        """PyG ``dropout_adj``: Bernoulli(1-p) keep mask over the edges, no rescale."""
        if not training or p == 0.0:
            return edge_index, edge_attr
        mask = torch.rand(edge_index.size(1)) >= p
        return edge_index[:, mask], (None if edge_attr is None else edge_attr[mask])

    sys.modules["torch_geometric.utils"].dropout_adj = dropout_adj
    # the reference's own package namespace: only recbole_gnn.utils.ModelType is read (abstract_recommender.py:4)
    for name in ("recbole_gnn", "recbole_gnn.model", "recbole_gnn.model.general_recommender"):
        m = _module(name)
        m.__path__ = []
    _module("recbole_gnn.utils", ModelType=GnnModelType)
    return rb


def load_ref_models(layers_mod):
    """Import abstract_recommender.py, lightgcn.py, ngcf.py, simgcl.py from /root/reference unmodified."""
    sys.modules["recbole_gnn.model.layers"] = layers_mod
    ar = load_ref("recbole_gnn/model/abstract_recommender.py", "recbole_gnn.model.abstract_recommender")
    sys.modules["recbole_gnn.model.abstract_recommender"] = ar
    lg = load_ref("recbole_gnn/model/general_recommender/lightgcn.py", "recbole_gnn.model.general_recommender.lightgcn")
    sys.modules["recbole_gnn.model.general_recommender"].LightGCN = lg.LightGCN      # simgcl.py:13
    ng = load_ref("recbole_gnn/model/general_recommender/ngcf.py", "recbole_gnn.model.general_recommender.ngcf")
    sg = load_ref("recbole_gnn/model/general_recommender/simgcl.py", "recbole_gnn.model.general_recommender.simgcl")
    return lg.LightGCN, ng.NGCF, sg.SimGCL


class _Patched:
    """Temporarily replace attributes (torch.rand_like / torch.nn.Dropout) around a call into reference code."""

    def __init__(self, *triples):
        self.triples = triples

    def __enter__(self):
        self.old = [getattr(o, n) for o, n, _ in self.triples]
        for o, n, v in self.triples:
            setattr(o, n, v)

    def __exit__(self, *a):
        for (o, n, _), v in zip(self.triples, self.old):
            setattr(o, n, v)


def ref_config(**kw):
    cfg = {"USER_ID_FIELD": "user_id", "ITEM_ID_FIELD": "item_id", "NEG_PREFIX": "neg_", "device": torch.device("cpu"),
           "enable_sparse": False, "embedding_size": 64, "n_layers": 3, "reg_weight": 1e-5, "require_pow": False,
           "lambda": 0.1, "eps": 0.1, "temperature": 0.2, "hidden_size_list": [64, 64, 64], "node_dropout": 0.0,
           "message_dropout": 0.0}
    cfg.update(kw)
    return cfg


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


def model_loops(layers, ds_cls, uid, iid, U, I, D, seed, models=None):
    """Run the reference layers, dataset methods and (``models`` = the three reference model classes) the
    reference's own forward()/calculate_loss()."""
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
    def set_tables(model, tu, ti):
        with torch.no_grad():
            model.user_embedding.weight.copy_(tu)
            model.item_embedding.weight.copy_(ti)

    LightGCN, NGCF, SimGCL = models
    # one recorded BPR batch (user, positive item, sampled negative item), ids >= 1
    gb = torch.Generator().manual_seed(seed + 123)
    B = min(256, uid.numel())
    pick = torch.randperm(uid.numel(), generator=gb)[:B]
    batch = {"user_id": uid[pick], "item_id": iid[pick],
             "neg_item_id": torch.randint(1, I, (B,), generator=gb, dtype=torch.int64)}
    out["batch_user"], out["batch_pos"], out["batch_neg"] = batch["user_id"], batch["item_id"], batch["neg_item_id"]

    # ---- LightGCN.forward / calculate_loss (lightgcn.py:70-110), dense-edge and SparseTensor graphs
    for L in (2, 3):
        for sparse in (False, True):
            m = LightGCN(ref_config(embedding_size=D, n_layers=L, enable_sparse=sparse), ds)
            set_tables(m, xu, xi)
            with torch.no_grad():
                u, i = m.forward()
            key = f"lightgcn_L{L}" + ("_sparse" if sparse else "")
            out[key] = torch.cat([u, i], 0)
    for rp in (False, True):
        m = LightGCN(ref_config(embedding_size=D, n_layers=3, require_pow=rp), ds)
        set_tables(m, xu, xi)
        loss = m.calculate_loss(batch)
        loss.backward()
        tag = "pow" if rp else "nopow"
        out[f"lightgcn_loss_{tag}"] = loss.detach().reshape(1)
        out[f"lightgcn_loss_{tag}_gu"] = m.user_embedding.weight.grad.clone()
        out[f"lightgcn_loss_{tag}_gi"] = m.item_embedding.weight.grad.clone()

    # ---- SimGCL.forward(perturbed) / calculate_loss (simgcl.py:24-60), L=3.  rand_like draws U[0,1); the
    # recorded draws are quantised to k/256 so that they store as uint8 (still valid U[0,1) samples).
    g = torch.Generator().manual_seed(seed + 7)
    L = 3
    noise_u8 = torch.randint(0, 256, (3 * L, N, D), generator=g, dtype=torch.uint8)   # 3 perturbed forwards
    queue = []

    def fake_rand_like(t, **kw):
        return queue.pop(0)

    m = SimGCL(ref_config(embedding_size=D, n_layers=L), ds)
    set_tables(m, xu, xi)
    with torch.no_grad():
        queue[:] = [noise_u8[l].float() / 256.0 for l in range(L)]
        with _Patched((torch, "rand_like", fake_rand_like)):
            u, i = m.forward(perturbed=True)
        assert not queue
        out["simgcl_L3"] = torch.cat([u, i], 0)
        u, i = m.forward()
        out["simgcl_clean_L3"] = torch.cat([u, i], 0)
    out["simgcl_L3_noise_u8"] = noise_u8[:L]
    queue[:] = [noise_u8[l].float() / 256.0 for l in range(L, 3 * L)]
    with _Patched((torch, "rand_like", fake_rand_like)):
        loss = m.calculate_loss(batch)                                  # 1 clean + 2 perturbed forwards
    assert not queue
    loss.backward()
    out["simgcl_loss"] = loss.detach().reshape(1)
    out["simgcl_loss_gu"] = m.user_embedding.weight.grad.clone()
    out["simgcl_loss_gi"] = m.item_embedding.weight.grad.clone()
    out["simgcl_loss_noise_u8"] = noise_u8[L:]

    # ---- NGCF.forward / calculate_loss (ngcf.py:73-125): hidden [D,D,D] (NGCF.yaml), xavier_normal weights;
    # message_dropout 0 and 0.1 with a recorded mask (nn.Dropout at ngcf.py:97 is patched to apply it)
    xun = O.xavier_normal_((U, D), seed + 2)
    xin = O.xavier_normal_((I, D), seed + 3)
    x0n = torch.cat([xun, xin], 0)
    out["ngcf_xu"], out["ngcf_xi"] = xun, xin
    gmask = torch.Generator().manual_seed(seed + 99)
    masks = [torch.rand(N, D, generator=gmask) >= 0.1 for _ in range(3)]
    out["ngcf_masks_packed"] = np.packbits(torch.stack(masks, 0).numpy().reshape(-1))

    def ngcf_model(p):
        m = NGCF(ref_config(embedding_size=D, hidden_size_list=[D, D, D], message_dropout=p), ds)
        set_tables(m, xun, xin)
        for l, layer in enumerate(m.GNNlayers):
            with torch.no_grad():
                layer.lin1.weight.copy_(O.xavier_normal_((D, D), seed + 10 + 2 * l))
                layer.lin2.weight.copy_(O.xavier_normal_((D, D), seed + 11 + 2 * l))
                # non-zero biases exercise the bias path (the reference initialises them to 0)
                layer.lin1.bias.copy_(0.01 * O.xavier_normal_((1, D), seed + 30 + l)[0])
                layer.lin2.bias.copy_(0.01 * O.xavier_normal_((1, D), seed + 40 + l)[0])
        return m

    class RecordedDropout(nn.Module):
        """nn.Dropout(p) in training mode with the Bernoulli keep-mask taken from ``masks`` in call order."""
        calls = 0

        def __init__(self, p=0.5):
            super().__init__()
            self.p = p

        def forward(self, x):
            if self.p == 0:
                return x
            k = RecordedDropout.calls
            RecordedDropout.calls += 1
            return x * masks[k % 3].float() / (1 - self.p)

    m = ngcf_model(0.0)
    for l, layer in enumerate(m.GNNlayers):
        out[f"ngcf_w1_{l}"], out[f"ngcf_b1_{l}"] = layer.lin1.weight.detach().clone(), layer.lin1.bias.detach().clone()
        out[f"ngcf_w2_{l}"], out[f"ngcf_b2_{l}"] = layer.lin2.weight.detach().clone(), layer.lin2.bias.detach().clone()
    with torch.no_grad():
        out["bignn_layer0"] = m.GNNlayers[0](x0n, edge_index, edge_weight)       # layers.py:54-58
        u, i = m.forward()
        out["ngcf_p0"] = torch.cat([u, i], 0)
    loss = m.calculate_loss(batch)
    loss.backward()
    out["ngcf_loss_p0"] = loss.detach().reshape(1)
    out["ngcf_loss_p0_gu"] = m.user_embedding.weight.grad.clone()
    out["ngcf_loss_p0_gw1_0"] = m.GNNlayers[0].lin1.weight.grad.clone()
    out["ngcf_loss_p0_gw2_2"] = m.GNNlayers[2].lin2.weight.grad.clone()
    out["ngcf_loss_p0_gb1_1"] = m.GNNlayers[1].lin1.bias.grad.clone()
    m = ngcf_model(0.1)
    with torch.no_grad(), _Patched((nn, "Dropout", RecordedDropout)):
        u, i = m.forward()
    # dropout variant: only the last layer's slice is stored (it depends on all earlier masks)
    out["ngcf_p01"] = torch.cat([u, i], 0)[:, 3 * D:]

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
    models = load_ref_models(layers)

    # G1: the reference's own fixture graph
    uid, iid, U, I = O.load_inter_file(os.path.join(REF, "tests/test_data/test/test.inter"))
    assert (U, I, uid.numel()) == (347, 1125, 5999), (U, I, uid.numel())
    g1 = model_loops(layers, ds_cls, uid, iid, U, I, 64, seed=0, models=models)
    g1.update(uid=uid, iid=iid, U=U, I=I)
    save(os.path.join(HERE, "g1_fixture.npz"), g1)

    # G2: 3 real users (+PAD) x 2 real items (+PAD, + one isolated item); duplicate (1,1) interaction
    uid = torch.tensor([1, 1, 2, 3, 1], dtype=torch.int64)
    iid = torch.tensor([1, 2, 1, 2, 1], dtype=torch.int64)
    g2 = model_loops(layers, ds_cls, uid, iid, 4, 4, 8, seed=5, models=models)
    g2.update(uid=uid, iid=iid, U=4, I=4)
    save(os.path.join(HERE, "g2_toy.npz"), g2)


if __name__ == "__main__":
    main()
