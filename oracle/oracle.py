"""CPU oracle for the bipartite graph-convolution hot path (TEST INFRASTRUCTURE ONLY).

This module is a CPU restatement, in plain torch/numpy, of the reference's algorithm for
the path SURVEY.md §8 names.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import it.  Nothing under ``recbole_gnn_b200/`` imports it, and the product path raises
when the CUDA library is missing instead of falling back to this file.

Parity status: PARTIALLY PINNED.  The reference (RUCAIBox/RecBole-GNN @ 632ef888) holds no
golden vectors or numeric assertions for this path (``tests/test_model.py`` is smoke-only),
and its arithmetic lives in third-party wheels that are absent here and have no source
under /root/reference:

* ``torch_geometric`` (README pins ``pyg>=2.0.4``): ``MessagePassing.propagate``,
  ``gcn_norm``, ``degree``, ``dropout_adj``;
* ``torch_sparse`` (unpinned): ``SparseTensor``, ``matmul``;
* ``recbole==1.1.1``: ``Dataset.inter_feat``, ``GeneralRecommender``.

What pins this oracle: ``tests/golden/make_golden.py`` executes the reference's OWN source
files (``recbole_gnn/model/layers.py``, ``recbole_gnn/data/dataset.py`` and, since round 2, the model files
``abstract_recommender.py``, ``lightgcn.py``, ``ngcf.py``, ``simgcl.py`` — ``forward()`` and
``calculate_loss()`` as written) in this container
over minimal stand-ins for those three packages (the stand-ins restate the packages'
documented semantics: sum-aggregation ``propagate``, ``gcn_norm(add_self_loops=False)``,
``degree``) and stores the outputs as fixtures; ``tests/test_oracle.py`` checks every function
below against them.  With respect to the third-party kernels themselves parity is UNPINNED.

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# graph construction
# --------------------------------------------------------------------------------------
def gcn_norm(edge_index: Tensor, edge_weight: Optional[Tensor], num_nodes: int) -> Tuple[Tensor, Tensor]:
    """PyG ``gcn_norm(edge_index, edge_weight, num_nodes, add_self_loops=False)`` with the default
    ``flow='source_to_target'``; call sites recbole_gnn/data/dataset.py:74,77 and sgl.py:121,124.

    deg[c] = sum of w over edges whose TARGET (row 1) is c; dis = deg^-1/2 with inf -> 0;
    w' = dis[row0] * w * dis[row1].
    """
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=torch.float32)
    row, col = edge_index[0], edge_index[1]
    deg = torch.zeros(num_nodes, dtype=edge_weight.dtype).scatter_add_(0, col, edge_weight)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    return edge_index, dis[row] * edge_weight * dis[col]


def build_norm_adj(uid: Tensor, iid: Tensor, user_num: int, item_num: int) -> Tuple[Tensor, Tensor]:
    """``GeneralGraphDataset.get_norm_adj_mat(enable_sparse=False)``
    (recbole_gnn/data/dataset.py:49-79).  ``uid``/``iid`` are the int64 ``inter_feat`` columns
    (ids >= 1, 0 is RecBole's [PAD]).  Duplicate interactions stay parallel edges."""
    row = uid                                     # dataset.py:60
    col = iid + user_num                          # dataset.py:61
    edge_index1 = torch.stack([row, col])         # dataset.py:62
    edge_index2 = torch.stack([col, row])         # dataset.py:63
    edge_index = torch.cat([edge_index1, edge_index2], dim=1)   # dataset.py:64
    edge_weight = torch.ones(edge_index.size(1))  # dataset.py:65
    num_nodes = user_num + item_num               # dataset.py:66
    return gcn_norm(edge_index, edge_weight, num_nodes)         # dataset.py:77


def degree(index: Tensor, num_nodes: int) -> Tensor:
    """PyG ``degree(index, num_nodes)`` (float32 count); call sites dataset.py:94,98-99."""
    return torch.zeros(num_nodes, dtype=torch.float32).scatter_add_(
        0, index, torch.ones(index.numel(), dtype=torch.float32))


def build_bipartite_inter_mat(row_ids: Tensor, col_ids: Tensor, n_row: int, n_col: int,
                              row_norm: bool = True) -> Tuple[Tensor, Tensor]:
    """``GeneralGraphDataset.get_bipartite_inter_mat`` (recbole_gnn/data/dataset.py:81-106)."""
    edge_index = torch.stack([row_ids, col_ids])                      # dataset.py:91
    if row_norm:
        deg = degree(edge_index[0], n_row)                            # dataset.py:94
        norm_deg = 1.0 / torch.where(deg == 0, torch.ones([1]), deg)  # dataset.py:95
        edge_weight = norm_deg[edge_index[0]]                         # dataset.py:96
    else:
        row_deg = degree(edge_index[0], n_row)                        # dataset.py:98
        col_deg = degree(edge_index[1], n_col)                        # dataset.py:99
        row_norm_deg = 1.0 / torch.sqrt(torch.where(row_deg == 0, torch.ones([1]), row_deg))
        col_norm_deg = 1.0 / torch.sqrt(torch.where(col_deg == 0, torch.ones([1]), col_deg))
        edge_weight = row_norm_deg[edge_index[0]] * col_norm_deg[edge_index[1]]   # dataset.py:104
    return edge_index, edge_weight


def dropout_adj(edge_index: Tensor, edge_weight: Tensor, keep_mask: Tensor) -> Tuple[Tensor, Tensor]:
    """PyG ``dropout_adj(edge_index, edge_attr, p, training=True)`` with the Bernoulli mask handed in:
    keeps edges where ``keep_mask`` is True, NO rescale (call sites ngcf.py:81,89)."""
    return edge_index[:, keep_mask], edge_weight[keep_mask]


# --------------------------------------------------------------------------------------
# one propagation layer
# --------------------------------------------------------------------------------------
def propagate_scatter(x: Tensor, edge_index: Tensor, edge_weight: Tensor, n_dst: Optional[int] = None) -> Tensor:
    """Dense-edge path of ``LightGCNConv``/``BipartiteGCNConv`` (recbole_gnn/model/layers.py:13-17, 31-35):
    x_j = x[edge_index[0]]; msg = w.view(-1,1)*x_j; out = scatter-add onto edge_index[1]."""
    n_dst = x.size(0) if n_dst is None else n_dst
    msg = edge_weight.view(-1, 1) * x.index_select(0, edge_index[0])
    return torch.zeros(n_dst, x.size(1), dtype=x.dtype).index_add_(0, edge_index[1], msg)


def adj_sparse(edge_index: Tensor, edge_weight: Tensor, n_dst: int, n_src: int, layout: str = "coo"):
    """The sparse object of the ``enable_sparse`` path: ``SparseTensor(row=ei[0], col=ei[1]).t()``
    (dataset.py:41-47) = matrix with rows = targets, cols = sources."""
    a = torch.sparse_coo_tensor(torch.stack([edge_index[1], edge_index[0]]), edge_weight,
                                (n_dst, n_src)).coalesce()
    return a.to_sparse_csr() if layout == "csr" else a


def propagate_sparse(a, x: Tensor) -> Tensor:
    """``matmul(adj_t, x, reduce='add')`` (layers.py:19-20) restated as ``torch.sparse.mm`` —
    the form BASELINE.json's north_star names."""
    return torch.sparse.mm(a, x)


def propagate_f64(x: Tensor, edge_index: Tensor, edge_weight: Tensor, n_dst: Optional[int] = None) -> Tensor:
    """float64 tie-breaker for fp32 ordering differences."""
    return propagate_scatter(x.double(), edge_index, edge_weight.double(), n_dst)


# --------------------------------------------------------------------------------------
# model forwards (K-layer loops)
# --------------------------------------------------------------------------------------
def lightgcn_forward(xu: Tensor, xi: Tensor, edge_index: Tensor, edge_weight: Tensor, n_layers: int,
                     prop=None) -> Tuple[Tensor, Tensor]:
    """``LightGCN.forward`` (recbole_gnn/model/general_recommender/lightgcn.py:70-81)."""
    prop = prop or (lambda x: propagate_scatter(x, edge_index, edge_weight))
    all_embeddings = torch.cat([xu, xi], dim=0)            # lightgcn.py:60-68
    embeddings_list = [all_embeddings]
    for _ in range(n_layers):                              # lightgcn.py:74-76
        all_embeddings = prop(all_embeddings)
        embeddings_list.append(all_embeddings)
    out = torch.stack(embeddings_list, dim=1).mean(dim=1)  # lightgcn.py:77-78
    return torch.split(out, [xu.size(0), xi.size(0)])      # lightgcn.py:80


def simgcl_forward(xu: Tensor, xi: Tensor, edge_index: Tensor, edge_weight: Tensor, n_layers: int,
                   eps: float, noises: Optional[Sequence[Tensor]] = None) -> Tuple[Tensor, Tensor]:
    """``SimGCL.forward(perturbed)`` (simgcl.py:24-38).  ``noises`` = the per-layer ``torch.rand_like``
    draws (U[0,1)), handed in so the device path can be compared element-wise; None = unperturbed."""
    all_embs = torch.cat([xu, xi], dim=0)
    embeddings_list = []                                   # simgcl.py:26 (ego layer excluded)
    for l in range(n_layers):
        all_embs = propagate_scatter(all_embs, edge_index, edge_weight)      # simgcl.py:29
        if noises is not None:                                               # simgcl.py:30-32
            all_embs = all_embs + torch.sign(all_embs) * F.normalize(noises[l], dim=-1) * eps
        embeddings_list.append(all_embs)
    out = torch.stack(embeddings_list, dim=1).mean(dim=1)  # simgcl.py:34-35
    return torch.split(out, [xu.size(0), xi.size(0)])


def bignn_layer(x: Tensor, edge_index: Tensor, edge_weight: Tensor, w1: Tensor, b1: Tensor,
                w2: Tensor, b2: Tensor) -> Tensor:
    """``BiGNNConv.forward`` (layers.py:54-58): lin1(Âx + x) + lin2(Âx ⊙ x)."""
    x_prop = propagate_scatter(x, edge_index, edge_weight)
    x_trans = F.linear(x_prop + x, w1, b1)
    x_inter = F.linear(torch.mul(x_prop, x), w2, b2)
    return x_trans + x_inter


def ngcf_forward(xu: Tensor, xi: Tensor, edge_index: Tensor, edge_weight: Tensor,
                 weights: Sequence[Tuple[Tensor, Tensor, Tensor, Tensor]],
                 message_dropout: float = 0.0,
                 drop_masks: Optional[Sequence[Tensor]] = None) -> Tuple[Tensor, Tensor]:
    """``NGCF.forward`` with node_dropout == 0 (ngcf.py:92-104).  ``nn.Dropout`` at ngcf.py:97 is a
    freshly built module, hence always in training mode; its Bernoulli keep-mask is handed in
    (``drop_masks[l]``, bool [N, D_l]) so results are comparable; None = message_dropout 0."""
    x = torch.cat([xu, xi], dim=0)
    embeddings_list = [x]
    for l, (w1, b1, w2, b2) in enumerate(weights):
        x = bignn_layer(x, edge_index, edge_weight, w1, b1, w2, b2)        # ngcf.py:95
        x = F.leaky_relu(x, negative_slope=0.2)                             # ngcf.py:96
        if drop_masks is not None and message_dropout > 0:                  # ngcf.py:97
            x = x * drop_masks[l].to(x.dtype) / (1.0 - message_dropout)
        x = F.normalize(x, p=2, dim=1)                                      # ngcf.py:98
        embeddings_list.append(x)
    out = torch.cat(embeddings_list, dim=1)                                 # ngcf.py:100
    return torch.split(out, [xu.size(0), xi.size(0)])                       # ngcf.py:102


def bipartite_forward(x_src: Tensor, edge_index: Tensor, edge_weight: Tensor, n_dst: int) -> Tensor:
    """``BipartiteGCNConv.forward(x=(x_src, x_dst), edge_index, edge_weight, size=(n_src, n_dst))``
    (layers.py:31-35): only x_src is read; edge_index[0] = source ids, [1] = destination ids."""
    return propagate_scatter(x_src, edge_index, edge_weight, n_dst)


# --------------------------------------------------------------------------------------
# per-epoch graph re-sampling of the callers
# --------------------------------------------------------------------------------------
def sgl_augmented_adj(uid: Tensor, iid: Tensor, user_num: int, item_num: int, aug_type: str,
                      keep_idx: Optional[Tensor] = None, drop_user: Optional[Tensor] = None,
                      drop_item: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """``SGL.random_graph_augment`` (sgl.py:92-126) with the ``np.random.choice`` draws handed in: ND drops the
    interactions that touch a dropped user / item (`:95-104`), ED / RW keep the sampled interaction indices
    (`:106-109`); then the symmetric COO + ``gcn_norm`` exactly as ``get_norm_adj_mat`` (`:111-124`)."""
    if aug_type == "ND":
        mask = torch.isin(uid, drop_user) | torch.isin(iid, drop_item)
        keep_idx = torch.nonzero(~mask).flatten()
    return build_norm_adj(uid[keep_idx], iid[keep_idx], user_num, item_num)


def sept_norm_edge_weight(edge_index: Tensor, node_num: int) -> Tensor:
    """``SEPT.get_norm_edge_weight`` (sept.py:81-87)."""
    deg = degree(edge_index[0], node_num)
    norm_deg = 1. / torch.sqrt(torch.where(deg == 0, torch.ones([1]), deg))
    return norm_deg[edge_index[0]] * norm_deg[edge_index[1]]


def sept_subgraph(uid: Tensor, iid: Tensor, src_user: Tensor, tgt_user: Tensor, user_num: int, item_num: int,
                  keep: Tensor, net_keep: Tensor) -> Tuple[Tensor, Tensor]:
    """``SEPT.subgraph_construction`` (sept.py:111-133) with the two index draws handed in."""
    row, col = uid[keep], iid[keep] + user_num
    edge_index = torch.cat([torch.stack([row, col]), torch.stack([col, row]),
                            torch.stack([src_user[net_keep], tgt_user[net_keep]])], dim=1)
    return edge_index, sept_norm_edge_weight(edge_index, user_num + item_num)


# --------------------------------------------------------------------------------------
# training losses around the path
# --------------------------------------------------------------------------------------
def bpr_loss(pos_score: Tensor, neg_score: Tensor, gamma: float = 1e-10) -> Tensor:
    """recbole 1.1.1 ``BPRLoss`` (third-party; call sites lightgcn.py:100, ngcf.py:119)."""
    return -torch.log(gamma + torch.sigmoid(pos_score - neg_score)).mean()


def emb_loss(*embeddings: Tensor, require_pow: bool = False, norm: int = 2) -> Tensor:
    """recbole 1.1.1 ``EmbLoss`` (third-party; call sites lightgcn.py:107, ngcf.py:121)."""
    loss = torch.zeros(1, dtype=embeddings[-1].dtype)
    if require_pow:
        for e in embeddings:
            loss = loss + torch.pow(torch.norm(e, p=norm), norm)
        return loss / embeddings[-1].shape[0] / norm
    for e in embeddings:
        loss = loss + torch.norm(e, p=norm)
    return loss / embeddings[-1].shape[0]


def lightgcn_loss(xu: Tensor, xi: Tensor, edge_index: Tensor, edge_weight: Tensor, n_layers: int,
                  user: Tensor, pos_item: Tensor, neg_item: Tensor, reg_weight: float = 1e-5,
                  require_pow: bool = False, forward=None) -> Tensor:
    """``LightGCN.calculate_loss`` (lightgcn.py:83-110)."""
    ua, ia = forward() if forward is not None else lightgcn_forward(xu, xi, edge_index, edge_weight, n_layers)
    u, pos, neg = ua[user], ia[pos_item], ia[neg_item]                      # lightgcn.py:93-95
    mf = bpr_loss(torch.mul(u, pos).sum(dim=1), torch.mul(u, neg).sum(dim=1))   # lightgcn.py:98-100
    reg = emb_loss(xu[user], xi[pos_item], xi[neg_item], require_pow=require_pow)   # lightgcn.py:103-107
    return mf + reg_weight * reg                                            # lightgcn.py:108


def simgcl_cl_loss(x1: Tensor, x2: Tensor, temperature: float) -> Tensor:
    """``SimGCL.calculate_cl_loss`` (simgcl.py:40-46): InfoNCE over the rows of x1/x2."""
    x1, x2 = F.normalize(x1, dim=-1), F.normalize(x2, dim=-1)
    pos_score = torch.exp((x1 * x2).sum(dim=-1) / temperature)
    ttl_score = torch.exp(torch.matmul(x1, x2.transpose(0, 1)) / temperature).sum(dim=1)
    return -torch.log(pos_score / ttl_score).sum()


def simgcl_loss(xu: Tensor, xi: Tensor, edge_index: Tensor, edge_weight: Tensor, n_layers: int, eps: float,
                noises1: Sequence[Tensor], noises2: Sequence[Tensor], user: Tensor, pos_item: Tensor,
                neg_item: Tensor, reg_weight: float = 1e-5, require_pow: bool = False, cl_rate: float = 0.1,
                temperature: float = 0.2) -> Tensor:
    """``SimGCL.calculate_loss`` (simgcl.py:48-60).  Note ``super().calculate_loss`` calls ``self.forward()``
    = the SimGCL forward with perturbed=False (mean of x_1..x_L, no ego term)."""
    loss = lightgcn_loss(xu, xi, edge_index, edge_weight, n_layers, user, pos_item, neg_item, reg_weight,
                         require_pow, forward=lambda: simgcl_forward(xu, xi, edge_index, edge_weight, n_layers, eps))
    u_ids, i_ids = torch.unique(user), torch.unique(pos_item)               # simgcl.py:51-52
    u1, i1 = simgcl_forward(xu, xi, edge_index, edge_weight, n_layers, eps, noises1)
    u2, i2 = simgcl_forward(xu, xi, edge_index, edge_weight, n_layers, eps, noises2)
    user_cl = simgcl_cl_loss(u1[u_ids], u2[u_ids], temperature)
    item_cl = simgcl_cl_loss(i1[i_ids], i2[i_ids], temperature)
    return loss + cl_rate * (user_cl + item_cl)                              # simgcl.py:60


def ngcf_loss(xu: Tensor, xi: Tensor, edge_index: Tensor, edge_weight: Tensor, weights, user: Tensor,
              pos_item: Tensor, neg_item: Tensor, reg_weight: float = 1e-5, message_dropout: float = 0.0,
              drop_masks=None) -> Tensor:
    """``NGCF.calculate_loss`` (ngcf.py:106-123): BPR on the concatenated layer outputs, EmbLoss on the SAME
    propagated rows (not the ego tables), require_pow left at its default False."""
    ua, ia = ngcf_forward(xu, xi, edge_index, edge_weight, weights, message_dropout, drop_masks)
    u, pos, neg = ua[user], ia[pos_item], ia[neg_item]
    mf = bpr_loss(torch.mul(u, pos).sum(dim=1), torch.mul(u, neg).sum(dim=1))
    return mf + reg_weight * emb_loss(u, pos, neg)


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
def synth_interactions(user_num: int, item_num: int, n_inter: int, seed: int = 0,
                       zipf_alpha: Optional[float] = None) -> Tuple[Tensor, Tensor]:
    """Seeded synthetic ``inter_feat``: ids drawn from [1, num) so that [PAD]=0 has degree 0.
    ``zipf_alpha`` draws items from a Zipf-like law over a random permutation (skewed degrees)."""
    g = torch.Generator().manual_seed(seed)
    u = torch.randint(1, user_num, (n_inter,), generator=g, dtype=torch.int64)
    if zipf_alpha is None:
        i = torch.randint(1, item_num, (n_inter,), generator=g, dtype=torch.int64)
    else:
        ranks = torch.arange(1, item_num, dtype=torch.float64)
        p = ranks.pow(-zipf_alpha)
        idx = torch.multinomial(p / p.sum(), n_inter, replacement=True, generator=g)
        perm = torch.randperm(item_num - 1, generator=g) + 1
        i = perm[idx]
    return u, i


def xavier_uniform_table(rows: int, dim: int, seed: int) -> Tensor:
    """``xavier_uniform_initialization`` on an ``nn.Embedding`` weight (lightgcn.py:57)."""
    g = torch.Generator().manual_seed(seed)
    bound = math.sqrt(6.0 / (rows + dim))
    return (torch.rand(rows, dim, generator=g) * 2 - 1) * bound


def xavier_normal_(shape: Tuple[int, int], seed: int) -> Tensor:
    """``xavier_normal_initialization`` (ngcf.py:59) for a [fan_out, fan_in] weight."""
    g = torch.Generator().manual_seed(seed)
    std = math.sqrt(2.0 / (shape[0] + shape[1]))
    return torch.randn(*shape, generator=g) * std


def load_inter_file(path: str) -> Tuple[Tensor, Tensor, int, int]:
    """Read a RecBole ``.inter`` atomic file (header ``user_id:token\\titem_id:token...``) and remap
    tokens to ids in first-appearance order starting at 1 (0 = [PAD]), as RecBole's Dataset does."""
    users, items = {}, {}
    u_ids: List[int] = []
    i_ids: List[int] = []
    with open(path) as f:
        next(f)
        for line in f:
            parts = line.rstrip("\n").split("\t")
            if len(parts) < 2:
                continue
            u_ids.append(users.setdefault(parts[0], len(users) + 1))
            i_ids.append(items.setdefault(parts[1], len(items) + 1))
    return (torch.tensor(u_ids, dtype=torch.int64), torch.tensor(i_ids, dtype=torch.int64),
            len(users) + 1, len(items) + 1)
