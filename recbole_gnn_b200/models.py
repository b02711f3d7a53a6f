"""The three models on the benchmarked path with the reference's structure, attribute names and objective:
``LightGCN`` (general_recommender/lightgcn.py:36-133), ``NGCF`` (ngcf.py:33-150) and ``SimGCL``
(simgcl.py:16-60).  ``calculate_loss`` / ``predict`` / ``full_sort_predict`` follow the reference (BPR +
EmbLoss; SimGCL's InfoNCE term over two perturbed views; NGCF's EmbLoss on the propagated rows).

Each ``forward`` has two routes that give the same numbers:
* the reference's layer-by-layer loop over the drop-in conv layers (``fused_propagation: False``), and
* the fused engine entry point (default): one kernel per layer, layer-combine in the epilogue.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as F_
from .abstract_recommender import GeneralGraphRecommender, _cfg
from .graph import GraphHandle
from .layers import BiGNNConv, LightGCNConv, _resolve
from .loss import BPRLoss, EmbLoss


def _xavier_uniform(module):     # recbole.model.init.xavier_uniform_initialization
    if isinstance(module, nn.Embedding):
        nn.init.xavier_uniform_(module.weight.data)
    elif isinstance(module, nn.Linear):
        nn.init.xavier_uniform_(module.weight.data)
        if module.bias is not None:
            nn.init.constant_(module.bias.data, 0)


def _xavier_normal(module):      # recbole.model.init.xavier_normal_initialization
    if isinstance(module, nn.Embedding):
        nn.init.xavier_normal_(module.weight.data)
    elif isinstance(module, nn.Linear):
        nn.init.xavier_normal_(module.weight.data)
        if module.bias is not None:
            nn.init.constant_(module.bias.data, 0)


class _PropagationModel(GeneralGraphRecommender):
    """What LightGCN and NGCF share verbatim in the reference: the cached full-sort tables, ``predict`` and
    ``full_sort_predict`` (lightgcn.py:112-133 == ngcf.py:125-150)."""

    restore_user_e = None
    restore_item_e = None

    def _graph(self) -> GraphHandle:
        n = self.n_users + self.n_items
        return _resolve(self.edge_index, self.edge_weight, n, n)

    def get_ego_embeddings(self):
        return torch.cat([self.user_embedding.weight, self.item_embedding.weight], dim=0)

    def _clear_restore(self):
        if self.restore_user_e is not None or self.restore_item_e is not None:
            self.restore_user_e, self.restore_item_e = None, None

    def _bpr_rows(self, interaction):
        user = interaction[self.USER_ID]
        pos_item, neg_item = interaction[self.ITEM_ID], interaction[self.NEG_ITEM_ID]
        user_all, item_all = self.forward()
        u, pos, neg = user_all[user], item_all[pos_item], item_all[neg_item]
        mf_loss = self.mf_loss(torch.mul(u, pos).sum(dim=1), torch.mul(u, neg).sum(dim=1))
        return (user, pos_item, neg_item), (u, pos, neg), mf_loss

    def predict(self, interaction):
        user, item = interaction[self.USER_ID], interaction[self.ITEM_ID]
        user_all, item_all = self.forward()
        return torch.mul(user_all[user], item_all[item]).sum(dim=1)

    def full_sort_predict(self, interaction):
        user = interaction[self.USER_ID]
        if self.restore_user_e is None or self.restore_item_e is None:
            self.restore_user_e, self.restore_item_e = self.forward()
        return F_.full_sort_scores(self.restore_user_e[user], self.restore_item_e).view(-1)

    def full_sort_topk(self, interaction, k: int, history=None):
        """Top-``k`` item ids/scores per user of the batch without materialising the ``[batch, n_items]``
        score matrix (what RecBole's evaluator does with ``full_sort_predict`` + ``topk``); ``history`` =
        ``(row_idx, item_idx)`` of seen interactions to exclude, as RecBole's full-sort eval passes them."""
        user = interaction[self.USER_ID]
        if self.restore_user_e is None or self.restore_item_e is None:
            self.restore_user_e, self.restore_item_e = self.forward()
        return F_.full_sort_topk(self.restore_user_e[user], self.restore_item_e, k, history=history)


class LightGCN(_PropagationModel):
    def __init__(self, config, dataset):
        super(LightGCN, self).__init__(config, dataset)
        self.latent_dim = config['embedding_size']
        self.n_layers = config['n_layers']
        self.reg_weight = _cfg(config, 'reg_weight', 1e-5)
        self.require_pow = _cfg(config, 'require_pow', False)
        self.fused = _cfg(config, 'fused_propagation', True)
        self.user_embedding = torch.nn.Embedding(num_embeddings=self.n_users, embedding_dim=self.latent_dim)
        self.item_embedding = torch.nn.Embedding(num_embeddings=self.n_items, embedding_dim=self.latent_dim)
        self.gcn_conv = LightGCNConv(dim=self.latent_dim)
        self.mf_loss = BPRLoss()
        self.reg_loss = EmbLoss()
        self.restore_user_e = None
        self.restore_item_e = None
        self.apply(_xavier_uniform)
        self.other_parameter_name = ['restore_user_e', 'restore_item_e']

    def forward(self):
        if self.fused:
            return F_.lightgcn_propagate(self._graph(), self.user_embedding.weight, self.item_embedding.weight,
                                         self.n_layers)
        all_embeddings = self.get_ego_embeddings()
        embeddings_list = [all_embeddings]
        for _ in range(self.n_layers):
            all_embeddings = self.gcn_conv(all_embeddings, self.edge_index, self.edge_weight)
            embeddings_list.append(all_embeddings)
        out = torch.mean(torch.stack(embeddings_list, dim=1), dim=1)
        return torch.split(out, [self.n_users, self.n_items])

    def calculate_loss(self, interaction):
        self._clear_restore()
        (user, pos_item, neg_item), _, mf_loss = self._bpr_rows(interaction)
        reg_loss = self.reg_loss(self.user_embedding(user), self.item_embedding(pos_item),
                                 self.item_embedding(neg_item), require_pow=self.require_pow)
        return mf_loss + self.reg_weight * reg_loss


class SimGCL(LightGCN):
    def __init__(self, config, dataset):
        super(SimGCL, self).__init__(config, dataset)
        self.cl_rate = _cfg(config, 'lambda', 0.1)
        self.eps = _cfg(config, 'eps', 0.1)
        self.temperature = _cfg(config, 'temperature', 0.2)

    def forward(self, perturbed=False, noises=None):
        if self.fused:
            return F_.simgcl_propagate(self._graph(), self.user_embedding.weight, self.item_embedding.weight,
                                       self.n_layers, self.eps, perturbed=perturbed, noises=noises)
        all_embs = self.get_ego_embeddings()
        embeddings_list = []
        for layer_idx in range(self.n_layers):
            all_embs = self.gcn_conv(all_embs, self.edge_index, self.edge_weight)
            if perturbed:
                random_noise = torch.rand_like(all_embs) if noises is None else noises[layer_idx]
                all_embs = all_embs + torch.sign(all_embs) * F.normalize(random_noise, dim=-1) * self.eps
            embeddings_list.append(all_embs)
        out = torch.mean(torch.stack(embeddings_list, dim=1), dim=1)
        return torch.split(out, [self.n_users, self.n_items])

    def calculate_cl_loss(self, x1, x2):
        x1, x2 = F.normalize(x1, dim=-1), F.normalize(x2, dim=-1)
        pos_score = torch.exp((x1 * x2).sum(dim=-1) / self.temperature)
        ttl_score = torch.exp(torch.matmul(x1, x2.transpose(0, 1)) / self.temperature).sum(dim=1)
        return -torch.log(pos_score / ttl_score).sum()

    def calculate_loss(self, interaction, noises1=None, noises2=None):
        """simgcl.py:48-60.  The fused route computes the clean view (``super().calculate_loss`` ->
        ``forward()``) and the two perturbed views in ONE engine call that shares their first layer
        (``functional.simgcl_views``: 1 + 3(L-1) SpMMs instead of 3L, one backward propagation instead of
        three).  ``noises1/2``: per-layer rand_like draws for parity runs; default = in-kernel Philox."""
        self._clear_restore()
        user, pos_item, neg_item = (interaction[k] for k in (self.USER_ID, self.ITEM_ID, self.NEG_ITEM_ID))
        if self.fused:
            (ua, ia), (u1, i1), (u2, i2) = F_.simgcl_views(
                self._graph(), self.user_embedding.weight, self.item_embedding.weight, self.n_layers, self.eps,
                noises1=noises1, noises2=noises2)
        else:
            ua, ia = self.forward()
            u1, i1 = self.forward(perturbed=True, noises=noises1)
            u2, i2 = self.forward(perturbed=True, noises=noises2)
        u, pos, neg = ua[user], ia[pos_item], ia[neg_item]
        mf_loss = self.mf_loss(torch.mul(u, pos).sum(dim=1), torch.mul(u, neg).sum(dim=1))
        reg_loss = self.reg_loss(self.user_embedding(user), self.item_embedding(pos_item),
                                 self.item_embedding(neg_item), require_pow=self.require_pow)
        loss = mf_loss + self.reg_weight * reg_loss
        cl_user, cl_item = torch.unique(user), torch.unique(pos_item)
        user_cl_loss = self.calculate_cl_loss(u1[cl_user], u2[cl_user])
        item_cl_loss = self.calculate_cl_loss(i1[cl_item], i2[cl_item])
        return loss + self.cl_rate * (user_cl_loss + item_cl_loss)


class NGCF(_PropagationModel):
    def __init__(self, config, dataset):
        super(NGCF, self).__init__(config, dataset)
        self.embedding_size = config['embedding_size']
        self.hidden_size_list = [self.embedding_size] + list(config['hidden_size_list'])
        self.node_dropout = _cfg(config, 'node_dropout', 0.0)
        self.message_dropout = _cfg(config, 'message_dropout', 0.1)
        self.reg_weight = _cfg(config, 'reg_weight', 1e-5)
        self.user_embedding = nn.Embedding(self.n_users, self.embedding_size)
        self.item_embedding = nn.Embedding(self.n_items, self.embedding_size)
        self.GNNlayers = torch.nn.ModuleList()
        for input_size, output_size in zip(self.hidden_size_list[:-1], self.hidden_size_list[1:]):
            self.GNNlayers.append(BiGNNConv(input_size, output_size))
        self.mf_loss = BPRLoss()
        self.reg_loss = EmbLoss()
        self.restore_user_e = None
        self.restore_item_e = None
        self.apply(_xavier_normal)
        self.other_parameter_name = ['restore_user_e', 'restore_item_e']

    def _graph(self, keep_edges=None) -> GraphHandle:
        g = super()._graph()
        if keep_edges is not None:
            return g.masked(keep_edges)
        if self.node_dropout != 0 and self.training:
            # dropout_adj(p, training=True): Bernoulli(1-p) edge mask, no rescale (ngcf.py:74-90); the
            # CSR is masked in place of the reference's COO round trip + SparseTensor rebuild
            keep = torch.rand(g.nnz(), device=g.device) >= self.node_dropout
            g = g.masked(keep)
        return g

    def forward(self, keep_masks=None, keep_edges=None):
        """``keep_masks`` (per-layer ``[N, d_out]`` bool) and ``keep_edges`` (one flag per CSR entry) replace the
        draws of ``nn.Dropout`` (ngcf.py:97) and ``dropout_adj`` (ngcf.py:81,89) in parity runs."""
        g = self._graph(keep_edges)
        fusable = all(d % 4 == 0 and d <= 256 for d in self.hidden_size_list)
        if not torch.is_grad_enabled() and fusable:
            # inference: SpMM + fused tail writing into the concat buffer.  nn.Dropout at ngcf.py:97 is a
            # fresh module (always training): its mask is drawn here and handed to the tail kernel.
            if keep_masks is None and self.message_dropout > 0:
                n = self.n_users + self.n_items
                keep_masks = [torch.rand(n, d, device=g.device) >= self.message_dropout
                              for d in self.hidden_size_list[1:]]
            weights = [(m.lin1.weight, m.lin1.bias, m.lin2.weight, m.lin2.bias) for m in self.GNNlayers]
            return F_.ngcf_forward(g, self.user_embedding.weight, self.item_embedding.weight, weights,
                                   message_dropout=self.message_dropout, keep_masks=keep_masks)
        all_embeddings = self.get_ego_embeddings()
        embeddings_list = [all_embeddings]
        for l, gnn in enumerate(self.GNNlayers):
            keep = None
            if keep_masks is not None:
                keep = keep_masks[l]
            elif self.message_dropout > 0:
                keep = torch.rand(all_embeddings.size(0), gnn.out_channels, device=g.device) >= self.message_dropout
            if fusable:
                # training: SpMM (own backward) + fused tail with its own backward kernels (BiGNNConv tail,
                # LeakyReLU, the always-on Dropout of ngcf.py:97 with the mask above, L2-normalise)
                x_prop = F_.spmm(g, all_embeddings)
                all_embeddings = F_.bignn_tail_autograd(
                    x_prop, all_embeddings, gnn.lin1.weight, gnn.lin1.bias, gnn.lin2.weight, gnn.lin2.bias,
                    slope=0.2, keep=keep, drop_p=self.message_dropout if keep is not None else 0.0, normalize=True)
            else:
                all_embeddings = gnn(all_embeddings, g, None)
                all_embeddings = nn.LeakyReLU(negative_slope=0.2)(all_embeddings)
                if keep is not None:
                    all_embeddings = all_embeddings * keep / (1.0 - self.message_dropout)
                all_embeddings = F.normalize(all_embeddings, p=2, dim=1)
            embeddings_list += [all_embeddings]
        out = torch.cat(embeddings_list, dim=1)
        return torch.split(out, [self.n_users, self.n_items])

    def calculate_loss(self, interaction):
        self._clear_restore()
        _, (u, pos, neg), mf_loss = self._bpr_rows(interaction)
        return mf_loss + self.reg_weight * self.reg_loss(u, pos, neg)      # ngcf.py:121: the PROPAGATED rows
