"""Forward bodies of the three models on the benchmarked path, with the reference's structure and
attribute names: ``LightGCN`` (general_recommender/lightgcn.py:36-81), ``NGCF`` (ngcf.py:33-104) and
``SimGCL`` (simgcl.py:16-38).  Losses / predict code are mini-batch dense torch ops in the reference
(out of scope) — only a plain BPR ``calculate_loss`` is kept so that a training step can be
driven end to end.

Each ``forward`` has two routes that give the same numbers:
* the reference's layer-by-layer loop over the drop-in conv layers (``fused=False``), and
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


class LightGCN(GeneralGraphRecommender):
    def __init__(self, config, dataset):
        super(LightGCN, self).__init__(config, dataset)
        self.latent_dim = config['embedding_size']
        self.n_layers = config['n_layers']
        self.reg_weight = _cfg(config, 'reg_weight', 1e-5)
        self.fused = _cfg(config, 'fused_propagation', True)
        self.user_embedding = torch.nn.Embedding(num_embeddings=self.n_users, embedding_dim=self.latent_dim)
        self.item_embedding = torch.nn.Embedding(num_embeddings=self.n_items, embedding_dim=self.latent_dim)
        self.gcn_conv = LightGCNConv(dim=self.latent_dim)
        self.restore_user_e = None
        self.restore_item_e = None
        self.apply(_xavier_uniform)
        self.other_parameter_name = ['restore_user_e', 'restore_item_e']

    def _graph(self) -> GraphHandle:
        n = self.n_users + self.n_items
        return _resolve(self.edge_index, self.edge_weight, n, n)

    def get_ego_embeddings(self):
        return torch.cat([self.user_embedding.weight, self.item_embedding.weight], dim=0)

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
        if self.restore_user_e is not None or self.restore_item_e is not None:
            self.restore_user_e, self.restore_item_e = None, None
        user, pos_item, neg_item = interaction[self.USER_ID], interaction[self.ITEM_ID], interaction[self.NEG_ITEM_ID]
        user_all, item_all = self.forward()
        u, pos, neg = user_all[user], item_all[pos_item], item_all[neg_item]
        pos_scores, neg_scores = (u * pos).sum(dim=1), (u * neg).sum(dim=1)
        mf_loss = -torch.log(1e-10 + torch.sigmoid(pos_scores - neg_scores)).mean()       # BPRLoss
        reg = sum(e.norm(2).pow(2) for e in (self.user_embedding(user), self.item_embedding(pos_item),
                                              self.item_embedding(neg_item))) / user.numel()
        return mf_loss + self.reg_weight * reg

    def full_sort_predict(self, interaction):
        user = interaction[self.USER_ID]
        if self.restore_user_e is None or self.restore_item_e is None:
            self.restore_user_e, self.restore_item_e = self.forward()
        return torch.matmul(self.restore_user_e[user], self.restore_item_e.transpose(0, 1)).view(-1)


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


class NGCF(GeneralGraphRecommender):
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
        self.restore_user_e = None
        self.restore_item_e = None
        self.apply(_xavier_normal)
        self.other_parameter_name = ['restore_user_e', 'restore_item_e']

    def get_ego_embeddings(self):
        return torch.cat([self.user_embedding.weight, self.item_embedding.weight], dim=0)

    def _graph(self) -> GraphHandle:
        n = self.n_users + self.n_items
        g = _resolve(self.edge_index, self.edge_weight, n, n)
        if self.node_dropout != 0 and self.training:
            # dropout_adj(p, training=True): Bernoulli(1-p) edge mask, no rescale (ngcf.py:74-90); the
            # CSR is masked in place of the reference's COO round trip + SparseTensor rebuild
            keep = torch.rand(g.nnz(), device=g.device) >= self.node_dropout
            g = g.masked(keep)
        return g

    def forward(self, keep_masks=None):
        g = self._graph()
        if not torch.is_grad_enabled():
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
        fusable = all(d % 4 == 0 and d <= 256 for d in self.hidden_size_list)
        for l, gnn in enumerate(self.GNNlayers):
            if fusable:
                # training: SpMM (own backward) + fused tail with autograd (BiGNNConv tail, LeakyReLU, the
                # always-on Dropout of ngcf.py:97 with a torch-drawn mask, L2-normalise) in one kernel
                x_prop = F_.spmm(g, all_embeddings)
                keep = None
                if keep_masks is not None:
                    keep = keep_masks[l]
                elif self.message_dropout > 0:
                    keep = torch.rand(all_embeddings.size(0), gnn.out_channels, device=g.device) >= self.message_dropout
                all_embeddings = F_.bignn_tail_autograd(
                    x_prop, all_embeddings, gnn.lin1.weight, gnn.lin1.bias, gnn.lin2.weight, gnn.lin2.bias,
                    slope=0.2, keep=keep, drop_p=self.message_dropout if keep is not None else 0.0, normalize=True)
            else:
                all_embeddings = gnn(all_embeddings, g, None)
                all_embeddings = nn.LeakyReLU(negative_slope=0.2)(all_embeddings)
                all_embeddings = nn.Dropout(self.message_dropout)(all_embeddings)
                all_embeddings = F.normalize(all_embeddings, p=2, dim=1)
            embeddings_list += [all_embeddings]
        out = torch.cat(embeddings_list, dim=1)
        return torch.split(out, [self.n_users, self.n_items])
