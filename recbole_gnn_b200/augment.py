"""Per-epoch graph re-sampling on the device: the augmented graphs that SGL
(`general_recommender/sgl.py:82-126`), SEPT (`social_recommender/sept.py:111-133`) and NGCF's node dropout
(`ngcf.py:74-90`, see ``models.NGCF._graph``) rebuild with numpy + PyG on the CPU every epoch / forward.

The reference samples interaction indices with ``np.random.choice`` on the host, rebuilds the int64 COO, normalises
it with ``gcn_norm`` and ships it to the device.  Here the interaction columns stay resident on the GPU:

* ``ND`` (node dropout): a per-entry keep mask is derived from the dropped-node flags on the resident, un-normalised
  CSR, the CSR is compacted in place (``b200gcn_csr_mask`` — no re-sort) and re-normalised (``b200gcn_gcn_norm_csr``);
* ``ED`` / ``RW`` (edge dropout; RW = a fresh ED graph per layer): the kept interactions are re-sorted on the device
  (radix sort, 14 ms for 200 M entries — `b200gcn_csr_from_interactions`) and normalised;
* SEPT's joint graph: interaction edges in both directions plus the DIRECTED social edges, with SEPT's own
  ``1/sqrt(deg_out)`` weights, returned in the dense-edge form the model feeds to ``LightGCNConv``.

Sampling uses torch's CUDA generator (exact-count sampling without replacement, like ``np.random.choice(...,
replace=False)``; the RNG stream is necessarily a different one).  Every function also takes the sampled indices /
flags explicitly, which is how the parity tests compare with the oracle.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import _lib
from .graph import GraphHandle

Tensor = torch.Tensor


def _sample_without_replacement(n: int, k: int, device, generator=None) -> Tensor:
    return torch.randperm(n, device=device, generator=generator)[:k]


class SGLAugmenter:
    """Resident state for SGL-style augmentation of one user-item graph: the interaction id columns and the
    un-normalised symmetric CSR.  ``graph_construction(aug_type, drop_ratio, n_layers)`` returns the two per-layer
    lists of ``(graph, None)`` pairs that ``SGL.forward(graph=...)`` iterates (sgl.py:82-90,136-139)."""

    def __init__(self, uid: Tensor, iid: Tensor, n_users: int, n_items: int, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("SGLAugmenter builds graphs on a CUDA device only (no CPU path exists)")
        self.n_users, self.n_items = int(n_users), int(n_items)
        self.uid, self.iid = uid.to(self.device), iid.to(self.device)
        self._raw: Optional[GraphHandle] = None
        self._row_ids: Optional[Tensor] = None

    def _raw_csr(self) -> GraphHandle:
        if self._raw is None:
            self._raw = GraphHandle.from_interactions(self.uid, self.iid, self.n_users, self.n_items).to(self.device)
            self._row_ids = self._raw.coo()[0]
        return self._raw

    def node_dropout(self, drop_user: Tensor, drop_item: Tensor) -> GraphHandle:
        """sgl.py:95-104 with the dropped ids handed in: interactions touching a dropped user or item vanish."""
        raw = self._raw_csr()
        n = self.n_users + self.n_items
        dead = torch.zeros(n, dtype=torch.bool, device=self.device)
        dead[drop_user.to(self.device)] = True
        dead[drop_item.to(self.device) + self.n_users] = True
        _, col, _ = raw.csr()
        keep = ~(dead[self._row_ids] | dead[col.long()])
        return raw.masked(keep, symmetric=True).gcn_norm()

    def edge_dropout(self, keep_idx: Tensor) -> GraphHandle:
        """sgl.py:106-109 with the kept interaction indices handed in."""
        keep_idx = keep_idx.to(self.device)
        return GraphHandle.from_interactions(self.uid[keep_idx].contiguous(), self.iid[keep_idx].contiguous(),
                                             self.n_users, self.n_items).gcn_norm().to(self.device)

    def random_graph_augment(self, aug_type: str, drop_ratio: float, generator=None) -> Tuple[GraphHandle, None]:
        if aug_type == "ND":
            du = _sample_without_replacement(self.n_users, int(self.n_users * drop_ratio), self.device, generator)
            di = _sample_without_replacement(self.n_items, int(self.n_items * drop_ratio), self.device, generator)
            return self.node_dropout(du, di), None
        if aug_type in ("ED", "RW"):
            E = self.uid.numel()
            keep = _sample_without_replacement(E, int(E * (1 - drop_ratio)), self.device, generator)
            return self.edge_dropout(keep), None
        raise ValueError("aug_type must be 'ND', 'ED' or 'RW'")

    def graph_construction(self, aug_type: str, drop_ratio: float, n_layers: int, generator=None):
        """sgl.py:82-90: ND / ED share one graph over the layers, RW draws one per layer."""
        if aug_type in ("ND", "ED"):
            g1 = [self.random_graph_augment(aug_type, drop_ratio, generator)] * n_layers
            g2 = [self.random_graph_augment(aug_type, drop_ratio, generator)] * n_layers
        elif aug_type == "RW":
            g1 = [self.random_graph_augment(aug_type, drop_ratio, generator) for _ in range(n_layers)]
            g2 = [self.random_graph_augment(aug_type, drop_ratio, generator) for _ in range(n_layers)]
        else:
            raise ValueError("aug_type must be 'ND', 'ED' or 'RW'")
        return g1, g2


def sgl_forward(conv, user_weight: Tensor, item_weight: Tensor, graphs: List[Tuple[GraphHandle, None]]):
    """``SGL.forward(graph)`` (sgl.py:128-145) for an augmented per-layer graph list: layer l propagates over
    ``graphs[l]``; mean over [x_0 .. x_L]."""
    x = torch.cat([user_weight, item_weight])
    outs = [x]
    for g, w in graphs:
        x = conv(x, g, w)
        outs.append(x)
    out = torch.mean(torch.stack(outs, dim=1), dim=1)
    return torch.split(out, [user_weight.size(0), item_weight.size(0)])


def sept_norm_edge_weight(edge_index: Tensor, node_num: int) -> Tensor:
    """``SEPT.get_norm_edge_weight`` (sept.py:81-87): ``deg = degree(edge_index[0])``, zero degrees read as 1,
    ``w = deg^-1/2[row] * deg^-1/2[col]`` — on the device."""
    _lib.require_cuda(edge_index, what="edge_index")
    deg = torch.bincount(edge_index[0], minlength=node_num).to(torch.float32)
    norm = 1.0 / torch.sqrt(torch.where(deg == 0, torch.ones_like(deg), deg))
    return norm[edge_index[0]] * norm[edge_index[1]]


def sept_subgraph_construction(uid: Tensor, iid: Tensor, src_user: Tensor, tgt_user: Tensor, n_users: int,
                               n_items: int, drop_ratio: float, device, keep: Optional[Tensor] = None,
                               net_keep: Optional[Tensor] = None, generator=None) -> Tuple[Tensor, Tensor]:
    """``SEPT.subgraph_construction`` (sept.py:111-133): edge dropout on the interaction graph AND on the social
    graph, concatenation ``[u->i | i->u | social]``, SEPT's normalisation.  Returns ``(edge_index, edge_weight)`` on
    the device, the form ``LightGCNConv`` takes (its CSR is built by the engine's device sort on first use)."""
    dev = torch.device(device)
    uid, iid, src_user, tgt_user = (t.to(dev) for t in (uid, iid, src_user, tgt_user))
    if keep is None:
        keep = _sample_without_replacement(uid.numel(), int(uid.numel() * (1 - drop_ratio)), dev, generator)
    if net_keep is None:
        net_keep = _sample_without_replacement(src_user.numel(), int(src_user.numel() * (1 - drop_ratio)), dev, generator)
    keep, net_keep = keep.to(dev), net_keep.to(dev)
    row, col = uid[keep], iid[keep] + n_users
    edge_index = torch.cat([torch.stack([row, col]), torch.stack([col, row]),
                            torch.stack([src_user[net_keep], tgt_user[net_keep]])], dim=1)
    return edge_index, sept_norm_edge_weight(edge_index, n_users + n_items)
