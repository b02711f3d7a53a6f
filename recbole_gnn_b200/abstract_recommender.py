"""``GeneralGraphRecommender`` (recbole_gnn/model/abstract_recommender.py:7-20): owns the graph state
``edge_index`` / ``edge_weight`` / ``use_sparse`` of every general graph model.

With RecBole installed this subclasses ``recbole.model.abstract_recommender.GeneralRecommender`` as the
reference does; without it, a minimal ``nn.Module`` base provides the attributes the hot path needs
(``n_users``, ``n_items``, ``device``).
"""
from __future__ import annotations

import torch
import torch.nn as nn

try:  # pragma: no cover
    from recbole.model.abstract_recommender import GeneralRecommender as _Base
    HAVE_RECBOLE = True
except Exception:
    HAVE_RECBOLE = False

    class _Base(nn.Module):
        """Stand-in for RecBole's GeneralRecommender: ids, sizes, device."""

        def __init__(self, config, dataset):
            super().__init__()
            self.USER_ID, self.ITEM_ID = "user_id", "item_id"
            self.NEG_ITEM_ID = "neg_item_id"
            self.n_users = dataset.num(dataset.uid_field)
            self.n_items = dataset.num(dataset.iid_field)
            self.device = torch.device(config["device"])


def _cfg(config, key, default=None):
    try:
        v = config[key]
    except Exception:
        return default
    return default if v is None else v


class GeneralGraphRecommender(_Base):
    def __init__(self, config, dataset):
        super(GeneralGraphRecommender, self).__init__(config, dataset)
        enable_sparse = _cfg(config, "enable_sparse")
        if enable_sparse not in (True, False, None):      # quick_start.py:21-24
            raise ValueError("Your config `enable_sparse` must be `True` or `False` or `None`")
        self.edge_index, self.edge_weight = dataset.get_norm_adj_mat(enable_sparse=enable_sparse)
        self.use_sparse = bool(enable_sparse) and dataset.is_sparse
        if self.use_sparse:
            self.edge_index, self.edge_weight = self.edge_index.to(self.device), None
        else:
            self.edge_index, self.edge_weight = self.edge_index.to(self.device), self.edge_weight.to(self.device)
