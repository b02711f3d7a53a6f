"""The two RecBole 1.1.1 loss modules the models on this path construct (``recbole.model.loss.BPRLoss`` /
``EmbLoss``; third-party, call sites lightgcn.py:53-54,100,107, ngcf.py:55-56,119,121), restated so that
``calculate_loss`` of the drop-in models optimises the reference's objective.  Mini-batch-sized dense torch
ops; the fused device versions live in ``functional.bpr_*``.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class BPRLoss(nn.Module):
    def __init__(self, gamma: float = 1e-10):
        super().__init__()
        self.gamma = gamma

    def forward(self, pos_score, neg_score):
        return -torch.log(self.gamma + torch.sigmoid(pos_score - neg_score)).mean()


class EmbLoss(nn.Module):
    def __init__(self, norm: int = 2):
        super().__init__()
        self.norm = norm

    def forward(self, *embeddings, require_pow: bool = False):
        last = embeddings[-1]
        total = torch.zeros(1, dtype=last.dtype, device=last.device)
        for e in embeddings:
            n = torch.norm(e, p=self.norm)
            total = total + (torch.pow(n, self.norm) if require_pow else n)
        total = total / last.shape[0]
        return total / self.norm if require_pow else total
