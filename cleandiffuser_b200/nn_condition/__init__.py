"""Condition embedders (the ``nn_condition`` plugin surface).

``forward(condition, mask=None) -> (b, *cond_out_shape)``.  In train mode a
Bernoulli label-dropout mask is drawn per row; in eval mode ``mask=None`` means
"keep everything".  Reference: cleandiffuser/nn_condition/base_nn_condition.py:7-57
and cleandiffuser/nn_condition/mlp.py:9-92.  These run ONCE per ``sample()`` call,
outside the reverse loop, so they stay plain PyTorch modules.
"""
from typing import List

import torch
import torch.nn as nn

from ..utils import at_least_ndim, Mlp


def get_mask(mask, mask_shape: tuple, dropout: float, train: bool, device):
    if train:
        return (torch.rand(mask_shape, device=device) > dropout).float()
    return 1. if mask is None else mask


class BaseNNCondition(nn.Module):
    def forward(self, condition: torch.Tensor, mask: torch.Tensor = None):
        raise NotImplementedError


class IdentityCondition(BaseNNCondition):
    """Pass the condition through (times the dropout / user mask)."""

    def __init__(self, dropout: float = 0.25):
        super().__init__()
        self.dropout = dropout

    def _row_mask(self, condition, mask, ndim=None):
        m = get_mask(mask, (condition.shape[0],), self.dropout, self.training, condition.device)
        return at_least_ndim(m, condition.dim() if ndim is None else ndim)

    def _embed(self, condition):
        return condition

    def forward(self, condition: torch.Tensor, mask: torch.Tensor = None):
        return self._embed(condition) * self._row_mask(condition, mask)


class LinearCondition(IdentityCondition):
    def __init__(self, in_dim: int, out_dim: int, dropout: float = 0.25):
        super().__init__(dropout)
        self.affine = nn.Linear(in_dim, out_dim)

    def _embed(self, condition):
        return self.affine(condition)


class MLPCondition(IdentityCondition):
    def __init__(self, in_dim: int, out_dim: int, hidden_dims: List[int],
                 act=nn.LeakyReLU(), dropout: float = 0.25):
        super().__init__(dropout)
        hidden_dims = [hidden_dims] if isinstance(hidden_dims, int) else hidden_dims
        self.mlp = Mlp(in_dim, hidden_dims, out_dim, act)

    def _embed(self, condition):
        return self.mlp(condition)


class MLPSieveObsCondition(IdentityCondition):
    def __init__(self, o_dim: int, emb_dim: int = 128, hidden_dim: int = 512, dropout: float = 0.25):
        super().__init__(dropout)
        self.mlp = Mlp(o_dim, [hidden_dim], emb_dim, nn.LeakyReLU())

    def forward(self, obs: torch.Tensor, mask: torch.Tensor = None):
        return torch.flatten(self.mlp(obs), 1) * self._row_mask(obs, mask, ndim=2)
