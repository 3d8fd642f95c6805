"""Small nn building blocks shared by the backbones and condition embedders.

Reference: cleandiffuser/utils/building_blocks.py:13-76 (``Mlp``, ``GroupNorm1d``)
and cleandiffuser/utils/utils.py:21-72 (``at_least_ndim``, ``to_tensor``).
Parameter names/shapes equal the reference's so checkpoints interchange.
"""
from typing import List, Union

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def at_least_ndim(x, ndim: int, pad: int = 0):
    """Right- (pad=0) or left- (pad=1) pad the shape of ``x`` with 1s up to ``ndim`` dims."""
    if isinstance(x, (int, float)):
        return x
    if not isinstance(x, (np.ndarray, torch.Tensor)):
        raise ValueError(f"Unsupported type {type(x)}")
    missing = ndim - x.ndim
    if missing <= 0:
        return x
    ones = (1,) * missing
    shape = tuple(x.shape) + ones if pad == 0 else ones + tuple(x.shape)
    return x.reshape(shape)


def to_tensor(x, device=None):
    if isinstance(x, torch.Tensor):
        return x.to(device)
    if isinstance(x, (np.ndarray, list, tuple, int, float)):
        return torch.tensor(x, device=device)
    raise ValueError(f"Unsupported type {type(x)}")


def count_parameters(model: nn.Module):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


class Mlp(nn.Module):
    """``in -> hidden_dims... -> out`` perceptron; state-dict keys ``mlp.<i>.0.*`` / ``mlp.<n>.*``."""

    def __init__(self, in_dim: int, hidden_dims: List[int], out_dim: int,
                 activation: nn.Module = nn.ReLU(), out_activation: nn.Module = nn.Identity()):
        super().__init__()
        widths = [in_dim] + list(hidden_dims)
        stages = [nn.Sequential(nn.Linear(a, b), activation) for a, b in zip(widths[:-1], widths[1:])]
        self.mlp = nn.Sequential(*stages, nn.Linear(widths[-1], out_dim), out_activation)

    def forward(self, x):
        return self.mlp(x)


class GroupNorm1d(nn.Module):
    """GroupNorm over ``(b, C, L)`` with ``G = min(num_groups, C // min_channels_per_group)``."""

    def __init__(self, dim, num_groups=32, min_channels_per_group=4, eps=1e-5):
        super().__init__()
        self.num_groups = min(num_groups, dim // min_channels_per_group)
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))

    def forward(self, x):
        y = F.group_norm(x.unsqueeze(2), self.num_groups,
                         self.weight.to(x.dtype), self.bias.to(x.dtype), self.eps)
        return y.squeeze(2)
