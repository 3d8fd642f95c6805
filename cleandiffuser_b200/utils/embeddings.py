"""Timestep / position embeddings used by the denoiser backbones.

Inside ``sample()`` the diffusion time is the same for every row of the batch
(diffusionsde.py:528 / :874), so the CUDA engine never evaluates these on the
device per element: the host evaluates the *bound* ``map_noise`` module on one
row per step and ships the rows as a table.  That also reproduces, for free,
the reference quirk that an int64 ``t`` degenerates the positional embedding to
``[cos t, 1, ..., sin t, 0, ...]`` (the frequency vector is cast to int64,
utils/utils.py:261), which trained checkpoints depend on.

Reference: cleandiffuser/utils/utils.py:248-336.
"""
import math

import numpy as np
import torch
import torch.nn as nn


def _edm_freqs(dim: int, max_positions: int, endpoint: bool, device):
    half = dim // 2
    ramp = torch.arange(start=0, end=half, dtype=torch.float32, device=device)
    ramp = ramp / (half - (1 if endpoint else 0))
    return (1 / max_positions) ** ramp


class PositionalEmbedding(nn.Module):
    """EDM/DDPM++ positional embedding: cat[cos(t f), sin(t f)] (utils/utils.py:248-263)."""

    def __init__(self, dim: int, max_positions: int = 10000, endpoint: bool = False):
        super().__init__()
        self.dim, self.max_positions, self.endpoint = dim, max_positions, endpoint

    def forward(self, x):
        freqs = _edm_freqs(self.dim, self.max_positions, self.endpoint, x.device)
        # the cast to x.dtype is what makes int64 timesteps degenerate -- keep it.
        phase = x.ger(freqs.to(x.dtype))
        return torch.cat([phase.cos(), phase.sin()], dim=1)


class UntrainablePositionalEmbedding(nn.Module):
    """Same table for inputs of any leading shape (utils/utils.py:266-281)."""

    def __init__(self, dim: int, max_positions: int = 10000, endpoint: bool = False):
        super().__init__()
        self.dim, self.max_positions, self.endpoint = dim, max_positions, endpoint

    def forward(self, x):
        freqs = _edm_freqs(self.dim, self.max_positions, self.endpoint, x.device)
        phase = torch.einsum('...i,j->...ij', x, freqs.to(x.dtype))
        return torch.cat([phase.cos(), phase.sin()], dim=1)


class SinusoidalEmbedding(nn.Module):
    """Transformer-style cat[sin, cos] embedding (utils/utils.py:286-300)."""

    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        half = self.dim // 2
        rate = math.log(10000) / (half - 1)
        freqs = torch.exp(torch.arange(half, device=x.device) * -rate)
        phase = torch.einsum('...i,j->...ij', x, freqs.to(x.dtype))
        return torch.cat((phase.sin(), phase.cos()), dim=-1)


class FourierEmbedding(nn.Module):
    """Random Fourier features + 2-layer MLP (utils/utils.py:305-318)."""

    def __init__(self, dim: int, scale=16):
        super().__init__()
        self.freqs = nn.Parameter(torch.randn(dim // 8) * scale, requires_grad=False)
        self.mlp = nn.Sequential(nn.Linear(dim // 4, dim), nn.Mish(), nn.Linear(dim, dim))

    def forward(self, x: torch.Tensor):
        phase = torch.einsum('...i,j->...ij', x, (2 * np.pi * self.freqs).to(x.dtype))
        return self.mlp(torch.cat([phase.cos(), phase.sin()], -1))


class UntrainableFourierEmbedding(nn.Module):
    """Random Fourier features without the MLP (utils/utils.py:321-330)."""

    def __init__(self, dim: int, scale=16):
        super().__init__()
        self.freqs = nn.Parameter(torch.randn(dim // 2) * scale, requires_grad=False)

    def forward(self, x: torch.Tensor):
        phase = torch.einsum('...i,j->...ij', x, (2 * np.pi * self.freqs).to(x.dtype))
        return torch.cat([phase.cos(), phase.sin()], -1)


SUPPORTED_TIMESTEP_EMBEDDING = {
    "positional": PositionalEmbedding,
    "fourier": FourierEmbedding,
    "untrainable_fourier": UntrainableFourierEmbedding,
    "untrainable_positional": UntrainablePositionalEmbedding,
}
