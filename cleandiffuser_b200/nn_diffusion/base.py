"""``nn_diffusion`` plugin base: ``forward(x, noise, condition) -> like x``.

Reference: cleandiffuser/nn_diffusion/base_nn_diffusion.py:9-43.  Any user
subclass is legal; only the four backbones the engine knows are lowered to the
CUDA program, everything else runs through ``forward`` in PyTorch.
"""
from typing import Optional

import torch
import torch.nn as nn

from ..utils import SUPPORTED_TIMESTEP_EMBEDDING


class BaseNNDiffusion(nn.Module):
    def __init__(self, emb_dim: int, timestep_emb_type: str = "positional",
                 timestep_emb_params: Optional[dict] = None):
        assert timestep_emb_type in SUPPORTED_TIMESTEP_EMBEDDING.keys()
        super().__init__()
        self.map_noise = SUPPORTED_TIMESTEP_EMBEDDING[timestep_emb_type](emb_dim, **(timestep_emb_params or {}))

    def forward(self, x: torch.Tensor, noise: torch.Tensor, condition: Optional[torch.Tensor] = None):
        raise NotImplementedError
