"""``DQLMlp``: the Diffusion-QL action denoiser (cleandiffuser/nn_diffusion/dqlmlp.py:9-49).

``cat[x, time_mlp(map_noise(t)), obs] -> 3 x (Linear 256 + Mish) -> Linear(act_dim)``.
"""
from typing import Optional

import torch
import torch.nn as nn

from .base import BaseNNDiffusion


class DQLMlp(BaseNNDiffusion):
    def __init__(self, obs_dim: int, act_dim: int, emb_dim: int = 16,
                 timestep_emb_type: str = "positional", timestep_emb_params: Optional[dict] = None):
        super().__init__(emb_dim, timestep_emb_type, timestep_emb_params)
        self.obs_dim = obs_dim
        self.time_mlp = nn.Sequential(nn.Linear(emb_dim, emb_dim * 2), nn.Mish(), nn.Linear(emb_dim * 2, emb_dim))
        hidden, layers, width = 256, [], obs_dim + act_dim + emb_dim
        for _ in range(3):
            layers += [nn.Linear(width, hidden), nn.Mish()]
            width = hidden
        self.mid_layer = nn.Sequential(*layers)
        self.final_layer = nn.Linear(hidden, act_dim)

    def forward(self, x: torch.Tensor, noise: torch.Tensor, condition: Optional[torch.Tensor] = None):
        """x (b, act_dim), noise (b,), condition (b, obs_dim)|None -> (b, act_dim)."""
        if condition is None:
            condition = torch.zeros(x.shape[0], self.obs_dim).to(x.device)
        temb = self.time_mlp(self.map_noise(noise))
        return self.final_layer(self.mid_layer(torch.cat([x, temb, condition], -1)))
