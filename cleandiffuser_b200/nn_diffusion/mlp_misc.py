"""``DVInvMlp``, ``SfBCUNet`` and ``PearceMlp``: the remaining MLP-class action denoisers of the reference.

* ``DVInvMlp`` (cleandiffuser/nn_diffusion/dvinvmlp.py:9-47): ``cat[x, time_mlp(map_noise(t)), cond] -> 3 x (Linear + Mish)
  -> Linear(act_dim)`` -- DQLMlp's graph with a configurable width and a mandatory condition (two stacked observations).
* ``SfBCUNet`` (cleandiffuser/nn_diffusion/sfbc_unet.py:9-82): a U-shaped stack of Linear residual blocks
  ``silu(W2 (silu(W1 x) + Wc c)) + skip(x)`` with ``c = t_layer(map_noise(t)) + condition``; the up path concatenates the
  down path's activations.

* ``PearceMlp`` (cleandiffuser/nn_diffusion/pearcemlp.py:35-79): the Diffusion-BC MLP -- ``act_emb`` (Linear, LeakyReLU, Linear)
  of the noisy action, three ``FCBlock``s (Linear -> GroupNorm1d(8 groups) -> exact GELU) whose inputs re-concatenate the raw
  action and the raw time, residual connections scaled by 1/1.414, a Linear head.

State-dict keys equal the reference's.
"""
from typing import List, Optional

import torch
import torch.nn as nn

from .base import BaseNNDiffusion
from ..utils import GroupNorm1d


class DVInvMlp(BaseNNDiffusion):
    def __init__(self, obs_dim: int, act_dim: int, emb_dim: int = 16, hidden_dim: int = 256,
                 timestep_emb_type: str = "positional", timestep_emb_params: Optional[dict] = None):
        super().__init__(emb_dim, timestep_emb_type, timestep_emb_params)
        self.time_mlp = nn.Sequential(nn.Linear(emb_dim, emb_dim * 2), nn.Mish(), nn.Linear(emb_dim * 2, emb_dim))
        layers, width = [], obs_dim * 2 + act_dim + emb_dim
        for _ in range(3):
            layers += [nn.Linear(width, hidden_dim), nn.Mish()]
            width = hidden_dim
        self.mid_layer = nn.Sequential(*layers)
        self.final_layer = nn.Linear(hidden_dim, act_dim)

    def forward(self, x: torch.Tensor, noise: torch.Tensor, condition: torch.Tensor = None):
        """x (b, act_dim), noise (b,), condition (b, 2 * obs_dim) -> (b, act_dim).  A missing condition raises (``torch.cat``
        of ``None``), as in the reference."""
        temb = self.time_mlp(self.map_noise(noise))
        return self.final_layer(self.mid_layer(torch.cat([x, temb, condition], -1)))


class SfBCResidualBlock(nn.Module):
    def __init__(self, in_dim: int, out_dim: int, emb_dim: int):
        super().__init__()
        self.linear1 = nn.Sequential(nn.Linear(in_dim, out_dim), nn.SiLU())
        self.linear2 = nn.Sequential(nn.Linear(out_dim, out_dim), nn.SiLU())
        self.linearc = nn.Linear(emb_dim, out_dim)
        self.skip = nn.Linear(in_dim, out_dim) if in_dim != out_dim else nn.Identity()

    def forward(self, x: torch.Tensor, c: torch.Tensor):
        return self.linear2(self.linear1(x) + self.linearc(c)) + self.skip(x)


class SfBCUNet(BaseNNDiffusion):
    def __init__(self, act_dim: int, emb_dim: int = 64, hidden_dims: List[int] = (512, 256, 128),
                 timestep_emb_type: str = "untrainable_fourier", timestep_emb_params: Optional[dict] = None):
        super().__init__(emb_dim, timestep_emb_type, timestep_emb_params)
        n = len(hidden_dims)
        self.t_layer = nn.Sequential(nn.Linear(emb_dim, emb_dim), nn.SiLU(), nn.Linear(emb_dim, emb_dim))
        self.down_blocks, self.up_blocks = nn.ModuleList(), nn.ModuleList()
        width = act_dim
        for i in range(n):
            self.down_blocks.append(SfBCResidualBlock(width, hidden_dims[i], emb_dim))
            width = hidden_dims[i]
        self.mid_block = SfBCResidualBlock(width, width, emb_dim)
        for i in range(n - 1):
            self.up_blocks.append(SfBCResidualBlock(width + hidden_dims[-1 - i], hidden_dims[-2 - i], emb_dim))
            width = hidden_dims[-2 - i]
        self.out_layer = nn.Linear(width, act_dim)

    def forward(self, x: torch.Tensor, noise: torch.Tensor, condition: Optional[torch.Tensor] = None):
        """x (b, act_dim) or (b, horizon, act_dim), noise (b,), condition (b, emb_dim)|None -> like x."""
        c = self.t_layer(self.map_noise(noise))
        if condition is not None:
            c = c + condition
        kept = []
        for block in self.down_blocks:
            x = block(x, c)
            kept.append(x)
        x = self.mid_block(x, c)
        for block in self.up_blocks:
            x = block(torch.cat([x, kept.pop()], dim=-1), c)
        return self.out_layer(x)


class FCBlock(nn.Module):
    def __init__(self, in_feats: int, out_feats: int):
        super().__init__()
        self.model = nn.Sequential(nn.Linear(in_feats, out_feats), GroupNorm1d(out_feats, 8, 4), nn.GELU())

    def forward(self, x):
        return self.model(x)


class PearceMlp(BaseNNDiffusion):
    def __init__(self, act_dim: int, To: int = 1, timestep_emb_type: str = "positional", emb_dim: int = 128, hidden_dim: int = 512,
                 timestep_emb_params: Optional[dict] = None):
        super().__init__(emb_dim, timestep_emb_type, timestep_emb_params)
        self.act_emb = nn.Sequential(nn.Linear(act_dim, emb_dim), nn.LeakyReLU(), nn.Linear(emb_dim, emb_dim))
        self.fcs = nn.ModuleList([FCBlock(emb_dim * (2 + To), hidden_dim), FCBlock(hidden_dim + act_dim + 1, hidden_dim),
                                  FCBlock(hidden_dim + act_dim + 1, hidden_dim), nn.Linear(hidden_dim + act_dim + 1, act_dim)])
        self.To, self.emb_dim = To, emb_dim

    def forward(self, x: torch.Tensor, noise: torch.Tensor, condition: Optional[torch.Tensor] = None):
        """x (b, act_dim), noise (b,), condition (b, To, emb_dim)|None -> (b, act_dim)."""
        x_e, t_e = self.act_emb(x), self.map_noise(noise)
        t = noise.unsqueeze(-1)
        if condition is None:
            condition = torch.zeros(x.shape[0], self.To, self.emb_dim).to(x.device)
        nn1 = self.fcs[0](torch.cat([x_e, t_e, torch.flatten(condition, 1)], -1))
        nn2 = self.fcs[1](torch.cat([nn1 / 1.414, x, t], -1)) + nn1 / 1.414
        nn3 = self.fcs[2](torch.cat([nn2 / 1.414, x, t], -1)) + nn2 / 1.414
        return self.fcs[3](torch.cat([nn3, x, t], -1))
