"""``DiT1d``: adaLN-Zero transformer denoiser over the horizon axis.

Reference: cleandiffuser/nn_diffusion/dit.py:10-132 (``modulate``, ``DiTBlock``,
``FinalLayer1d``, ``DiT1d``).  Same parameter names (``blocks.<i>.attn.in_proj_weight`` ...)
so Decision-Diffuser / DP checkpoints load.  The engine lowers a block to
h = LN+modulate(x) -> QKV GEMM -> per-(trajectory, head) softmax(QK^T)V -> out-proj GEMM with
gated residual onto h (sic) -> LN+modulate -> MLP GEMMs (GELU-tanh) with gated residual.
"""
from typing import Optional

import torch
import torch.nn as nn

from .base import BaseNNDiffusion
from ..utils import SinusoidalEmbedding


def modulate(x, shift, scale):
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


class DiTBlock(nn.Module):
    def __init__(self, hidden_size: int, n_heads: int, dropout: float = 0.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.attn = nn.MultiheadAttention(hidden_size, n_heads, dropout, batch_first=True)
        self.norm2 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.mlp = nn.Sequential(
            nn.Linear(hidden_size, hidden_size * 4), nn.GELU(approximate="tanh"), nn.Dropout(dropout),
            nn.Linear(hidden_size * 4, hidden_size))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, hidden_size * 6))

    def forward(self, x: torch.Tensor, t: torch.Tensor):
        s_att, k_att, g_att, s_mlp, k_mlp, g_mlp = self.adaLN_modulation(t).chunk(6, dim=1)
        # Load-bearing reference quirk (dit.py:33-34): the attention residual is taken around the
        # *modulated, normalised* tokens, not around the block input.
        x = modulate(self.norm1(x), s_att, k_att)
        x = x + g_att.unsqueeze(1) * self.attn(x, x, x)[0]
        return x + g_mlp.unsqueeze(1) * self.mlp(modulate(self.norm2(x), s_mlp, k_mlp))


class FinalLayer1d(nn.Module):
    def __init__(self, hidden_size: int, out_dim: int):
        super().__init__()
        self.norm_final = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.linear = nn.Linear(hidden_size, out_dim)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 2 * hidden_size))

    def forward(self, x: torch.Tensor, t: torch.Tensor):
        shift, scale = self.adaLN_modulation(t).chunk(2, dim=1)
        return self.linear(modulate(self.norm_final(x), shift, scale))


class DiT1d(BaseNNDiffusion):
    def __init__(self, in_dim: int, emb_dim: int, d_model: int = 384, n_heads: int = 6, depth: int = 12,
                 dropout: float = 0.0, timestep_emb_type: str = "positional",
                 timestep_emb_params: Optional[dict] = None):
        super().__init__(emb_dim, timestep_emb_type, timestep_emb_params)
        self.in_dim, self.emb_dim, self.d_model = in_dim, emb_dim, d_model
        self.x_proj = nn.Linear(in_dim, d_model)
        self.map_emb = nn.Sequential(nn.Linear(emb_dim, d_model), nn.Mish(), nn.Linear(d_model, d_model), nn.Mish())
        self.pos_emb = SinusoidalEmbedding(d_model)
        self.pos_emb_cache = None
        self.blocks = nn.ModuleList([DiTBlock(d_model, n_heads, dropout) for _ in range(depth)])
        self.final_layer = FinalLayer1d(d_model, in_dim)
        self.initialize_weights()

    def initialize_weights(self):
        """Xavier for Linear, N(0,.02) for the time MLP, zeros for every adaLN and the output head (dit.py:78-104)."""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        for idx in (0, 2):
            nn.init.normal_(self.map_emb[idx].weight, std=0.02)
        zeroed = [blk.adaLN_modulation[-1] for blk in self.blocks]
        zeroed += [self.final_layer.adaLN_modulation[-1], self.final_layer.linear]
        for lin in zeroed:
            nn.init.constant_(lin.weight, 0)
            nn.init.constant_(lin.bias, 0)

    def forward(self, x: torch.Tensor, noise: torch.Tensor, condition: Optional[torch.Tensor] = None):
        """x (b, horizon, in_dim), noise (b,), condition (b, emb_dim)|None -> (b, horizon, in_dim)."""
        if self.pos_emb_cache is None or self.pos_emb_cache.shape[0] != x.shape[1]:
            # int64 positions: the table degenerates to 4 distinct columns -- keep.
            self.pos_emb_cache = self.pos_emb(torch.arange(x.shape[1], device=x.device))
        x = self.x_proj(x) + self.pos_emb_cache[None, ]
        emb = self.map_noise(noise)
        emb = emb + (condition if condition is not None else torch.zeros_like(emb))
        emb = self.map_emb(emb)
        for blk in self.blocks:
            x = blk(x, emb)
        return self.final_layer(x, emb)
