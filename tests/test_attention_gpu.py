"""Operator-level parity of CDS_OP_ATTN in the TF32 tensor-core programs (persistent TMA-fed kernel, csrc/attention_tma.cuh)
against softmax(q k^T / sqrt(d)) v in fp64 on the same TF32-rounded q / k / v.

Reference behaviour: nn.MultiheadAttention's core inside DiTBlock (cleandiffuser/nn_diffusion/dit.py:10-36)."""
import ctypes as C

import numpy as np
import pytest
import torch

from cleandiffuser_b200.engine import cabi, lower

pytestmark = pytest.mark.gpu


def _run_attn(qkv, heads, out_dtype):
    B, L, C3 = qkv.shape
    out = torch.full((B, L, C3 // 3), float("nan"), device=qkv.device, dtype=torch.float32)
    op = cabi.Op()
    op.kind = cabi.OP_ATTN
    a = op.u.attn
    a.batch, a.L, a.C, a.heads = B, L, C3 // 3, heads
    a.qkv = qkv.data_ptr()
    a.out = out.data_ptr()
    a.out_dtype, a.qkv_dtype = out_dtype, cabi.TF32
    cabi.run_op(qkv.device.index or 0, op, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("B,L,heads", [(3, 100, 10), (2, 128, 4), (5, 17, 2), (1, 7, 1), (4, 96, 3), (700, 100, 10), (2, 113, 5)])
def test_tf32_attention_matches_fp64(B, L, heads):
    g = torch.Generator().manual_seed(100 * B + L)
    C_ = 32 * heads
    qkv = torch.randn(B, L, 3 * C_, generator=g) * 1.5
    qkv = lower.round_tf32(qkv).cuda()
    out = _run_attn(qkv, heads, cabi.F32)
    q, k, v = (t.reshape(B, L, heads, 32).permute(0, 2, 1, 3).double() for t in qkv.split(C_, dim=-1))
    ref = torch.softmax(q @ k.transpose(-1, -2) / np.sqrt(32.0), dim=-1) @ v
    ref = ref.permute(0, 2, 1, 3).reshape(B, L, C_)
    err = (out.double() - ref).abs()
    assert torch.isfinite(out).all()
    # P is rounded to TF32 (10 mantissa bits) before P V: |err| <= 2^-11 * sum_j P_j |v_j| <= ~5e-4 * max|v|
    assert err.max().item() < 4e-3, err.max().item()
    assert err.mean().item() < 3e-4, err.mean().item()


def test_tf32_attention_rounded_output_and_repeatability():
    g = torch.Generator().manual_seed(7)
    qkv = lower.round_tf32(torch.randn(64, 100, 960, generator=g)).cuda()
    a = _run_attn(qkv, 10, cabi.TF32)
    b = _run_attn(qkv, 10, cabi.TF32)
    assert torch.equal(a, b)
    bits = a.view(torch.int32)
    assert int((bits & 0x1FFF).abs().max()) == 0          # stored values are TF32-representable
    c = _run_attn(qkv, 10, cabi.F32)
    assert (a - c).abs().max().item() < 2e-3
