"""Data-parallel sampling: shard the candidate batch over the ranks of a ``torch.distributed`` group.

The reverse process has no cross-trajectory operation (GroupNorm / LayerNorm / attention are per trajectory), so
the only exchange on the path is ONE all-gather of the finished samples.  Each rank owns a contiguous
chunk of the batch, replicated weights, and its own noise stream (rank r seeds ``seed + r``): N-GPU parity is
defined as the concatenation of the per-shard single-GPU results, never as a reproduction of one device's
full-batch Philox layout.  Backend: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from typing import Optional

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous, balanced split of ``n`` trajectories: the first ``n % world`` ranks get one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _slice(t, lo, hi):
    return t[lo:hi] if isinstance(t, torch.Tensor) else t


def sample_sharded(agent, prior: torch.Tensor, *, group=None, seed: Optional[int] = None, gather: bool = True,
                   condition_cfg=None, mask_cfg=None, warm_start_reference=None, **sample_kwargs):
    """``agent.sample`` on this rank's shard of ``prior`` (+ per-trajectory kwargs), then all-gather.

    Returns ``(x0_full, log)`` with ``x0_full`` of shape ``prior.shape`` on every rank (or the local shard if
    ``gather=False``).  ``n_samples`` is derived from the shard."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = prior.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    if seed is not None:
        torch.manual_seed(seed + rank)
    sample_kwargs = dict(sample_kwargs)
    sample_kwargs["n_samples"] = hi - lo
    x_local, log = agent.sample(prior[lo:hi], condition_cfg=_slice(condition_cfg, lo, hi),
                                mask_cfg=_slice(mask_cfg, lo, hi),
                                warm_start_reference=_slice(warm_start_reference, lo, hi), **sample_kwargs)
    if world == 1 or not gather:
        return x_local, log
    sizes = [shard_bounds(n, world, r) for r in range(world)]
    if all(b - a == sizes[0][1] - sizes[0][0] for a, b in sizes):
        out = torch.empty((n, *x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
        dist.all_gather_into_tensor(out, x_local.contiguous(), group=group)
        return out, log
    # ragged split: pad every shard to the largest one (collectives need equal sizes), gather, trim
    big = max(b - a for a, b in sizes)
    padded = torch.zeros((big, *x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    padded[:hi - lo] = x_local
    out = torch.empty((world * big, *x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * big:r * big + (b - a)] for r, (a, b) in enumerate(sizes)], 0), log
