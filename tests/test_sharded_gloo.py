"""N > 1 path on CPU: world_size-2 gloo processes shard the batch, sample independently and all-gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from common import product_net
from cleandiffuser_b200.diffusion import DiscreteDiffusionSDE
from cleandiffuser_b200.engine.sharded import sample_sharded, shard_bounds


def _agent():
    net, _ = product_net(cases.SAMPLER_NETS["dql_tiny"])
    return DiscreteDiffusionSDE(net, None, diffusion_steps=10, x_max=torch.ones(1, 3), x_min=-torch.ones(1, 3))


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(3)
    prior, cond = torch.zeros(n, 3), torch.randn(n, 4, generator=g)
    x, _ = sample_sharded(_agent(), prior, seed=2, condition_cfg=cond, w_cfg=1.0, solver="ddpm", sample_steps=5)
    torch.save(x, os.path.join(out_dir, f"x{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 7])
def test_two_rank_shard_and_gather(tmp_path, n):
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    got = [torch.load(tmp_path / f"x{r}.pt") for r in range(world)]
    assert torch.equal(got[0], got[1]) and got[0].shape == (n, 3)
    # definition of N-GPU parity: concat over ranks of single-rank samples with seed + rank
    g = torch.Generator().manual_seed(3)
    prior, cond = torch.zeros(n, 3), torch.randn(n, 4, generator=g)
    agent, want = _agent(), []
    for r in range(world):
        lo, hi = shard_bounds(n, world, r)
        torch.manual_seed(2 + r)
        want.append(agent.sample(prior[lo:hi], solver="ddpm", n_samples=hi - lo, sample_steps=5,
                                 condition_cfg=cond[lo:hi], w_cfg=1.0)[0])
    assert torch.allclose(got[0], torch.cat(want, 0), atol=1e-6)


def test_shard_bounds_cover_everything():
    for n in (1, 7, 8, 4096, 16385):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
