"""Product host layer (nn.Modules + PyTorch sampling path) against the reference goldens, on CPU.

This is the path every non-CUDA / autograd / custom-backbone call takes, and the scaffold the
engine plugs into; it must reproduce the reference exactly (up to fp32 rounding)."""
import numpy as np
import pytest
import torch

import cases
from common import product_condition, product_net, tape_of
from cleandiffuser_b200.diffusion import (ContinuousConsistencyModel, ContinuousDiffusionSDE,
                                          DiscreteDiffusionSDE)
from cleandiffuser_b200.nn_condition import IdentityCondition
from cleandiffuser_b200.testing import NoiseTape
from cleandiffuser_b200 import utils as U

NET_ATOL, SAMPLER_ATOL = 1e-5, 2e-4


@pytest.mark.parametrize("name", list(cases.NETS))
def test_module_state_dict_and_forward(golden, name):
    case = cases.NETS[name]
    net, sd = product_net(case)
    keys = ["%s|%s" % (k, ",".join(map(str, v.shape))) for k, v in sd.items()]
    assert keys == list(golden["nets"][name + "/keys"]), "state_dict layout differs from the reference"
    x, t, cond = cases.net_inputs(case)
    with torch.no_grad():
        y = net(x, t, cond)
    np.testing.assert_allclose(y.numpy(), golden["nets"][name + "/y"], rtol=0, atol=NET_ATOL)


def test_schedule_tables(golden):
    g = golden["tables"]
    for T in (5, 10, 100):
        grid = U.SUPPORTED_DISCRETIZATIONS["uniform"](T, 1e-3)
        for kind in ("linear", "cosine"):
            a, s = U.SUPPORTED_NOISE_SCHEDULES[kind]["forward"](grid)
            assert np.array_equal(a.numpy(), g[f"alpha/{kind}/{T}"])
            assert np.array_equal(s.numpy(), g[f"sigma/{kind}/{T}"])
    for key in [k for k in g.files if k.startswith("steps/")]:
        _, name, ts = key.split("/")
        T, S = map(int, ts.split("_"))
        span = [1e-3, 0.9946] if name.endswith("continuous") else T
        assert np.array_equal(U.SUPPORTED_SAMPLING_STEP_SCHEDULE[name](span, S).numpy(), g[key]), key
    tl, tf = torch.tensor([0, 3, 99]), torch.tensor([0.001, 0.5, 1.0])
    for kind in ("positional", "untrainable_positional"):
        emb = U.SUPPORTED_TIMESTEP_EMBEDDING[kind](32)
        assert np.array_equal(emb(tl).numpy(), g[f"emb/{kind}/long"])
        assert np.array_equal(emb(tf).numpy(), g[f"emb/{kind}/float"])
    assert np.array_equal(U.SinusoidalEmbedding(32)(torch.arange(10)).numpy(), g["emb/sinusoidal/long"])
    assert np.array_equal(U.SinusoidalEmbedding(32)(torch.arange(10).float()).numpy(), g["emb/sinusoidal/float"])


def build_agent(spec, device="cpu"):
    net, _ = product_net(cases.SAMPLER_NETS[spec["net"]])
    inp = cases.sampler_inputs(spec)
    common = dict(nn_condition=product_condition(spec), fix_mask=inp["fix_mask"], x_max=inp["x_max"],
                  x_min=inp["x_min"], predict_noise=spec["predict_noise"], device=device,
                  noise_schedule=spec.get("schedule", "cosine"))
    if spec["kind"] == "discrete":
        agent = DiscreteDiffusionSDE(net, diffusion_steps=spec["T"], **common)
        sched = spec.get("step_schedule", "uniform")
    else:
        agent = ContinuousDiffusionSDE(net, **common)
        sched = spec.get("step_schedule", "uniform_continuous")
    kw = dict(solver=spec["solver"], n_samples=cases.SAMPLER_BATCH, sample_steps=spec["steps"],
              sample_step_schedule=sched, use_ema=True, temperature=spec["temperature"],
              condition_cfg=inp["cond"], w_cfg=spec["w_cfg"],
              diffusion_x_sampling_steps=spec.get("diffusion_x", 0))
    if inp["warm"] is not None:
        kw.update(warm_start_reference=inp["warm"], warm_start_forward_level=spec["warm"])
    return agent, inp, kw


@pytest.mark.parametrize("name", list(cases.sampler_cases()))
def test_sampler_torch_path(golden, name):
    spec = cases.sampler_cases()[name]
    agent, inp, kw = build_agent(spec)
    tape = NoiseTape(tape_of(golden["samplers"], name))
    with tape.active(), torch.no_grad():
        x0, log = agent.sample(inp["prior"], **kw)
    assert tape.pos == len(tape.draws), "noise draw count/order differs from the reference"
    assert log["sample_history"] is None
    np.testing.assert_allclose(x0.numpy(), golden["samplers"][name + "/x0"], rtol=1e-5, atol=SAMPLER_ATOL)


@pytest.mark.parametrize("steps", [1, 3])
def test_consistency_torch_path(golden, steps):
    g = golden["consistency"]
    net, _ = product_net(cases.NETS["chi_cm_fourier"])
    cm = ContinuousConsistencyModel(net, IdentityCondition(dropout=0.0), x_max=torch.ones(1, 8, 3),
                                    x_min=-torch.ones(1, 8, 3), device="cpu")
    tape = NoiseTape(tape_of(g, f"cm{steps}"))
    with tape.active(), torch.no_grad():
        x0, _ = cm.sample(torch.zeros(4, 8, 3), n_samples=4, sample_steps=steps,
                          condition_cfg=torch.as_tensor(g[f"cm{steps}/cond"]), w_cfg=1.0)
    np.testing.assert_allclose(x0.numpy(), g[f"cm{steps}/x0"], rtol=1e-5, atol=SAMPLER_ATOL)


def test_api_surface_and_errors():
    net, _ = product_net(cases.SAMPLER_NETS["dql_tiny"])
    agent = DiscreteDiffusionSDE(net, None, diffusion_steps=5, device="cpu")
    assert isinstance(agent.model, torch.nn.ModuleDict) and set(agent.model.keys()) == {"diffusion", "condition"}
    assert not any(p.requires_grad for p in agent.model_ema.parameters())
    assert agent.fix_mask == 0. and agent.loss_weight == 1.
    assert agent.alpha.shape == (5,) and agent.sigma.shape == (5,)
    with pytest.raises(AssertionError):
        agent.sample(torch.zeros(2, 3), solver="euler", n_samples=2)
    with pytest.raises(ValueError):
        agent.sample(torch.zeros(2, 3), solver="ddpm", n_samples=2, sample_step_schedule="nope")
    with pytest.raises(ValueError):
        DiscreteDiffusionSDE(net, None, diffusion_steps=5000, epsilon=1e-3)
    # preserve_history reproduces the reference's (n, S+1, n, *x_shape) float64 allocation (quirk 10)
    x0, log = agent.sample(torch.zeros(2, 3), solver="ddim", n_samples=2, sample_steps=3, preserve_history=True)
    assert log["sample_history"].shape == (2, 4, 2, 3) and log["sample_history"].dtype == np.float64
    # training step + EMA + save/load round trip
    out = agent.update(torch.randn(8, 3), torch.randn(8, 4))
    assert set(out) == {"loss", "grad_norm"}
    import tempfile, os
    with tempfile.TemporaryDirectory() as d:
        agent.save(os.path.join(d, "ck.pt"))
        agent.load(os.path.join(d, "ck.pt"))
    # requires_grad=True keeps the graph (Diffusion-QL actor loss)
    agent.model.train()
    act, _ = agent.sample(torch.zeros(2, 3), solver="ddpm", n_samples=2, sample_steps=5, use_ema=False,
                          condition_cfg=torch.randn(2, 4), w_cfg=1.0, requires_grad=True)
    assert act.requires_grad
