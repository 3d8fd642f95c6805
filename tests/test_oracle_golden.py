"""Pin the oracle (oracle/*.py) against vectors produced by the unmodified reference (CPU)."""
import numpy as np
import pytest
import torch

import cases
import oracle.nets as onets
import oracle.sampler as osamp
from common import oracle_cond_emb, oracle_net, product_net, tape_of
from cleandiffuser_b200.testing import state_checksum

NET_ATOL = 1e-5       # fp32 rounding only (outputs are O(1))
SAMPLER_ATOL = 2e-4   # the same, accumulated over <= 10 reverse steps (CFG w=2.5 amplifies it)


@pytest.mark.parametrize("name", list(cases.NETS))
def test_oracle_net_matches_reference(golden, name):
    case = cases.NETS[name]
    _, sd = product_net(case)               # product module only supplies the state-dict template
    np.testing.assert_allclose(state_checksum(sd), golden["nets"][name + "/checksum"], rtol=1e-12)
    x, t, cond = cases.net_inputs(case)
    with torch.no_grad():
        y = oracle_net(case, sd)(x, t, cond)
    # same torch CPU primitives in the same order; what is left is fp32 rounding from oneDNN/MKL
    # blocking (depends on the thread count) and, for DiT, the restated MHA vs the fused fast path
    np.testing.assert_allclose(y.numpy(), golden["nets"][name + "/y"], rtol=0, atol=NET_ATOL)


def test_oracle_tables(golden):
    g = golden["tables"]
    for T in (5, 10, 100):
        for kind in ("linear", "cosine"):
            a, s = osamp.alpha_sigma(osamp.discrete_grid(T), kind)
            assert np.array_equal(a.numpy(), g[f"alpha/{kind}/{T}"])
            assert np.array_equal(s.numpy(), g[f"sigma/{kind}/{T}"])
    for key in [k for k in g.files if k.startswith("steps/")]:
        _, name, ts = key.split("/")
        T, S = map(int, ts.split("_"))
        span = [1e-3, 0.9946] if name.endswith("continuous") else T
        assert np.array_equal(osamp.step_schedule(name, span, S).numpy(), g[key]), key
    tl, tf = torch.tensor([0, 3, 99]), torch.tensor([0.001, 0.5, 1.0])
    assert np.array_equal(onets.positional_embedding(tl, 32).numpy(), g["emb/positional/long"])
    assert np.array_equal(onets.positional_embedding(tf, 32).numpy(), g["emb/positional/float"])
    assert np.array_equal(onets.sinusoidal_embedding(torch.arange(10), 32).numpy(), g["emb/sinusoidal/long"])
    # the load-bearing quirk: int64 timesteps collapse the embedding to [cos t, 1.., sin t, 0..]
    e = g["emb/positional/long"]
    assert np.all(e[:, 1:16] == 1.0) and np.all(e[:, 17:] == 0.0)


@pytest.mark.parametrize("name", list(cases.sampler_cases()))
def test_oracle_sampler_matches_reference(golden, name):
    spec = cases.sampler_cases()[name]
    netcase = cases.SAMPLER_NETS[spec["net"]]
    _, sd = product_net(netcase)
    net = oracle_net(netcase, sd)
    inp = cases.sampler_inputs(spec)
    tape = osamp.Tape(tape_of(golden["samplers"], name))
    kw = dict(steps=spec["steps"], solver=spec["solver"], schedule=spec.get("schedule", "cosine"),
              temperature=spec["temperature"], fix_mask=inp["fix_mask"][None] if inp["fix_mask"] is not None else 0.,
              predict_noise=spec["predict_noise"], cond_emb=oracle_cond_emb(spec, inp["cond"]), w_cfg=spec["w_cfg"],
              x_min=inp["x_min"], x_max=inp["x_max"], diffusion_x=spec.get("diffusion_x", 0),
              warm_start=inp["warm"], warm_level=spec.get("warm", 0.3))
    with torch.no_grad():
        if spec["kind"] == "discrete":
            x0 = osamp.sample_discrete(net, inp["prior"], tape, T=spec["T"],
                                       step_schedule_name=spec.get("step_schedule", "uniform"), **kw)
        else:
            x0 = osamp.sample_continuous(net, inp["prior"], tape,
                                         step_schedule_name=spec.get("step_schedule", "uniform_continuous"), **kw)
    assert tape.pos == len(tape.draws)
    np.testing.assert_allclose(x0.numpy(), golden["samplers"][name + "/x0"], rtol=1e-5, atol=SAMPLER_ATOL)


@pytest.mark.parametrize("steps", [1, 3])
def test_oracle_consistency_matches_reference(golden, steps):
    g = golden["consistency"]
    case = cases.NETS["chi_cm_fourier"]
    _, sd = product_net(case)
    net = oracle_net(case, sd)
    tape = osamp.Tape(tape_of(g, f"cm{steps}"))
    with torch.no_grad():
        x0 = osamp.sample_consistency(net, torch.zeros(4, 8, 3), tape, steps=steps,
                                      cond_emb=torch.as_tensor(g[f"cm{steps}/cond"]),
                                      x_min=-torch.ones(1, 8, 3), x_max=torch.ones(1, 8, 3))
    np.testing.assert_allclose(x0.numpy(), g[f"cm{steps}/x0"], rtol=1e-5, atol=SAMPLER_ATOL)
