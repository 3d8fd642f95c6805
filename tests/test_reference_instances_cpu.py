"""Drop-in boundary against the REAL reference: modules built from the reference's own
classes must lower to the engine's operator program with no fallback, and ``cleandiffuser_b200.install()`` must let the
call sequence of an unmodified pipeline run on this package's sampler classes.

Needs the reference tree (``/root/reference``, present in the build container only): every test here is skipped where it is
absent (the GPU box), and nothing under ``-m gpu`` depends on it.  The lowered programs run on the numpy interpreter of the
ABI (tests/emulator.py), so what is pinned is the lowering of REFERENCE instances; the kernels are the same ones the GPU
parity tests check.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import cases
import emulator
from cleandiffuser_b200.engine import cabi, runtime
from cleandiffuser_b200.testing import synth_state_dict

REF = os.environ.get("CDS_REFERENCE_PATH", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "cleandiffuser")), reason="reference tree not present")



@pytest.fixture()
def ref(monkeypatch):
    """The reference package imported under its own name (and removed from sys.modules afterwards)."""
    monkeypatch.syspath_prepend(REF)
    for k in [k for k in sys.modules if k == "cleandiffuser" or k.startswith("cleandiffuser.")]:
        monkeypatch.delitem(sys.modules, k)
    import cleandiffuser.nn_diffusion  # noqa: F401
    import cleandiffuser
    assert os.path.realpath(cleandiffuser.__file__).startswith(os.path.realpath(REF))
    monkeypatch.setattr(runtime, "_device_ok", lambda device: True)
    monkeypatch.setattr(runtime, "_make_handle", lambda device, ops, n: emulator.Handle(ops, n))
    monkeypatch.setenv("CDS_BACKEND", "cuda")            # any fallback to the PyTorch loop raises
    yield cleandiffuser
    for k in [k for k in sys.modules if k == "cleandiffuser" or k.startswith("cleandiffuser.")]:
        sys.modules.pop(k, None)


@pytest.mark.parametrize("math", ["fp32", "tf32"])
@pytest.mark.parametrize("name", list(cases.NETS))
def test_reference_backbone_instances_lower_without_fallback(golden, ref, name, math, monkeypatch):
    """JannerUNet1d / ChiUNet1d / DiT1d / DQLMlp INSTANCES OF THE REFERENCE'S CLASSES through ``engine_forward``."""
    monkeypatch.setenv("CDS_MATH", math)
    import cleandiffuser.nn_diffusion as rnn
    case = cases.NETS[name]
    net = getattr(rnn, case["cls"])(**case["ctor"]).eval()
    assert type(net).__module__.startswith("cleandiffuser.")
    net.load_state_dict(synth_state_dict(net.state_dict(), seed=0))
    x, t, cond = cases.net_inputs(case)
    want = golden["nets"][name + "/y"]
    for i in range(cases.NET_BATCH):
        y = runtime.engine_forward(net, x, t[i:i + 1], cond)           # raises lower.Unsupported if anything is not recognised
        err = np.abs(y[i].numpy() - want[i])
        if math == "fp32":
            assert err.max() < 2e-5, (name, i, err.max())
        else:
            assert err.max() < 2e-2 and err.mean() < 2e-3, (name, i, err.max(), err.mean())


def test_reference_instances_inside_product_sampler(golden, ref):
    """A reference JannerUNet1d handed to this package's DiscreteDiffusionSDE: engine call, no fallback, golden result."""
    from common import tape_of
    from cleandiffuser_b200.diffusion import DiscreteDiffusionSDE
    from cleandiffuser_b200.testing import NoiseTape
    import cleandiffuser.nn_diffusion as rnn
    name = "disc_dup_ddpm_x0"
    spec = cases.sampler_cases()[name]
    ncase = cases.SAMPLER_NETS[spec["net"]]
    net = getattr(rnn, ncase["cls"])(**ncase["ctor"]).eval()
    net.load_state_dict(synth_state_dict(net.state_dict(), seed=0))
    inp = cases.sampler_inputs(spec)
    agent = DiscreteDiffusionSDE(net, None, fix_mask=inp["fix_mask"], x_max=inp["x_max"], x_min=inp["x_min"],
                                 predict_noise=spec["predict_noise"], diffusion_steps=spec["T"], device="cpu")
    before, fb = runtime.STATS["engine_calls"], runtime.STATS["fallbacks"]
    os.environ["CDS_MATH"] = "fp32"
    tape = NoiseTape(tape_of(golden["samplers"], name))
    with tape.active(), torch.no_grad():
        x0, _ = agent.sample(inp["prior"], solver=spec["solver"], n_samples=cases.SAMPLER_BATCH, sample_steps=spec["steps"],
                             temperature=spec["temperature"], w_cfg=spec["w_cfg"])
    assert runtime.STATS["engine_calls"] == before + 1 and runtime.STATS["fallbacks"] == fb
    np.testing.assert_allclose(x0.numpy(), golden["samplers"][name + "/x0"], rtol=1e-4, atol=3e-4)


PIPELINE_STUB = r"""
import sys, types, torch
# the pipeline's imports of packages that are not in this image (simulators, config system) are stubbed; everything from
# `cleandiffuser` is the REAL reference
for name in ("d4rl", "gym", "hydra"):
    sys.modules[name] = types.ModuleType(name)
import cleandiffuser_b200
patched = cleandiffuser_b200.install()
assert "cleandiffuser.diffusion.DiscreteDiffusionSDE" in patched, patched

# ---- pipelines/diffuser_d4rl_mujoco.py:11-17 (imports), :39-66 (construction), :136-148 (inference call) -------------
from cleandiffuser.classifier import CumRewClassifier
from cleandiffuser.diffusion import DiscreteDiffusionSDE
from cleandiffuser.nn_classifier import HalfJannerUNet1d
from cleandiffuser.nn_diffusion import JannerUNet1d
from cleandiffuser.utils import report_parameters
assert DiscreteDiffusionSDE is cleandiffuser_b200.diffusion.DiscreteDiffusionSDE
assert JannerUNet1d.__module__.startswith("cleandiffuser.")

obs_dim, act_dim, horizon, model_dim, dim_mult = 11, 3, 32, 32, [1, 2, 2, 2]
nn_diffusion = JannerUNet1d(obs_dim + act_dim, model_dim=model_dim, emb_dim=model_dim, dim_mult=dim_mult,
                            timestep_emb_type="positional", attention=False, kernel_size=5)
nn_classifier = HalfJannerUNet1d(horizon, obs_dim + act_dim, out_dim=1, model_dim=model_dim, emb_dim=model_dim,
                                 dim_mult=dim_mult, timestep_emb_type="positional", kernel_size=3)
report_parameters(nn_diffusion)
classifier = CumRewClassifier(nn_classifier, device="cpu")
fix_mask = torch.zeros((horizon, obs_dim + act_dim)); fix_mask[0, :obs_dim] = 1.
loss_weight = torch.ones((horizon, obs_dim + act_dim)); loss_weight[0, obs_dim:] = 10.
agent = DiscreteDiffusionSDE(nn_diffusion, None, fix_mask=fix_mask, loss_weight=loss_weight, classifier=classifier,
                             ema_rate=0.9999, device="cpu", diffusion_steps=20, predict_noise=False)
# one training step of each (pipeline :75-90)
x = torch.randn(8, horizon, obs_dim + act_dim); R = torch.randn(8, 1)
log = agent.update(x); assert "loss" in log
log = agent.update_classifier(x, R); assert "loss" in log
# inference (:136-148)
agent.eval()
num_envs, num_candidates = 2, 4
prior = torch.zeros((num_envs, horizon, obs_dim + act_dim)); prior[:, 0, :obs_dim] = torch.randn(num_envs, obs_dim)
traj, log = agent.sample(prior.repeat(num_candidates, 1, 1), solver="ddpm", n_samples=num_candidates * num_envs,
                         sample_steps=20, use_ema=True, w_cg=0.3, temperature=0.5)
logp = log["log_p"].view(num_candidates, num_envs, -1).sum(-1)
idx = logp.argmax(0)
act = traj.view(num_candidates, num_envs, horizon, -1)[idx, torch.arange(num_envs), 0, obs_dim:]
assert act.shape == (num_envs, act_dim) and torch.isfinite(traj).all()
assert torch.equal(traj[:, 0, :obs_dim], prior.repeat(num_candidates, 1, 1)[:, 0, :obs_dim])
cleandiffuser_b200.uninstall()
import cleandiffuser.diffusion as rd
assert rd.DiscreteDiffusionSDE.__module__.startswith("cleandiffuser.")
print("ok")
"""


def test_install_overlay_runs_the_diffuser_pipeline_call_sequence():
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([REF, ROOT]), CDS_BACKEND="auto")
    out = subprocess.run([sys.executable, "-c", PIPELINE_STUB], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-3000:]
