"""Drop-in route B: with compat/ on the path, the reference's own import lines resolve to this package."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SNIPPET = """
import torch
from cleandiffuser.diffusion import DiscreteDiffusionSDE, ContinuousDiffusionSDE
from cleandiffuser.nn_diffusion import JannerUNet1d, DiT1d, ChiUNet1d, DQLMlp
from cleandiffuser.nn_condition import MLPCondition, IdentityCondition
from cleandiffuser.utils import GroupNorm1d, at_least_ndim, SUPPORTED_NOISE_SCHEDULES
import cleandiffuser_b200
assert DiscreteDiffusionSDE is cleandiffuser_b200.diffusion.DiscreteDiffusionSDE
# the call sequence of pipelines/dql_d4rl_mujoco.py:39-51,184-190 on CPU (BASELINE config 1 shapes)
nn_diffusion = DQLMlp(11, 3, emb_dim=64, timestep_emb_type="positional")
actor = DiscreteDiffusionSDE(nn_diffusion, IdentityCondition(dropout=0.0), predict_noise=True, optim_params={"lr": 3e-4},
                             x_max=+1. * torch.ones((1, 3)), x_min=-1. * torch.ones((1, 3)), diffusion_steps=5,
                             ema_rate=0.995, device="cpu")
actor.eval()
act, log = actor.sample(torch.zeros((64, 3)), solver="ddpm", n_samples=64, sample_steps=5,
                        condition_cfg=torch.randn(64, 11), w_cfg=1.0, use_ema=True, temperature=0.5)
assert act.shape == (64, 3) and float(act.abs().max()) <= 1.0
print("ok")
"""


def test_reference_import_lines_resolve():
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "compat"), ROOT]))
    out = subprocess.run([sys.executable, "-c", SNIPPET], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
