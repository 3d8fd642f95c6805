"""Deterministic synthetic weights / inputs and a replayable noise tape (used by tests, bench, smoke).

SURVEY 8(d) fixes the synthetic-input recipe so builder and judge measure the same thing:
every parameter is overwritten (DiT's zero-initialised adaLN / output head included -- otherwise
parity tests are vacuous, quirk 7): matrices ~ N(0, 1/fan_in), biases ~ N(0, 0.02^2),
norm gains 1 + N(0, 0.1^2), random-Fourier ``freqs`` ~ 16*N(0,1).  Values come from a CPU
``torch.Generator`` walked in ``state_dict`` order, so they are identical on every machine.
"""
import contextlib
from typing import Dict, List

import numpy as np
import torch


def synth_state_dict(template: Dict[str, torch.Tensor], seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = {}
    for name, ref in template.items():
        shape = tuple(ref.shape)
        z = torch.randn(shape, generator=g, dtype=torch.float32)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "freqs":
            v = z * 16.0
        elif leaf in ("bias", "in_proj_bias", "b"):
            v = z * 0.02
        elif leaf == "g" or (ref.dim() == 1 and leaf == "weight"):
            v = 1.0 + 0.1 * z
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            v = z / float(np.sqrt(fan_in))
        out[name] = v.to(ref.dtype)
    return out


def load_synth(module: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    sd = synth_state_dict(module.state_dict(), seed)
    module.load_state_dict(sd)
    return module


def state_checksum(sd: Dict[str, torch.Tensor]) -> List[float]:
    """Order-dependent fingerprint used by the golden files to detect recipe drift."""
    s1 = sum(float(v.double().sum()) for v in sd.values())
    s2 = sum(float(v.double().abs().sum()) for v in sd.values())
    return [s1, s2, float(sum(v.numel() for v in sd.values()))]


class NoiseTape:
    """Record (``draws=None``) or replay every ``torch.randn_like`` issued inside the ``with`` block.

    The reference draws its noise with bare ``torch.randn_like`` calls (diffusionsde.py:493,548,571...);
    patching that one symbol in the *harness* lets the unmodified reference, our PyTorch path and the
    CUDA engine all consume the same normal deviates.  Replayed tensors are moved to the device/dtype
    of the ``like`` argument."""

    def __init__(self, draws=None):
        self.replay = draws is not None
        self.draws = list(draws) if self.replay else []
        self.pos = 0

    def _randn_like(self, like, **kw):
        if self.replay:
            z = torch.as_tensor(self.draws[self.pos])
            self.pos += 1
            assert tuple(z.shape) == tuple(like.shape), (z.shape, like.shape)
            return z.to(device=like.device, dtype=like.dtype)
        z = self._orig(like, **kw)
        self.draws.append(z.detach().cpu().clone())
        return z

    @contextlib.contextmanager
    def active(self):
        self._orig = torch.randn_like
        torch.randn_like = self._randn_like
        try:
            yield self
        finally:
            torch.randn_like = self._orig


class ToyClassifier:
    """A small deterministic stand-in with the reference's classifier surface (cleandiffuser/classifier/base.py:9-79:
    ``model`` / ``model_ema`` / ``logp`` / ``gradients`` / ``train`` / ``eval``) for guided-sampling tests and goldens:
    log p(c | x_t, t) = MLP([flatten(x_t), 0.01 t]) with synthetic weights.  The real classifiers (CumRewClassifier over
    HalfJannerUNet1d ...) live in the reference and are used through the overlay; they are not part of this package."""

    def __init__(self, x_shape, device="cpu", seed=5):
        d = int(np.prod(x_shape))
        net = torch.nn.Sequential(torch.nn.Linear(d + 1, 16), torch.nn.Tanh(), torch.nn.Linear(16, 1))
        self.model = load_synth(net, seed).to(device)
        self.model_ema = self.model
        self.device = device

    def train(self):
        self.model.train()

    def eval(self):
        self.model.eval()

    def logp(self, x, noise, c=None):
        t = noise.to(torch.float32).reshape(-1, 1) * 0.01
        return self.model_ema(torch.cat([torch.flatten(x, 1), t], 1))

    def gradients(self, x, noise, c=None):
        x.requires_grad_()
        logp = self.logp(x, noise, c)
        grad = torch.autograd.grad([logp.sum()], [x])[0]
        return logp.detach(), grad.detach()
