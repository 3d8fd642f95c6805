"""ORACLE (test infrastructure, NOT product code): CPU restatement of the reverse-diffusion loop.

Restates ``DiscreteDiffusionSDE.sample`` / ``ContinuousDiffusionSDE.sample``
(cleandiffuser/diffusion/diffusionsde.py:401-606 / :743-952), the guidance / clipping helpers
(:153-223) and ``ContinuousConsistencyModel.sample`` (consistency_model.py:241-262, :366-428) as
plain functions over fp32 CPU tensors.  The denoiser is any callable ``net(x, t, cond)``; noise
comes from an explicit tape (list of tensors, consumed in draw order) so that the CUDA engine,
the PyTorch path and this oracle can be driven with identical randomness.

Pinned against golden vectors generated from the unmodified reference
(tests/golden/make_golden.py -> tests/golden/sampler_*.npz; checked by tests/test_oracle_golden.py).
Independent of ``cleandiffuser_b200`` on purpose: nothing here imports the product.
"""
import math

import numpy as np
import torch

SOLVERS = ("ddpm", "ddim", "ode_dpmsolver_1", "ode_dpmsolver++_1", "ode_dpmsolver++_2M",
           "sde_dpmsolver_1", "sde_dpmsolver++_1", "sde_dpmsolver++_2M")


# ----------------------------------------------------------------------------- schedules
def discrete_grid(T, eps=1e-3):
    """utils/utils.py:89-90."""
    return torch.linspace(eps, 1.0, T)


def alpha_sigma(t, kind="cosine", **p):
    """utils/utils.py:99-105 (linear), :124-128 (cosine)."""
    if kind == "linear":
        b0, b1 = p.get("beta0", 0.1), p.get("beta1", 20.0)
        alpha = (-(b1 - b0) / 4.0 * (t ** 2) - b0 / 2.0 * t).exp()
    elif kind == "cosine":
        s = p.get("s", 0.008)
        alpha = (np.pi / 2.0 * (t.clip(0., 0.9946) + s) / (1 + s)).cos() / np.cos(np.pi / 2.0 * s / (1 + s))
    else:
        raise ValueError(kind)
    return alpha, (1.0 - alpha ** 2).sqrt()


def step_schedule(name, span, steps):
    """utils/utils.py:157-233.  ``span`` is T (int) for discrete names, [lo, hi] for *_continuous."""
    u = torch.linspace(0, 1, steps + 1, dtype=torch.float32)
    cont = name.endswith("_continuous")
    base = name[:-len("_continuous")] if cont else name
    if base == "uniform":
        if cont:
            return torch.linspace(span[0], span[1], steps + 1, dtype=torch.float32)
        return torch.linspace(0, span - 1, steps + 1, dtype=torch.long)
    if base == "quad":
        ramp = u ** 1.5
    elif base == "cat_cos":
        ramp = 0.5 * (2 * (u > 0.5) - 1) * torch.sin(np.pi * torch.abs(u - 0.5)) ** (1 / 2.0) + 0.5
    elif base == "quad_cos":
        ramp = ((torch.sin(np.pi * (u - 0.5)) + 1) / 2) ** 2.0
    else:
        raise ValueError(name)
    if cont:
        return (span[1] - span[0]) * ramp + span[0]
    return ((span - 1) * ramp).to(torch.long)


# ----------------------------------------------------------------------------- one reverse loop
class Tape:
    """Replays pre-drawn standard-normal tensors in the order the sampler asks for them."""

    def __init__(self, draws):
        self.draws, self.pos = list(draws), 0

    def __call__(self, like):
        z = torch.as_tensor(self.draws[self.pos], dtype=torch.float32)
        self.pos += 1
        assert z.shape == like.shape, (z.shape, like.shape)
        return z


def guided_prediction(net, x, t, cond_emb, w_cfg):
    """classifier_free_guidance, diffusionsde.py:175-206 (no classifier guidance: w_cg = 0)."""
    if w_cfg != 0.0 and w_cfg != 1.0:
        b = x.shape[0]
        both = net(torch.cat([x, x], 0), torch.cat([t, t], 0),
                   torch.cat([cond_emb, torch.zeros_like(cond_emb)], 0))
        pc, pu = both[:b], both[b:]
    elif w_cfg == 0.0:
        pc, pu = 0., net(x, t, None)
    else:
        pc, pu = net(x, t, cond_emb), 0.
    return w_cfg * pc + (1 - w_cfg) * pu


def clip_prediction(pred, x, alpha, sigma, x_min, x_max, predict_noise):
    """diffusionsde.py:208-223."""
    if x_min is None and x_max is None:
        return pred
    if predict_noise:
        hi = (x - alpha * x_min) / sigma if x_min is not None else None
        lo = (x - alpha * x_max) / sigma if x_max is not None else None
        return pred.clip(lo, hi)
    return pred.clip(x_min, x_max)


def reverse_loop(net, x, prior, fix_mask, alphas, sigmas, t_values, t_dtype, solver, steps, tape, *,
                 predict_noise=True, cond_emb=None, w_cfg=0.0, x_min=None, x_max=None, diffusion_x=0,
                 trace=None):
    """diffusionsde.py:514-594.  ``alphas/sigmas/t_values`` have steps+1 entries (index 0 = data end)."""
    lam = torch.log(alphas / sigmas)
    h = torch.zeros_like(lam)
    h[1:] = lam[:-1] - lam[1:]
    std = torch.zeros(steps + 1)
    std[1:] = sigmas[:-1] / sigmas[1:] * (1 - (alphas[1:] / alphas[:-1]) ** 2).sqrt()
    n = x.shape[0]
    hist = []
    for i in reversed([1] * diffusion_x + list(range(1, steps + 1))):
        a, s, ap, sp = alphas[i], sigmas[i], alphas[i - 1], sigmas[i - 1]
        t = torch.full((n,), t_values[i], dtype=t_dtype)
        pred = guided_prediction(net, x, t, cond_emb, w_cfg)
        pred = clip_prediction(pred, x, a, s, x_min, x_max, predict_noise)
        eps = pred if predict_noise else (x - a * pred) / s                    # :21-32
        x0 = pred if not predict_noise else (x - s * pred) / a
        if solver == "ddpm":
            x_new = (ap / a) * (x - s * eps) + (sp ** 2 - std[i] ** 2 + 1e-8).sqrt() * eps
            if i > 1:
                x_new = x_new + std[i] * tape(x_new)
        elif solver == "ddim":
            x_new = ap * ((x - s * eps) / a) + sp * eps
        elif solver == "ode_dpmsolver_1":
            x_new = (ap / a) * x - sp * torch.expm1(h[i]) * eps
        elif solver == "ode_dpmsolver++_1":
            x_new = (sp / s) * x - ap * torch.expm1(-h[i]) * x0
        elif solver == "sde_dpmsolver_1":
            x_new = (ap / a) * x - 2 * sp * torch.expm1(h[i]) * eps + sp * torch.expm1(2 * h[i]).sqrt() * tape(x)
        elif solver in ("ode_dpmsolver++_2M", "sde_dpmsolver++_1", "sde_dpmsolver++_2M"):
            hist.append(x0)
            d = x0
            if solver.endswith("2M") and i < steps:
                r = h[i + 1] / h[i]
                d = (1 + 0.5 / r) * hist[-1] - 0.5 / r * hist[-2]
            if solver.startswith("ode"):
                x_new = (sp / s) * x - ap * torch.expm1(-h[i]) * d
            else:
                x_new = ((sp / s) * (-h[i]).exp() * x - ap * torch.expm1(-2 * h[i]) * d
                         + sp * (-torch.expm1(-2 * h[i])).sqrt() * tape(x))
        else:
            raise AssertionError(solver)
        x = x_new * (1. - fix_mask) + prior * fix_mask                          # :592
        if trace is not None:
            trace.append(x.clone())
    return x


def sample_discrete(net, prior, tape, *, T, steps, solver="ddpm", schedule="cosine", schedule_params=None,
                    step_schedule_name="uniform", eps=1e-3, temperature=1.0, fix_mask=0., predict_noise=True,
                    cond_emb=None, w_cfg=0.0, x_min=None, x_max=None, diffusion_x=0,
                    warm_start=None, warm_level=0.3, trace=None):
    """DiscreteDiffusionSDE.sample (diffusionsde.py:401-606) with ``cond_emb`` = output of nn_condition."""
    assert solver in SOLVERS
    alpha_T, sigma_T = alpha_sigma(discrete_grid(T, eps), schedule, **(schedule_params or {}))
    if warm_start is not None:
        T_eff = int(warm_level * T)
        x = warm_start * alpha_T[T_eff] + sigma_T[T_eff] * tape(warm_start)
    else:
        T_eff = T
        x = tape(prior) * temperature
    x = x * (1. - fix_mask) + prior * fix_mask
    idx = step_schedule(step_schedule_name, T_eff, steps)
    x = reverse_loop(net, x, prior, fix_mask, alpha_T[idx], sigma_T[idx], idx, torch.long, solver, steps, tape,
                     predict_noise=predict_noise, cond_emb=cond_emb, w_cfg=w_cfg, x_min=x_min, x_max=x_max,
                     diffusion_x=diffusion_x, trace=trace)
    if x_min is not None or x_max is not None:
        x = x.clip(x_min, x_max)                                                # :603-604
    return x


def sample_continuous(net, prior, tape, *, steps, solver="ddpm", schedule="cosine", schedule_params=None,
                      step_schedule_name="uniform_continuous", eps=1e-3, temperature=1.0, fix_mask=0.,
                      predict_noise=True, cond_emb=None, w_cfg=0.0, x_min=None, x_max=None, diffusion_x=0,
                      warm_start=None, warm_level=0.3, trace=None):
    """ContinuousDiffusionSDE.sample (diffusionsde.py:743-952)."""
    assert solver in SOLVERS
    span = [eps, 0.9946] if schedule == "cosine" else [eps, 1.]
    if warm_start is not None and warm_level > 0.:
        lvl = eps + warm_level * (1. - eps)
        a_w, s_w = alpha_sigma(torch.ones((1,)) * lvl, schedule, **(schedule_params or {}))
        x = warm_start * a_w + s_w * tape(warm_start)
        span = [span[0], lvl]
    else:
        x = tape(prior) * temperature
    x = x * (1. - fix_mask) + prior * fix_mask
    times = step_schedule(step_schedule_name, span, steps)
    alphas, sigmas = alpha_sigma(times, schedule, **(schedule_params or {}))
    x = reverse_loop(net, x, prior, fix_mask, alphas, sigmas, times, torch.float32, solver, steps, tape,
                     predict_noise=predict_noise, cond_emb=cond_emb, w_cfg=w_cfg, x_min=x_min, x_max=x_max,
                     diffusion_x=diffusion_x, trace=trace)
    if x_min is not None or x_max is not None:
        x = x.clip(x_min, x_max)
    return x


# ----------------------------------------------------------------------------- consistency model
def cm_denoise(net, x, sigma, cond_emb, *, sigma_data=0.5, sigma_min=0.002, x_min=None, x_max=None):
    """ContinuousConsistencyModel.f, consistency_model.py:241-262; ``sigma`` is a (b,) tensor."""
    c_skip = sigma_data ** 2 / (sigma_data ** 2 + (sigma - sigma_min) ** 2)
    c_out = (sigma - sigma_min) * sigma_data / (sigma_data ** 2 + sigma ** 2).sqrt()
    c_in = 1 / (sigma_data ** 2 + sigma ** 2).sqrt()
    shape = (-1,) + (1,) * (x.dim() - 1)
    out = c_skip.reshape(shape) * x + c_out.reshape(shape) * net(c_in.reshape(shape) * x, 0.25 * sigma.log(), cond_emb)
    if x_min is not None or x_max is not None:
        out = out.clip(x_min, x_max)
    return out


def sample_consistency(net, prior, tape, *, steps=1, temperature=1.0, fix_mask=0., cond_emb=None,
                       sigma_data=0.5, sigma_min=0.002, sigma_max=80., rho=7.0, x_min=None, x_max=None,
                       diffusion_x=0):
    """ContinuousConsistencyModel.sample, consistency_model.py:366-428."""
    n = prior.shape[0]
    x = tape(prior) * sigma_max * temperature
    x = x * (1. - fix_mask) + prior * fix_mask
    sig = ((sigma_min ** (1 / rho) + torch.arange(steps + 1) / steps
            * (sigma_max ** (1 / rho) - sigma_min ** (1 / rho))) ** rho)
    kw = dict(sigma_data=sigma_data, sigma_min=sigma_min, x_min=x_min, x_max=x_max)
    t = torch.full((n,), sig[-1], dtype=torch.float32)
    px = cm_denoise(net, x, t, cond_emb, **kw)
    px = px * (1. - fix_mask) + prior * fix_mask
    for i in reversed([1] * diffusion_x + list(range(1, steps))):
        t = torch.full((n,), sig[i], dtype=torch.float32)
        tt = t.reshape((-1,) + (1,) * (x.dim() - 1))
        x = px + (tt ** 2 - sigma_min ** 2).sqrt() * tape(x)
        px = cm_denoise(net, x, t, cond_emb, **kw)
        px = px * (1. - fix_mask) + prior * fix_mask
    return px


# ----------------------------------------------------------------------------- EDM
def edm_denoise(net, x, sigma, cond_emb, w_cfg, *, sigma_data=0.5):
    """ContinuousEDM.D under classifier-free guidance, newedm.py:142-148, :240-269; ``sigma`` is a (b,) tensor."""
    def D(xx, ss, cc):
        shape = (-1,) + (1,) * (xx.dim() - 1)
        c_skip = sigma_data ** 2 / (sigma_data ** 2 + ss ** 2)
        c_out = ss * sigma_data / (sigma_data ** 2 + ss ** 2).sqrt()
        c_in = 1 / (sigma_data ** 2 + ss ** 2).sqrt()
        return c_skip.reshape(shape) * xx + c_out.reshape(shape) * net(c_in.reshape(shape) * xx, 0.25 * ss.log(), cc)
    if w_cfg != 0.0 and w_cfg != 1.0:
        b = x.shape[0]
        both = D(torch.cat([x, x], 0), torch.cat([sigma, sigma], 0), torch.cat([cond_emb, torch.zeros_like(cond_emb)], 0))
        pc, pu = both[:b], both[b:]
    elif w_cfg == 0.0:
        pc, pu = 0., D(x, sigma, None)
    else:
        pc, pu = D(x, sigma, cond_emb), 0.
    return w_cfg * pc + (1 - w_cfg) * pu


def sample_edm(net, prior, tape, *, steps, solver="euler", temperature=1.0, fix_mask=0., cond_emb=None, w_cfg=0.0,
               sigma_data=0.5, sigma_min=0.002, sigma_max=80., rho=7.0, x_min=None, x_max=None, diffusion_x=0,
               warm_start=None, warm_level=0.3):
    """ContinuousEDM.sample (no classifier), newedm.py:372-438: Euler / Heun on the Karras grid."""
    assert solver in ("euler", "heun")
    n = prior.shape[0]
    if warm_start is not None and warm_level > 0.:
        top = sigma_min + (sigma_max - sigma_min) * warm_level
        x = warm_start + top * tape(warm_start)
    else:
        top = sigma_max
        x = tape(prior) * sigma_max * temperature
    x = x * (1. - fix_mask) + prior * fix_mask
    sig = (sigma_min ** (1 / rho) + torch.arange(steps + 1) / steps * (top ** (1 / rho) - sigma_min ** (1 / rho))) ** rho
    clip = x_min is not None or x_max is not None

    def slope(xx, s_eval, s_div):
        pred = edm_denoise(net, xx, torch.full((n,), s_eval, dtype=torch.float32), cond_emb, w_cfg, sigma_data=sigma_data)
        if clip:
            pred = pred.clip(x_min, x_max)
        return (xx - pred) / s_div
    for i in reversed([1] * diffusion_x + list(range(1, steps + 1))):
        d = slope(x, sig[i], sig[i])
        dt = sig[i] - sig[i - 1]
        nxt = (x - d * dt) * (1. - fix_mask) + prior * fix_mask
        if solver == "heun" and i > 1:
            d2 = slope(nxt, sig[i] / sig[i] * sig[i - 1], sig[i - 1])
            nxt = (x - (d + d2) / 2. * dt) * (1. - fix_mask) + prior * fix_mask
        x = nxt
    if clip:
        x = x.clip(x_min, x_max)
    return x


# ----------------------------------------------------------------------------- legacy DDPM (beta-schedule parameterisation)
def sample_legacy_ddpm(net, prior, tape, *, T, beta, predict_noise=True, temperature=1.0, fix_mask=0., cond_emb=None, w_cfg=0.0,
                       x_min=None, x_max=None, extra_steps=0):
    """DDPM.sample / DDPM.sample_x, ddpm.py:168-253, :256-378 (no classifier); ``beta``: the fp32 beta schedule tensor [T]."""
    alpha = 1 - beta
    bar = torch.cumprod(alpha.clone(), 0)
    n = prior.shape[0]
    x = tape(prior) * temperature
    x = x * (1. - fix_mask) + prior * fix_mask
    for k, t in enumerate(list(range(T - 1, -1, -1)) + [0] * extra_steps):
        ab = bar[t]
        ab_prev = bar[t - 1] if t > 0 else torch.tensor(1.0)
        a, b = alpha[t], beta[t]
        pred = guided_prediction(net, x, torch.full((n,), t, dtype=torch.long), cond_emb, w_cfg)
        if predict_noise:
            if x_min is not None or x_max is not None:
                hi = (x - ab.sqrt() * x_min) / (1 - ab).sqrt() if x_min is not None else None
                lo = (x - ab.sqrt() * x_max) / (1 - ab).sqrt() if x_max is not None else None
                pred = pred.clip(lo, hi)
            x = 1 / a.sqrt() * (x - b / (1 - ab).sqrt() * pred)
        else:
            if x_min is not None or x_max is not None:
                pred = pred.clip(x_min, x_max)
            x = 1 / (1 - ab) * (a.sqrt() * (1 - ab_prev) * x + b * ab_prev.sqrt() * pred)
        if t != 0 and k < T:
            x = x + (b * (1 - ab_prev) / (1 - ab)).sqrt() * tape(x)
        x = x * (1. - fix_mask) + prior * fix_mask
    return x
