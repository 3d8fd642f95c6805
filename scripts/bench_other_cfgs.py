"""Throughput of the OTHER BASELINE.json configs on one GPU (cfg3 ChiUNet DDIM 50, cfg4 DiT1d DPM-Solver++2M 20 with two CFG
branches, cfg5 consistency ChiUNet 1 step), synthetic weights/inputs as in SURVEY 8(d); per-GPU batch = the config's batch / its
GPU count.  Not the bench contract (bench.py = cfg2): numbers for DESIGN.md, one JSON line per config.
  python scripts/bench_other_cfgs.py [cfg3 cfg4 cfg5] [--math bf16|fp32] [--reps 3]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("cfgs", nargs="*", default=["cfg3", "cfg5", "cfg4"])
ap.add_argument("--math", default="bf16")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
os.environ.update(CDS_BACKEND="cuda", CDS_MATH=args.math)

from cleandiffuser_b200.diffusion import ContinuousConsistencyModel, ContinuousDiffusionSDE, DiscreteDiffusionSDE  # noqa: E402
from cleandiffuser_b200.engine import runtime  # noqa: E402
from cleandiffuser_b200.nn_condition import IdentityCondition, MLPCondition  # noqa: E402
from cleandiffuser_b200.nn_diffusion import ChiUNet1d, DiT1d  # noqa: E402
from cleandiffuser_b200.testing import load_synth  # noqa: E402

DEV = "cuda:0"
PEAK_TF = 1443.3          # MEASURED_PEAKS.json bf16_tflops_sustained


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def per_op(agent, tag):
    """cds_plan_profile of the agent's (only) plan: device time of every operator of one iteration, to stderr."""
    plan = next(iter(agent._engine_plans.values()))
    ops = plan.program.ops
    st = torch.cuda.current_stream().cuda_stream
    plan.handle.profile(1, st, len(ops))
    t = plan.handle.profile(1, st, len(ops))
    kinds = {0: "conv", 1: "update", 2: "lnmod", 3: "attn", 4: "prep", 5: "cast"}
    tot = sum(t)
    for i, (op, ms) in enumerate(zip(ops, t)):
        d = ""
        if op.kind == 0:
            c = op.u.conv
            d = f"{'tc ' if c.math == 1 else 'f32'} rows {c.batch} L {c.L_in}->{c.L_out} C {c.C_in}->{c.C_out} k{c.taps}"
        print(f"[{tag}] op {i:2d} {kinds[op.kind]:6s} {d:48s} {ms * 1e3:9.1f} us ({ms / tot * 100:4.1f} %)", file=sys.stderr)
    print(f"[{tag}] iteration total {tot * 1e3:.1f} us", file=sys.stderr)


def report(name, batch, ms, gflop_per_traj, note):
    v = batch / (ms * 1e-3)
    print(json.dumps({"config": name, "math": args.math, "batch": batch, "ms_per_sample_call": ms, "trajectories_per_s": v,
                      "tflops_effective": v * gflop_per_traj / 1e3, "frac_of_bf16_tensor_peak": v * gflop_per_traj / 1e3 / PEAK_TF,
                      "engine": dict(calls=runtime.STATS["engine_calls"], fallbacks=runtime.STATS["fallbacks"]), "note": note}),
          flush=True)


g = torch.Generator().manual_seed(1)
with torch.no_grad():
    if "cfg3" in args.cfgs:
        B = 2048
        net = load_synth(ChiUNet1d(7, 20, 2, model_dim=256, emb_dim=256, kernel_size=5, dim_mult=[1, 2, 2]), seed=0)
        agent = DiscreteDiffusionSDE(net, IdentityCondition(dropout=0.0), predict_noise=True, diffusion_steps=1000,
                                     x_max=torch.ones(1, 16, 7), x_min=-torch.ones(1, 16, 7), device=DEV)
        prior, cond = torch.zeros(B, 16, 7, device=DEV), torch.randn(B, 40, generator=g).to(DEV)
        ms = timed(lambda: agent.sample(prior, solver="ddim", n_samples=B, sample_steps=50, condition_cfg=cond, w_cfg=1.0), args.reps)
        per_op(agent, "cfg3")
        report("cfg3 ChiUNet1d DDIM 50 w_cfg=1", B, ms, 29.85, "SURVEY 8d: 29.85 GFLOP / trajectory")
        del agent, net
    if "cfg5" in args.cfgs:
        B = 8192
        net = load_synth(ChiUNet1d(7, 20, 2, model_dim=256, emb_dim=256, kernel_size=5, dim_mult=[1, 2, 2],
                                   timestep_emb_type="untrainable_fourier"), seed=0)
        cm = ContinuousConsistencyModel(net, IdentityCondition(dropout=0.0), x_max=torch.ones(1, 16, 7),
                                        x_min=-torch.ones(1, 16, 7), device=DEV)
        prior, cond = torch.zeros(B, 16, 7, device=DEV), torch.randn(B, 40, generator=g).to(DEV)
        ms = timed(lambda: cm.sample(prior, n_samples=B, sample_steps=1, condition_cfg=cond, w_cfg=1.0), args.reps)
        report("cfg5 consistency ChiUNet1d 1 step (8192 = 65536 / 8 GPUs)", B, ms, 0.597, "SURVEY 8d: 0.597 GFLOP / trajectory")
        del cm, net
    if "cfg4" in args.cfgs:
        B = 2048
        net = load_synth(DiT1d(29, emb_dim=128, d_model=320, n_heads=10, depth=2, timestep_emb_type="fourier"), seed=0)
        mask = torch.zeros(100, 29)
        mask[0] = 1.
        agent = ContinuousDiffusionSDE(net, MLPCondition(1, 128, [128], torch.nn.SiLU(), dropout=0.25), fix_mask=mask,
                                       predict_noise=True, noise_schedule="linear", device=DEV)
        agent.model.eval(); agent.model_ema.eval()
        prior = torch.zeros(B, 100, 29, device=DEV)
        prior[:, 0] = torch.randn(B, 29, generator=g).to(DEV)
        cond = torch.rand(B, 1, generator=g).to(DEV)
        ms = timed(lambda: agent.sample(prior, solver="ode_dpmsolver++_2M", n_samples=B, sample_steps=20,
                                        sample_step_schedule="uniform_continuous", temperature=0.5, condition_cfg=cond, w_cfg=6.0),
                   args.reps)
        per_op(agent, "cfg4")
        report("cfg4 DiT1d DPM-Solver++2M 20 steps, 2 CFG branches (2048 = 16384 / 8 GPUs)", B, ms, 20.96,
               "SURVEY 8d: 20.96 GFLOP / trajectory; Linear layers on tcgen05 (bf16), attention / LayerNorm on CUDA cores (fp32)")
