"""Run ONE short sample() of the bench workload (for profilers): python scripts/one_sample.py [math] [steps] [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402

math = sys.argv[1] if len(sys.argv) > 1 else "bf16"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
os.environ.update(CDS_BACKEND="cuda", CDS_MATH=math)
agent, _, _ = bench.build_agent("cuda:0")
prior = bench.make_prior(batch).cuda()
with torch.no_grad():
    x, _ = agent.sample(prior, solver="ddpm", n_samples=batch, sample_steps=steps, temperature=0.5)
torch.cuda.synchronize()
print("ok", float(x.abs().mean()))
