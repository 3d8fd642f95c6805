"""Timeline of tensor-core conv launches (clock64 per CTA, see csrc/conv_tc.cuh kTraceSlots):
python scripts/trace_tc.py [ops...]   -> per traced op: prologue, per-tile producer/MMA/epilogue times, CTA span."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from cleandiffuser_b200.engine import cabi  # noqa: E402

os.environ.update(CDS_BACKEND="cuda", CDS_GRAPH="0")
os.environ.setdefault("CDS_MATH", "tf32")
targets = [int(a) for a in sys.argv[1:]] or [0, 4, 17, 18, 39]        # ordinal of the conv_tc launch inside one iteration
agent, _, _ = bench.build_agent("cuda:0")
B = 4096
prior = bench.make_prior(B).cuda()
lib = cabi.load()
with torch.no_grad():
    agent.sample(prior, solver="ddpm", n_samples=B, sample_steps=3, temperature=0.5)      # warm: plans, modules
torch.cuda.synchronize()
plan = next(iter(agent._engine_plans.values()))
n_tc = sum(1 for op in plan.program.ops if op.kind == 0 and op.u.conv.math in (1, 2))
SLOTS = 64
buf = torch.zeros(1024 * SLOTS, dtype=torch.int64, device="cuda")
for tgt in targets:
    buf.zero_()
    lib.cds_debug_trace(C.c_void_p(buf.data_ptr()), buf.numel(), n_tc + tgt)             # 2nd iteration of the next call
    with torch.no_grad():
        agent.sample(prior, solver="ddpm", n_samples=B, sample_steps=3, temperature=0.5)
    torch.cuda.synchronize()
    grid = lib.cds_debug_trace(None, 0, -1)
    t = buf.cpu().view(-1, SLOTS)[:grid].numpy()
    convs = [op for op in plan.program.ops if op.kind == 0 and op.u.conv.math in (1, 2)]
    c = convs[tgt].u.conv
    print(f"== tc launch {tgt}: L {c.L_in}->{c.L_out} C {c.C_in}->{c.C_out} k{c.taps} gn{c.groups} grid {grid}")
    g0 = t[:, 0].min()
    span = (t[:, 4].max() - g0) / 1e3
    print(f"   kernel span (globaltimer, first entry -> last exit): {span:.2f} us; entry spread {(t[:,0].max()-g0)/1e3:.2f} us")
    cyc = 1.0 / 1.965e3     # us per cycle at 1965 MHz
    pro = (t[:, 2] - t[:, 1]) * cyc
    tot = (t[:, 3] - t[:, 1]) * cyc
    print(f"   prologue {pro.mean():.2f} us (max {pro.max():.2f}); CTA lifetime {tot.mean():.2f} us (max {tot.max():.2f}); tiles/CTA {t[:,5].mean():.2f}")
    ntile = int(t[:, 5].max())
    for k in range(min(ntile, 6)):
        sel = t[:, 5] > k
        base = t[sel, 1]
        prod = (t[sel, 8 + 4 * k] - base) * cyc
        mma = (t[sel, 9 + 4 * k] - base) * cyc
        e0 = (t[sel, 10 + 4 * k] - base) * cyc
        e1 = (t[sel, 11 + 4 * k] - base) * cyc
        print(f"   tile {k}: producer done {prod.mean():6.2f}  mma first operands {mma.mean():6.2f}  epi start {e0.mean():6.2f}  "
              f"epi end {e1.mean():6.2f}  (epi {((e1-e0)).mean():.2f} us, max end {e1.max():.2f})")
