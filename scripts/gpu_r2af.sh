#!/bin/bash
# Round 2, visit AF: final validation of HEAD: full GPU suite, smoke(), cfg4 timing.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu_af.log 2>&1
echo "[tests] exit $?"; tail -3 gpurun_out/pytest_gpu_af.log; grep -E "^E |^FAILED" gpurun_out/pytest_gpu_af.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_af.log 2>&1
echo "[smoke] exit $?"; tail -1 gpurun_out/smoke_af.log
timeout 600 python scripts/bench_other_cfgs.py cfg4 --math tf32 --once > gpurun_out/other_cfgs_tf32.jsonl 2> gpurun_out/other_cfgs_tf32.err
echo "[cfg4] exit $?"; cut -c1-330 gpurun_out/other_cfgs_tf32.jsonl
