#!/bin/bash
# Round 2, visit AE: schedule-identity fast path (no device->host reads between calls): full GPU suite, smoke(), default bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu_ae.log 2>&1
echo "[tests] exit $?"; tail -3 gpurun_out/pytest_gpu_ae.log; grep -E "^E |^FAILED" gpurun_out/pytest_gpu_ae.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_ae.log 2>&1
echo "[smoke] exit $?"; tail -3 gpurun_out/smoke_ae.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_tf32_ae.json 2> gpurun_out/bench_tf32_ae.err
echo "[bench tf32 default] exit $?"; grep -E "timed:|e2e:|cfg[345]:" gpurun_out/bench_tf32_ae.err
timeout 600 python scripts/bench_other_cfgs.py cfg5 cfg3 --math tf32 > gpurun_out/other_cfgs_tf32.jsonl 2> gpurun_out/other_cfgs_tf32.err
echo "[other cfgs] exit $?"; cut -c1-330 gpurun_out/other_cfgs_tf32.jsonl
