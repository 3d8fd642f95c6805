#!/bin/bash
# Round 2, visit AC: bulk-store path in the generic epilogue lane (ChiUNet1d's FiLM convs): full GPU suite, cfg3 / cfg5 / cfg4 timings.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu_ac.log 2>&1
echo "[tests] exit $?"; tail -3 gpurun_out/pytest_gpu_ac.log; grep -E "^E |^FAILED" gpurun_out/pytest_gpu_ac.log | head
timeout 600 python scripts/bench_other_cfgs.py cfg3 cfg5 cfg4 --math tf32 > gpurun_out/other_cfgs_tf32.jsonl 2> gpurun_out/other_cfgs_tf32.err
echo "[other cfgs] exit $?"; cut -c1-330 gpurun_out/other_cfgs_tf32.jsonl; grep -E "^\[cfg3\]" gpurun_out/other_cfgs_tf32.err | cut -c1-125
timeout 600 python scripts/bench_other_cfgs.py cfg3 --math bf16 --once > gpurun_out/other_cfgs_bf16.jsonl 2> gpurun_out/other_cfgs_bf16.err
echo "[cfg3 bf16] exit $?"; cut -c1-330 gpurun_out/other_cfgs_bf16.jsonl
