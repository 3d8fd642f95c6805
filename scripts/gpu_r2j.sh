#!/bin/bash
# Round 2, visit J: 160 / 192-wide runtime column tiles (DiT1d Linear layers); row-shifted UMMA descriptor micro-test; full suite; bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 60 scripts/micro/umma_rowshift > gpurun_out/umma_rowshift.log 2>&1
echo "[micro rowshift] exit $?"; tail -20 gpurun_out/umma_rowshift.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu_j.log 2>&1
echo "[tests] exit $?"; tail -8 gpurun_out/pytest_gpu_j.log
timeout 600 python scripts/bench_other_cfgs.py cfg4 cfg5 cfg3 --math tf32 > gpurun_out/other_cfgs_tf32.jsonl 2> gpurun_out/other_cfgs_tf32.err
echo "[other cfgs] exit $?"; cut -c1-400 gpurun_out/other_cfgs_tf32.jsonl; grep -E "cfg4\]" gpurun_out/other_cfgs_tf32.err | head -30
timeout 600 python scripts/bench_other_cfgs.py cfg4 --math bf16 > gpurun_out/other_cfgs_bf16.jsonl 2> gpurun_out/other_cfgs_bf16.err
echo "[cfg4 bf16] exit $?"; cut -c1-400 gpurun_out/other_cfgs_bf16.jsonl; grep -E "cfg4\]" gpurun_out/other_cfgs_bf16.err | head -30
timeout 400 python bench.py --math tf32 --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --no-eager-baseline > gpurun_out/bench_tf32_j.json 2> gpurun_out/bench_tf32_j.err
echo "[bench tf32] exit $?"; grep -E "timed:|iteration total" gpurun_out/bench_tf32_j.err; cut -c1-600 gpurun_out/bench_tf32_j.json
