#!/bin/bash
# Round 2, visit A: parity suite (all math modes), TF32 + bf16 bench lines with per-op times.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1; nproc >> gpurun_out/smi.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -x -s > gpurun_out/pytest_gpu.log 2>&1
echo "[tests] exit $?"; grep -E "^cfg[0-9] |passed|failed|Error|error" gpurun_out/pytest_gpu.log | tail -25
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -s > gpurun_out/pytest_gpu_all.log 2>&1
echo "[tests, no -x] exit $?"; tail -15 gpurun_out/pytest_gpu_all.log
for m in tf32 bf16; do
  timeout 400 python bench.py --math $m --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
  echo "[bench $m] exit $?"; head -c 700 gpurun_out/bench_$m.json; echo; grep -E "timed:|iteration total" gpurun_out/bench_$m.err
done
