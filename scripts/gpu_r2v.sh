#!/bin/bash
# Round 2, visit V: TF32 split policy (64-wide layers whole) + conv_ps residual prefetch: full GPU suite, cfg2 bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu_v.log 2>&1
echo "[tests] exit $?"; tail -4 gpurun_out/pytest_gpu_v.log
for i in 1 2; do
timeout 400 python bench.py --math tf32 --steps 8 --warmup 3 --no-cpu-baseline --no-other-configs --no-eager-baseline > gpurun_out/bench_tf32_v.json 2> gpurun_out/bench_tf32_v.err
echo "[bench tf32] exit $?"; grep -E "timed:|iteration total" gpurun_out/bench_tf32_v.err
done
grep -E "op +[0-9]+ conv" gpurun_out/bench_tf32_v.err | cut -c18-120
