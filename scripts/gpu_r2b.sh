#!/bin/bash
# Round 2, visit B: fixed parity tests, TF32 per-CTA timelines, ncu full captures of three TF32 conv kernels, split experiment.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -s -k "baseline_configs or tf32" > gpurun_out/pytest_gpu_b.log 2>&1
echo "[tests] exit $?"; grep -E "^cfg[0-9] |passed|failed" gpurun_out/pytest_gpu_b.log | tail -15
CDS_MATH=tf32 timeout 300 python scripts/trace_tc.py 0 2 3 6 12 13 17 18 24 28 38 39 > gpurun_out/trace_tf32.txt 2>&1; echo "[trace] exit $?"; cat gpurun_out/trace_tf32.txt
for v in 1 2; do
  CDS_TC_NOSPLIT=$v timeout 300 python bench.py --math tf32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tf32_nosplit$v.json 2> gpurun_out/bench_tf32_nosplit$v.err
  echo "[bench tf32 nosplit=$v] exit $?"; grep -E "timed:|iteration total" gpurun_out/bench_tf32_nosplit$v.err
done
NCU_MATH=tf32 NCU_KERNEL='conv_(tc|ps)_kernel' NCU_SPECS="42:1:tf32_L32,52:1:tf32_L8,57:1:tf32_L4" bash scripts/ncu_tc.sh > gpurun_out/ncu_full.log 2>&1; echo "[ncu full] exit $?"; tail -5 gpurun_out/ncu_full.log
rm -f gpurun_out/*.source.csv
du -sh gpurun_out
