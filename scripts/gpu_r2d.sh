#!/bin/bash
# Round 2, visit D: TMA-store epilogue -- parity suite, A/B bench (CDS_NO_TMA_STORE=1 = direct stores), traces.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -x > gpurun_out/pytest_gpu_d.log 2>&1
echo "[tests] exit $?"; tail -12 gpurun_out/pytest_gpu_d.log
for m in tf32 bf16; do
  for v in 0 1; do
    if [ $v = 1 ]; then export CDS_NO_TMA_STORE=1; else unset CDS_NO_TMA_STORE; fi
    timeout 400 python bench.py --math $m --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --no-eager-baseline > gpurun_out/bench_${m}_nt$v.json 2> gpurun_out/bench_${m}_nt$v.err
    echo "[bench $m no_tma_store=$v] exit $?"; grep -E "timed:|iteration total" gpurun_out/bench_${m}_nt$v.err
  done
done
unset CDS_NO_TMA_STORE
CDS_MATH=tf32 timeout 300 python scripts/trace_tc.py 2 3 12 13 17 18 > gpurun_out/trace_tf32_d.txt 2>&1; echo "[trace] exit $?"; cat gpurun_out/trace_tf32_d.txt
grep -E "op +[0-9]+ " gpurun_out/bench_tf32_nt0.err | head -45
