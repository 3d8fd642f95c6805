#!/bin/bash
# Round 2, visit E: cluster multicast of activation tiles -- parity suite, A/B bench (CDS_NO_MULTICAST=1), per-op times.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -x > gpurun_out/pytest_gpu_e.log 2>&1
echo "[tests] exit $?"; tail -12 gpurun_out/pytest_gpu_e.log
for m in tf32 bf16; do
  for v in 0 1; do
    if [ $v = 1 ]; then export CDS_NO_MULTICAST=1; else unset CDS_NO_MULTICAST; fi
    timeout 400 python bench.py --math $m --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --no-eager-baseline > gpurun_out/bench_${m}_nm$v.json 2> gpurun_out/bench_${m}_nm$v.err
    echo "[bench $m no_multicast=$v] exit $?"; grep -E "timed:|iteration total" gpurun_out/bench_${m}_nm$v.err
  done
done
unset CDS_NO_MULTICAST
grep -E "op +[0-9]+ " gpurun_out/bench_tf32_nm0.err | awk '{print $4,$5,$6,$7,$8,$9,$10,$11,$12,$13,$14,$15,$16,$17}' | head -45
