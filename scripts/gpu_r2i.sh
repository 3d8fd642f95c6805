#!/bin/bash
# Round 2, visit I: full GPU suite (residual prefetch, gated lane, vectorised LayerNorm, legacy EDM, 2M-history fix); A/B bench of the
# residual prefetch; per-operator times of cfg4; cfg2 tf32 launch list with DRAM bytes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu_i.log 2>&1
echo "[tests] exit $?"; tail -8 gpurun_out/pytest_gpu_i.log
for v in 0 1; do
  if [ $v = 1 ]; then export CDS_NO_TMA_RES=1; else unset CDS_NO_TMA_RES; fi
  timeout 400 python bench.py --math tf32 --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --no-eager-baseline > gpurun_out/bench_tf32_nr$v.json 2> gpurun_out/bench_tf32_nr$v.err
  echo "[bench tf32 no_tma_res=$v] exit $?"; grep -E "timed:|iteration total" gpurun_out/bench_tf32_nr$v.err
done
unset CDS_NO_TMA_RES
timeout 600 python scripts/bench_other_cfgs.py cfg4 cfg5 cfg3 --math tf32 > gpurun_out/other_cfgs_tf32.jsonl 2> gpurun_out/other_cfgs_tf32.err
echo "[other cfgs] exit $?"; cut -c1-400 gpurun_out/other_cfgs_tf32.jsonl; grep -E "cfg4\]" gpurun_out/other_cfgs_tf32.err | head -30
CDS_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
   -k regex:'conv_tc_kernel|conv_ps_kernel|solver_update_kernel' -s 164 -c 82 --csv --log-file gpurun_out/r02_launches_tf32.csv python scripts/one_sample.py tf32 8 > gpurun_out/ncu_list.log 2>&1
echo "[ncu launch list] exit $?"; tail -2 gpurun_out/ncu_list.log
