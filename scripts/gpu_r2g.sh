#!/bin/bash
# Round 2, visit G: new GPU tests (legacy DDPM, RF); per-operator times of cfg3/4/5 (tf32); ncu launch list with DRAM bytes for cfg2 tf32;
# ncu full-set captures of the dominant cfg2 kernels; tensor-pipe metrics for DiT1d and ChiUNet1d kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -k "legacy or rectifiedflow or rf_" > gpurun_out/pytest_gpu_g.log 2>&1
echo "[tests] exit $?"; tail -5 gpurun_out/pytest_gpu_g.log
timeout 600 python scripts/bench_other_cfgs.py cfg3 cfg4 cfg5 --math tf32 > gpurun_out/other_cfgs_tf32.jsonl 2> gpurun_out/other_cfgs_tf32.err
echo "[other cfgs] exit $?"; cat gpurun_out/other_cfgs_tf32.jsonl; grep -E "cfg4\]|cfg3\] iteration|cfg5\] iteration" gpurun_out/other_cfgs_tf32.err | head -60
# launch list, one iteration of cfg2 tf32 (graph off so that every kernel is listed), warm L2 (no cache control)
CDS_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
   -s 200 -c 90 --csv --log-file gpurun_out/r02_launches_tf32.csv python scripts/one_sample.py tf32 8 > gpurun_out/ncu_list.log 2>&1
echo "[ncu launch list] exit $?"
NCU_MATH=tf32 NCU_KERNEL='conv_(tc|ps)_kernel' NCU_SPECS="42:1:r02_tf32_L32,57:1:r02_tf32_L4" bash scripts/ncu_tc.sh > gpurun_out/ncu_full.log 2>&1; echo "[ncu full] exit $?"
rm -f gpurun_out/*.source.csv gpurun_out/*.raw.csv
# tensor-pipe utilisation of the DiT1d / ChiUNet1d kernels (cfg4 / cfg3, tf32): one sample() call, first 70 launches after the warm-up call
for c in cfg4 cfg3; do
  CDS_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed \
     --clock-control none -k regex:'conv_tc_kernel|conv_ps_kernel|attention|ln_modulate|solver_update' -s 60 -c 60 --csv --log-file gpurun_out/r02_tensorpipe_$c.csv python scripts/bench_other_cfgs.py $c --math tf32 --once > gpurun_out/ncu_tp_$c.log 2>&1
  echo "[ncu tensor pipe $c] exit $?"
done
du -sh gpurun_out; ls gpurun_out | head -50
