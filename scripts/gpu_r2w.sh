#!/bin/bash
# Round 2, visit W: the round's final reference measurements (after the TF32 split policy change): full GPU suite, smoke(), default
# bench line with every block, reference arm, bf16 line, per-operator times of cfg3 / cfg4 / cfg5, cfg2 launch list with DRAM bytes,
# tensor-pipe activity per launch of one cfg4 sample().
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu_w.log 2>&1
echo "[tests] exit $?"; tail -3 gpurun_out/pytest_gpu_w.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_w.log 2>&1
echo "[smoke] exit $?"; tail -3 gpurun_out/smoke_w.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_tf32_w.json 2> gpurun_out/bench_tf32_w.err
echo "[bench tf32 default] exit $?"; grep -E "timed:|iteration total|e2e:" gpurun_out/bench_tf32_w.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_w.json 2> gpurun_out/bench_reference_w.err
echo "[bench reference] exit $?"; cut -c1-200 gpurun_out/bench_reference_w.json
timeout 400 python bench.py --math bf16 --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs --no-eager-baseline > gpurun_out/bench_bf16_w.json 2> gpurun_out/bench_bf16_w.err
echo "[bench bf16] exit $?"; grep -E "timed:" gpurun_out/bench_bf16_w.err
timeout 600 python scripts/bench_other_cfgs.py cfg3 cfg4 cfg5 --math tf32 > gpurun_out/other_cfgs_tf32.jsonl 2> gpurun_out/other_cfgs_tf32.err
echo "[other cfgs] exit $?"; cut -c1-330 gpurun_out/other_cfgs_tf32.jsonl
CDS_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
   -k regex:'conv_tc_kernel|conv_ps_kernel|solver_update_kernel' -s 164 -c 82 --csv --log-file gpurun_out/r02_launches_tf32.csv python scripts/one_sample.py tf32 8 > gpurun_out/ncu_list.log 2>&1
echo "[ncu launch list] exit $?"; tail -1 gpurun_out/ncu_list.log
CDS_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed \
   --clock-control none -k regex:'conv_tc_kernel|linear_ln|attention|ln_modulate|solver_update' -s 44 -c 44 --csv --log-file gpurun_out/r02_tensorpipe_cfg4.csv python scripts/bench_other_cfgs.py cfg4 --math tf32 --once > gpurun_out/ncu_tp_cfg4.log 2>&1
echo "[ncu tensor pipe cfg4] exit $?"; tail -1 gpurun_out/ncu_tp_cfg4.log | cut -c1-200
