#!/bin/bash
# Round 2, visit R: the round's reference measurements -- full GPU suite, smoke(), the default bench line (tf32) with every block,
# the reference arm, the bf16 line, per-operator times of cfg3 / cfg4 / cfg5, the cfg2 tf32 launch list with DRAM bytes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu_r.log 2>&1
echo "[tests] exit $?"; tail -4 gpurun_out/pytest_gpu_r.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r.log 2>&1
echo "[smoke] exit $?"; tail -3 gpurun_out/smoke_r.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_tf32_r.json 2> gpurun_out/bench_tf32_r.err
echo "[bench tf32 default] exit $?"; grep -E "timed:|iteration total" gpurun_out/bench_tf32_r.err; cut -c1-300 gpurun_out/bench_tf32_r.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_r.json 2> gpurun_out/bench_reference_r.err
echo "[bench reference] exit $?"; cut -c1-300 gpurun_out/bench_reference_r.json
timeout 400 python bench.py --math bf16 --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs --no-eager-baseline > gpurun_out/bench_bf16_r.json 2> gpurun_out/bench_bf16_r.err
echo "[bench bf16] exit $?"; grep -E "timed:" gpurun_out/bench_bf16_r.err
timeout 600 python scripts/bench_other_cfgs.py cfg3 cfg4 cfg5 --math tf32 > gpurun_out/other_cfgs_tf32.jsonl 2> gpurun_out/other_cfgs_tf32.err
echo "[other cfgs] exit $?"; cut -c1-330 gpurun_out/other_cfgs_tf32.jsonl
CDS_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
   -k regex:'conv_tc_kernel|conv_ps_kernel|solver_update_kernel' -s 164 -c 82 --csv --log-file gpurun_out/r02_launches_tf32.csv python scripts/one_sample.py tf32 8 > gpurun_out/ncu_list.log 2>&1
echo "[ncu launch list] exit $?"; tail -2 gpurun_out/ncu_list.log
