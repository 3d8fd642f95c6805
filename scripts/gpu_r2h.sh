#!/bin/bash
# Round 2, visit H: full GPU suite; per-operator times of cfg4/cfg5 (tf32) after the LayerNorm / gated-lane changes; cfg2 tf32 launch list
# with DRAM bytes (kernels of this library only, two iterations).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu_h.log 2>&1
echo "[tests] exit $?"; tail -8 gpurun_out/pytest_gpu_h.log
timeout 600 python scripts/bench_other_cfgs.py cfg4 cfg5 cfg3 --math tf32 > gpurun_out/other_cfgs_tf32.jsonl 2> gpurun_out/other_cfgs_tf32.err
echo "[other cfgs] exit $?"; cut -c1-400 gpurun_out/other_cfgs_tf32.jsonl; grep -E "cfg4\]" gpurun_out/other_cfgs_tf32.err | head -30
CDS_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
   -k regex:'conv_tc_kernel|conv_ps_kernel|solver_update_kernel' -s 164 -c 82 --csv --log-file gpurun_out/r02_launches_tf32.csv python scripts/one_sample.py tf32 8 > gpurun_out/ncu_list.log 2>&1
echo "[ncu launch list] exit $?"; tail -2 gpurun_out/ncu_list.log
