#!/bin/bash
# Round 2, visit X: final build -- full GPU suite, smoke(), default bench line, cfg4 per-operator times, ncu --set full of the final
# linear_ln kernel (gated form, K = 320).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu_x.log 2>&1
echo "[tests] exit $?"; tail -3 gpurun_out/pytest_gpu_x.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_x.log 2>&1
echo "[smoke] exit $?"; tail -3 gpurun_out/smoke_x.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_tf32_x.json 2> gpurun_out/bench_tf32_x.err
echo "[bench tf32 default] exit $?"; grep -E "timed:|iteration total|e2e:|cfg[345]:" gpurun_out/bench_tf32_x.err
timeout 600 python scripts/bench_other_cfgs.py cfg4 --math tf32 > gpurun_out/other_cfgs_tf32_cfg4.jsonl 2> gpurun_out/other_cfgs_tf32_cfg4.err
echo "[cfg4] exit $?"; cut -c1-330 gpurun_out/other_cfgs_tf32_cfg4.jsonl
CDS_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:linear_ln -s 5 -c 1 -o gpurun_out/prof_linear_ln_final -f python scripts/bench_other_cfgs.py cfg4 --math tf32 --once > gpurun_out/ncu_linear_ln.log 2>&1
echo "[ncu linear_ln] exit $?"
ncu -i gpurun_out/prof_linear_ln_final.ncu-rep --page details > gpurun_out/prof_linear_ln_final.details.txt 2>&1
grep -E "Duration|Executed Ipc Active|DRAM Throughput|No Eligible" gpurun_out/prof_linear_ln_final.details.txt
