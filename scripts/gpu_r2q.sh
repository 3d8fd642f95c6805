#!/bin/bash
# Round 2, visit Q: ncu --set full (with source) of the fused Linear+LayerNorm kernel inside cfg4.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
CDS_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:linear_ln -s 2 -c 1 -o gpurun_out/prof_linear_ln -f python scripts/bench_other_cfgs.py cfg4 --math tf32 --once > gpurun_out/ncu_linear_ln.log 2>&1
echo "[ncu linear_ln] exit $?"; tail -3 gpurun_out/ncu_linear_ln.log
ncu -i gpurun_out/prof_linear_ln.ncu-rep --page details > gpurun_out/prof_linear_ln.details.txt 2>&1
grep -E "Duration|Executed Ipc Active|Issue Slots Busy|No Eligible|Warp Cycles Per Issued" gpurun_out/prof_linear_ln.details.txt
