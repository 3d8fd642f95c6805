#!/bin/bash
# Round 2, visit Z: ncu --set full of DiT1d's fc1 and QKV launches (conv_tc<64,256> / <64,192>, TF32) of the final build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
CDS_GRAPH=0 timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'conv_tc_kernel<64, 256' -s 2 -c 1 -o gpurun_out/prof_r02f_cfg4_fc1 -f python scripts/bench_other_cfgs.py cfg4 --math tf32 --once > gpurun_out/ncu_fc1.log 2>&1
echo "[ncu fc1] exit $?"; tail -2 gpurun_out/ncu_fc1.log | cut -c1-200
CDS_GRAPH=0 timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'conv_tc_kernel<64, 192' -s 2 -c 1 -o gpurun_out/prof_r02f_cfg4_qkv -f python scripts/bench_other_cfgs.py cfg4 --math tf32 --once > gpurun_out/ncu_qkv.log 2>&1
echo "[ncu qkv] exit $?"; tail -2 gpurun_out/ncu_qkv.log | cut -c1-200
CDS_GRAPH=0 timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'linear_ln_kernel' -s 7 -c 1 -o gpurun_out/prof_r02f_cfg4_fc2ln -f python scripts/bench_other_cfgs.py cfg4 --math tf32 --once > gpurun_out/ncu_fc2.log 2>&1
echo "[ncu fc2+ln] exit $?"; tail -2 gpurun_out/ncu_fc2.log | cut -c1-200
ls -la gpurun_out/prof_r02f_cfg4* | tail
