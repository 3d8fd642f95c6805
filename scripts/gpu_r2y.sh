#!/bin/bash
# Round 2, visit Y: ncu --set full of three TF32 GEMM-class launches of the final build, for the shared-memory / L2 / tensor-pipe
# utilisation figures quoted in DESIGN.md section 5: cfg2 conv L=8 128->128 (conv_tc<64,64,0,2>), cfg2 conv L=16 64->64, cfg4 fc1
# (conv_tc<64,256>), cfg4 QKV (conv_tc<64,192>).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for spec in "52:r02f_cfg2_L8_128" "47:r02f_cfg2_L16_64" "42:r02f_cfg2_L32_32" "57:r02f_cfg2_L4_256_ps"; do
  s=${spec%%:*}; n=${spec##*:}
  CDS_GRAPH=0 timeout 300 ncu --set full --clock-control none -k regex:'conv_(tc|ps)_kernel' -s $s -c 1 -o gpurun_out/prof_$n -f python scripts/one_sample.py tf32 4 > gpurun_out/ncu_$n.log 2>&1
  echo "[ncu $n] exit $?"
done
CDS_GRAPH=0 timeout 300 ncu --set full --clock-control none -k regex:'conv_tc_kernel<64, 256' -s 2 -c 1 -o gpurun_out/prof_r02f_cfg4_fc1 -f python scripts/bench_other_cfgs.py cfg4 --math tf32 --once > gpurun_out/ncu_fc1.log 2>&1
echo "[ncu fc1] exit $?"
CDS_GRAPH=0 timeout 300 ncu --set full --clock-control none -k regex:'conv_tc_kernel<64, 192' -s 2 -c 1 -o gpurun_out/prof_r02f_cfg4_qkv -f python scripts/bench_other_cfgs.py cfg4 --math tf32 --once > gpurun_out/ncu_qkv.log 2>&1
echo "[ncu qkv] exit $?"
ls -la gpurun_out/*.ncu-rep | tail -8
