#!/bin/bash
# Round 2, visit L: attention operator tests; attention kernels alone (timing + ncu --set full of both variants).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_attention_gpu.py -m gpu -q --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_attn_l.log 2>&1
echo "[attention tests] exit $?"; tail -5 gpurun_out/pytest_attn_l.log; grep -E "^E " gpurun_out/pytest_attn_l.log | head -20
for v in 1 0; do
  CDS_ATTN_TMA=$v timeout 120 python scripts/attn_bench.py
  CDS_ATTN_TMA=$v timeout 300 ncu --set full --import-source on --clock-control none -k regex:attention -s 3 -c 1 -o gpurun_out/prof_attn_tma$v -f python scripts/attn_bench.py 4096 100 10 3 > gpurun_out/ncu_attn_tma$v.log 2>&1
  echo "[ncu attn tma=$v] exit $?"
  ncu -i gpurun_out/prof_attn_tma$v.ncu-rep --page details > gpurun_out/prof_attn_tma$v.details.txt 2>&1
done
ls -la gpurun_out | head -40
