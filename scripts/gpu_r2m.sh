#!/bin/bash
# Round 2, visit M: attention kernel v2 (templated score tiles, XOR addressing, ex2.approx, relaxed producer wait); new MLP backbones on the GPU.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_attention_gpu.py -m gpu -q --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_attn_m.log 2>&1
echo "[attention tests] exit $?"; tail -5 gpurun_out/pytest_attn_m.log; grep -E "^E " gpurun_out/pytest_attn_m.log | head -20
timeout 120 python scripts/attn_bench.py
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu_m.log 2>&1
echo "[tests] exit $?"; tail -8 gpurun_out/pytest_gpu_m.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attention -s 3 -c 1 -o gpurun_out/prof_attn_v2 -f python scripts/attn_bench.py 4096 100 10 3 > gpurun_out/ncu_attn_v2.log 2>&1
echo "[ncu attn v2] exit $?"
timeout 600 python scripts/bench_other_cfgs.py cfg4 --math tf32 > gpurun_out/other_cfgs_tf32.jsonl 2> gpurun_out/other_cfgs_tf32.err
echo "[cfg4 tf32] exit $?"; cut -c1-400 gpurun_out/other_cfgs_tf32.jsonl; grep -E "cfg4\]" gpurun_out/other_cfgs_tf32.err | head -30
