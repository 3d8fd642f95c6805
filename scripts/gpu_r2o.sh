#!/bin/bash
# Round 2, visit O: fused Linear+LayerNorm kernel v2 (pipelined residual loads, two staging buffers, single-pass moments).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q -k "fused_linear or dit" --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_fuse_o.log 2>&1
echo "[fused tests] exit $?"; tail -5 gpurun_out/pytest_fuse_o.log; grep -E "^E " gpurun_out/pytest_fuse_o.log | head -20
timeout 600 python scripts/bench_other_cfgs.py cfg4 --math tf32 > gpurun_out/other_cfgs_tf32.jsonl 2> gpurun_out/other_cfgs_tf32.err
echo "[cfg4 tf32] exit $?"; cut -c1-400 gpurun_out/other_cfgs_tf32.jsonl; grep -E "cfg4\]" gpurun_out/other_cfgs_tf32.err | head -30
timeout 300 python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q -k "cfg4" --timeout 280 --timeout-method=thread -p no:cacheprovider -s > gpurun_out/pytest_cfg4_o.log 2>&1
echo "[cfg4 parity] exit $?"; grep -E "cfg4 |passed|failed" gpurun_out/pytest_cfg4_o.log
