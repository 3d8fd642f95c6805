#!/bin/bash
# Round 2, visit K: persistent TMA-fed TF32 attention, table / bf16-residual modes of the plain epilogue lane, half-warp LayerNorm rows.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_attention_gpu.py -m gpu -q --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_attn_k.log 2>&1
echo "[attention tests] exit $?"; tail -12 gpurun_out/pytest_attn_k.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu_k.log 2>&1
echo "[tests] exit $?"; tail -8 gpurun_out/pytest_gpu_k.log
timeout 600 python scripts/bench_other_cfgs.py cfg4 --math tf32 > gpurun_out/other_cfgs_tf32.jsonl 2> gpurun_out/other_cfgs_tf32.err
echo "[cfg4 tf32] exit $?"; cut -c1-400 gpurun_out/other_cfgs_tf32.jsonl; grep -E "cfg4\]" gpurun_out/other_cfgs_tf32.err | head -30
timeout 600 python scripts/bench_other_cfgs.py cfg4 --math bf16 > gpurun_out/other_cfgs_bf16.jsonl 2> gpurun_out/other_cfgs_bf16.err
echo "[cfg4 bf16] exit $?"; cut -c1-400 gpurun_out/other_cfgs_bf16.jsonl; grep -E "cfg4\]" gpurun_out/other_cfgs_bf16.err | head -30
