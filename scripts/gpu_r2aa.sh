#!/bin/bash
# Round 2, visit AA: compute-sanitizer memcheck over the round's new kernels (attention_tma, linear_ln in both forms) on small shapes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_attention_gpu.py -m gpu -q -k "5-17-2 or 1-7-1 or 2-113-5 or 2-128-4" -p no:cacheprovider > gpurun_out/sanitizer_attn.log 2>&1
echo "[memcheck attention] exit $?"; tail -6 gpurun_out/sanitizer_attn.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_engine_gpu.py -m gpu -q -k "fused_linear or run_range_splits" -p no:cacheprovider > gpurun_out/sanitizer_ll.log 2>&1
echo "[memcheck linear_ln] exit $?"; tail -6 gpurun_out/sanitizer_ll.log | cut -c1-200
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/sanitizer_attn.log gpurun_out/sanitizer_ll.log
