#!/bin/bash
# Round 2, visit N: fused Linear + gate + residual + LayerNorm kernel (DiT1d); split-policy A/B of the conv kernel in TF32.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q -k "fused_linear or dit" --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_fuse_n.log 2>&1
echo "[fused tests] exit $?"; tail -5 gpurun_out/pytest_fuse_n.log; grep -E "^E " gpurun_out/pytest_fuse_n.log | head -20
timeout 600 python scripts/bench_other_cfgs.py cfg4 --math tf32 > gpurun_out/other_cfgs_tf32.jsonl 2> gpurun_out/other_cfgs_tf32.err
echo "[cfg4 tf32] exit $?"; cut -c1-400 gpurun_out/other_cfgs_tf32.jsonl; grep -E "cfg4\]" gpurun_out/other_cfgs_tf32.err | head -30
CDS_FUSE_LN=0 timeout 600 python scripts/bench_other_cfgs.py cfg4 --math tf32 --once > gpurun_out/other_cfgs_tf32_nofuse.jsonl 2> gpurun_out/other_cfgs_tf32_nofuse.err
echo "[cfg4 tf32 unfused] exit $?"; cut -c1-400 gpurun_out/other_cfgs_tf32_nofuse.jsonl
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu_n.log 2>&1
echo "[tests] exit $?"; tail -8 gpurun_out/pytest_gpu_n.log
for knob in "" "CDS_TC_NOSPLIT=1" "CDS_TC_NOSPLIT=2" "CDS_TC_SPLIT64=1"; do
  env $knob timeout 400 python bench.py --math tf32 --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --no-eager-baseline > gpurun_out/bench_tf32_knob.json 2> gpurun_out/bench_tf32_knob.err
  echo "[bench tf32 $knob] exit $?"; grep -E "timed:" gpurun_out/bench_tf32_knob.err
done
