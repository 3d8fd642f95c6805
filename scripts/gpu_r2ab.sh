#!/bin/bash
# Round 2, visit AB: compute-sanitizer memcheck over smoke() (cfg2's whole program in the three math modes, B = 64, 10 steps) and over
# one DiT1d / ChiUNet1d forward each (golden-size nets).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_smoke.log 2>&1
echo "[memcheck smoke] exit $?"; tail -6 gpurun_out/sanitizer_smoke.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_engine_gpu.py -m gpu -q -k "test_denoiser_forward_tf32_tensor_cores" -p no:cacheprovider > gpurun_out/sanitizer_nets.log 2>&1
echo "[memcheck nets tf32] exit $?"; tail -5 gpurun_out/sanitizer_nets.log | cut -c1-200
