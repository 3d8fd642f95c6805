#!/bin/bash
# Round 2, visit AH: EDM / legacy-EDM / consistency GPU tests after the schedule-cache change (prebuilt libcds.so, no rebuild).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_edm.py tests/test_legacy_ddpm.py tests/test_engine_gpu.py -m gpu -q -k "edm or consistency or legacy" --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu_ah.log 2>&1
echo "[tests] exit $?"; tail -3 gpurun_out/pytest_gpu_ah.log; grep -E "^E |^FAILED" gpurun_out/pytest_gpu_ah.log | head
