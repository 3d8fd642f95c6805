#!/bin/bash
# Round 2, visit AG: the two new sampler goldens on the GPU + the whole GPU suite once more.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu_ag.log 2>&1
echo "[tests] exit $?"; tail -3 gpurun_out/pytest_gpu_ag.log; grep -E "^E |^FAILED" gpurun_out/pytest_gpu_ag.log | head
