#!/bin/bash
# Round 2, visit U (2 GPUs): the bench line under torchrun with 2 ranks after the collective-order fix.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_tf32_2gpu.json 2> gpurun_out/bench_tf32_2gpu.err
echo "[bench 2 gpus] exit $?"; grep -E "timed:|e2e:|cfg[345]:" gpurun_out/bench_tf32_2gpu.err | cut -c1-200; cut -c1-400 gpurun_out/bench_tf32_2gpu.json
