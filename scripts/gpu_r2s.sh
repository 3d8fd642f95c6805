#!/bin/bash
# Round 2, visit S (2 GPUs): new MLP backbones on the GPU; the bench line under torchrun with 2 ranks (both arms).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q -k "pearce or sfbc or dvinv or idql" --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_mlp_s.log 2>&1
echo "[mlp backbones] exit $?"; tail -4 gpurun_out/pytest_mlp_s.log; grep -E "^E " gpurun_out/pytest_mlp_s.log | head
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_tf32_2gpu.json 2> gpurun_out/bench_tf32_2gpu.err
echo "[bench 2 gpus] exit $?"; tail -3 gpurun_out/bench_tf32_2gpu.err | cut -c1-300; cut -c1-400 gpurun_out/bench_tf32_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err
echo "[bench reference 2 gpus] exit $?"; cut -c1-300 gpurun_out/bench_ref_2gpu.json
