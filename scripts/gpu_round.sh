#!/bin/bash
# One GPU-box visit: build check, parity tests, bench, launch list.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/smi.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; cat gpurun_out/bench.json | head -c 3000; tail -5 gpurun_out/bench.err
if [ "${WITH_NCU:-0}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv \
     --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --batch 4096 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "ncu exit $?"
fi
