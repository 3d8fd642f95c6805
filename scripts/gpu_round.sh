#!/bin/bash
# One GPU-box visit: build check, parity tests, bench, launch list.  Everything lands in gpurun_out/.
# Stages are selected with STAGES="fp32 bench32 tc bench16 ncu" (default: all but ncu).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
STAGES=${STAGES:-"fp32 bench32 tc bench16"}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/smi.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
PYT="python -m pytest tests -m gpu -q --timeout 240 --timeout-method=thread -p no:cacheprovider"
for st in $STAGES; do
  case $st in
    fp32)
      timeout 600 $PYT -k "not tensor_cores" > gpurun_out/pytest_gpu_fp32.log 2>&1
      echo "[fp32 tests] exit $?"; tail -3 gpurun_out/pytest_gpu_fp32.log ;;
    tc)
      timeout 400 $PYT -k "tensor_cores" > gpurun_out/pytest_gpu_tc.log 2>&1
      echo "[tc tests] exit $?"; tail -15 gpurun_out/pytest_gpu_tc.log ;;
    bench32)
      timeout 500 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
      echo "[bench fp32] exit $?"; head -c 2500 gpurun_out/bench_fp32.json; tail -4 gpurun_out/bench_fp32.err ;;
    bench16)
      export CDS_DEBUG=1
      timeout 400 python bench.py --math bf16 --steps ${BENCH_STEPS:-3} --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
      echo "[bench bf16] exit $?"; head -c 2500 gpurun_out/bench_bf16.json; tail -4 gpurun_out/bench_bf16.err ;;
    ncu)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-300} -c ${NCU_COUNT:-130} --csv \
         --log-file gpurun_out/launches_${NCU_MATH:-bf16}.csv python bench.py --math ${NCU_MATH:-bf16} --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
      echo "[ncu] exit $?"; tail -3 gpurun_out/ncu_bench.log ;;
  esac
done
