#!/bin/bash
# Round-end GPU visit: full parity suite, smoke, both bench arms, ncu launch list (time + DRAM bytes per launch, warm L2)
# and full-set captures of the two dominant kernels.  Everything lands in gpurun_out/ (copy what is judged into profiles/).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1; nproc >> gpurun_out/smi.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "[tests] exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "[smoke] exit $?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "[bench] exit $?"; head -c 600 gpurun_out/bench.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "[bench reference] exit $?"; head -c 400 gpurun_out/bench_reference.json; echo
if [ -z "$SKIP_NCU" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
     -s ${NCU_SKIP:-1000} -c ${NCU_COUNT:-130} --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "[ncu launch list] exit $?"
  NCU_KERNEL='conv_(tc|ps)_kernel' NCU_SPECS="${NCU_SPECS:-40:1:first,57:1:ps}" bash scripts/ncu_tc.sh > gpurun_out/ncu_full.log 2>&1; echo "[ncu full] exit $?"
  rm -f gpurun_out/*.source.csv
fi
du -sh gpurun_out
