#!/bin/bash
# one GPU-box visit for A/B experiments: each VARIANT "name:ENV=val,ENV=val" runs the bf16 bench in its own process
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
if [ -n "$TESTS" ]; then
  timeout 600 python -m pytest tests -m gpu -q --timeout 240 --timeout-method=thread -p no:cacheprovider -k "$TESTS" > gpurun_out/pytest_exp.log 2>&1
  echo "[tests] exit $?"; tail -5 gpurun_out/pytest_exp.log
fi
for v in ${VARIANTS:-base:}; do
  name=${v%%:*}; envs=${v#*:}
  ( for kv in ${envs//,/ }; do export "$kv"; done
    timeout 300 python bench.py --math bf16 --steps ${BENCH_STEPS:-3} --warmup 3 --no-cpu-baseline > gpurun_out/exp_$name.json 2> gpurun_out/exp_$name.err )
  echo "[$name] exit $? $(python -c "import json;d=json.load(open('gpurun_out/exp_$name.json'));print(round(d['value']),'traj/s',round(d['ms_per_step'],2),'ms', 'e2e',round(d['e2e']['value']))" 2>&1 | tail -1)"
  grep "iteration total" gpurun_out/exp_$name.err
done
if [ -n "$TRACE" ]; then
  ( for kv in ${TRACE_ENV//,/ }; do export "$kv"; done; python scripts/trace_tc.py $TRACE > gpurun_out/trace.txt 2>&1 ); tail -3 gpurun_out/trace.txt
fi
