#!/bin/bash
# Round 2, visit F: parity suite on the current tree + the full default bench line (headline tf32, other configs, eager baseline, CPU arm).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -x > gpurun_out/pytest_gpu_f.log 2>&1
echo "[tests] exit $?"; tail -5 gpurun_out/pytest_gpu_f.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "[bench] exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','dtype','loop_ms_per_step','gpu_launches')})
print('e2e',d['e2e']); print('roofline frac',d['roofline']['frac'],'tensor',d['roofline']['tensor']); print('eager',d['gpu_eager_baseline']); print('cpu',d['cpu_baseline'])
for k,v in (d['other_configs'] or {}).items(): print(k, {kk:v.get(kk) for kk in ('value','ms_per_call','tflops_per_gpu','error')}, v.get('roofline',{}).get('frac'))
PY
grep -E "timed:|e2e:|cfg[345]|eager" gpurun_out/bench_full.err
