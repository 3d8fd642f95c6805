#!/bin/bash
# Round 2, visit T: run_range split test; cfg2 tf32 knob sweep (branches, split policy, graph length), 8 timed steps each.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q -k "run_range_splits or pearce or sfbc or dvinv" --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_t.log 2>&1
echo "[tests] exit $?"; tail -4 gpurun_out/pytest_t.log; grep -E "^E " gpurun_out/pytest_t.log | head
for knob in "" "CDS_BRANCHES=2" "CDS_TC_NOSPLIT=1" "CDS_TC_NOSPLIT=2" "CDS_GRAPH_ITERS=25" "CDS_GRAPH_ITERS=50" ""; do
  env $knob timeout 400 python bench.py --math tf32 --steps 8 --warmup 3 --no-cpu-baseline --no-other-configs --no-eager-baseline > gpurun_out/bench_tf32_knob.json 2> gpurun_out/bench_tf32_knob.err
  echo "[bench tf32 '$knob'] exit $?"; grep -E "timed:" gpurun_out/bench_tf32_knob.err
done
