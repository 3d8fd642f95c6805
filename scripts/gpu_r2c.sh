#!/bin/bash
# Round 2, visit C: full GPU suite on the current tree; ncu capture of the TF32 L=8 and L=32 conv_tc kernels WITH the source page.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu_c.log 2>&1
echo "[tests] exit $?"; tail -12 gpurun_out/pytest_gpu_c.log
export CDS_GRAPH=0
for spec in 52:tf32_L8 42:tf32_L32; do
  skip=${spec%%:*}; tag=${spec##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'conv_(tc|ps)_kernel' -s $skip -c 1 -o gpurun_out/prof_$tag -f python scripts/one_sample.py tf32 3 > gpurun_out/ncu_$tag.log 2>&1
  echo "ncu $tag exit $?"
  ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/prof_$tag.source.csv 2>&1
  gzip -f gpurun_out/prof_$tag.source.csv
  rm -f gpurun_out/prof_$tag.ncu-rep
done
ls -la gpurun_out | tail -8
