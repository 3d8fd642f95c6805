#!/bin/bash
# full-set ncu capture of selected tensor-core conv launches (direct launches, no graph).  The .ncu-rep files are too big
# to travel back (gpurun_out/ is capped at 64 MiB), so the pages we read are exported on the box and the reports deleted.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
export CDS_GRAPH=0
SPECS=${NCU_SPECS:-40:1:first,57:2:mid}
KREGEX=${NCU_KERNEL:-conv_tc_kernel}
for spec in ${SPECS//,/ }; do
  skip=$(echo "$spec" | cut -d: -f1); count=$(echo "$spec" | cut -d: -f2); tag=$(echo "$spec" | cut -d: -f3)
  echo "== capture $tag: skip $skip count $count"
  if [ -n "$DRY" ]; then continue; fi
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s "$skip" -c "$count" \
     -o "gpurun_out/prof_tc_$tag" -f python scripts/one_sample.py ${NCU_MATH:-bf16} 3 > "gpurun_out/ncu_tc_$tag.log" 2>&1
  echo "ncu $tag exit $?"; tail -n 2 "gpurun_out/ncu_tc_$tag.log"
  ncu -i "gpurun_out/prof_tc_$tag.ncu-rep" --page details > "gpurun_out/prof_tc_$tag.details.txt" 2>&1
  ncu -i "gpurun_out/prof_tc_$tag.ncu-rep" --page raw --csv > "gpurun_out/prof_tc_$tag.raw.csv" 2>&1
  ncu -i "gpurun_out/prof_tc_$tag.ncu-rep" --page source --csv > "gpurun_out/prof_tc_$tag.source.csv" 2>&1
  rm -f "gpurun_out/prof_tc_$tag.ncu-rep"
done
du -sh gpurun_out; ls -la gpurun_out/
