#!/bin/bash
# full-set ncu capture of selected tensor-core conv launches (direct launches, no graph); reports stay small
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
export CDS_GRAPH=0
SPECS=${NCU_SPECS:-40:1:first,57:2:mid}
for spec in ${SPECS//,/ }; do
  skip=$(echo "$spec" | cut -d: -f1); count=$(echo "$spec" | cut -d: -f2); tag=$(echo "$spec" | cut -d: -f3)
  echo "== capture $tag: skip $skip count $count"
  if [ -n "$DRY" ]; then continue; fi
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s "$skip" -c "$count" \
     -o "gpurun_out/prof_tc_$tag" -f python scripts/one_sample.py bf16 3 > "gpurun_out/ncu_tc_$tag.log" 2>&1
  echo "ncu $tag exit $?"; tail -n 2 "gpurun_out/ncu_tc_$tag.log"
  ncu -i "gpurun_out/prof_tc_$tag.ncu-rep" --page details > "gpurun_out/prof_tc_$tag.details.txt" 2>&1
done
ls -la gpurun_out/
