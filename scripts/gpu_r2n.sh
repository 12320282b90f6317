#!/bin/bash
TAG=${1:-r2n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for e in 0 1 2 3 4 7; do
  NTF_FIX_EXP=$e timeout 300 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none -k regex:"out_fix_kernel|out_tc2_kernel" --csv --log-file $OUT/fix_exp$e.csv python scripts/fix_timing.py > /dev/null 2>&1
  echo "== NTF_FIX_EXP=$e (warm caches)"; python scripts/summarize_launches.py $OUT/fix_exp$e.csv 2>&1 | tail -3
done
NTF_FIX_EXP=0 timeout 300 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none -k regex:"out_fix_kernel|out_tc2_kernel" --csv --log-file $OUT/fix_hot.csv python scripts/fix_timing.py zipf > /dev/null 2>&1
echo "== hot expert"; python scripts/summarize_launches.py $OUT/fix_hot.csv 2>&1 | tail -3
echo "== topk timing"; timeout 300 python scripts/topk_timing.py 2>&1 | head -60 | tee $OUT/topk_timeline.txt
