#!/bin/bash
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for e in 0 32 16; do
  echo "== stress2 NTF_TC_EXP=$e"; NTF_TC_EXP=$e timeout 300 python scripts/flip_stress2.py 40 2>&1 | tail -12 | tee $OUT/stress2_exp$e.txt
done
echo "== flip_stress"; timeout 300 python scripts/flip_stress.py 30 2>&1 | tail -8 | tee $OUT/flip_stress.txt
echo "== fused top-K tests"; timeout 600 python -m pytest tests/test_gpu_topk_fused.py -q --timeout=300 2>&1 | tail -30 | tee $OUT/topk_fused_tests.txt
echo "== topk prof"; for K in 10 100; do timeout 300 python scripts/topk_prof.py $K 2>&1 | tail -2 | tee -a $OUT/topk_prof.txt; done
echo "== ncu launch list (topk)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/topk_launches.csv python scripts/topk_prof.py 10 > $OUT/topk_under_ncu.log 2>&1
python scripts/summarize_launches.py $OUT/topk_launches.csv 2>&1 | tail -25 | tee $OUT/topk_launches_summary.txt
