#!/bin/bash
TAG=${1:-r2c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
bash scripts/flip_variants.sh run 40 2>&1 | tee $OUT/flip_variants.txt
echo "== flip_stress (200)"; timeout 900 python scripts/flip_stress.py 200 2>&1 | tail -4 | cut -c1-300 | tee $OUT/flip_stress200.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --timeout=600 2>&1 | tail -25 | tee $OUT/gpu_tests.txt
echo "== topk prof"; for K in 10 100; do timeout 300 python scripts/topk_prof.py $K 2>&1 | tail -2 | tee -a $OUT/topk_prof.txt; done
echo "== ncu launch list (topk)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/topk_launches.csv python scripts/topk_prof.py 10 > $OUT/topk_under_ncu.log 2>&1
python scripts/summarize_launches.py $OUT/topk_launches.csv 2>&1 | tail -25 | tee $OUT/topk_launches_summary.txt
