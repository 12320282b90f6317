#!/bin/bash
# top-K kernels after the radix threshold / warp final / two issuing warps change: tests, event timings, launch list, CTA timeline
TAG=${1:-r2q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest topk + shard + tc"; timeout 1200 python -m pytest tests/test_gpu_topk_fused.py tests/test_gpu_tc.py -q -x -m gpu --timeout=600 2>&1 | tail -5 | tee $OUT/tests.txt
for K in 10 100; do timeout 300 python scripts/topk_prof.py $K 2>&1 | tail -2 | tee -a $OUT/topk_prof.txt; done
timeout 300 python scripts/topk_prof.py 10 4096 2>&1 | tail -2 | tee -a $OUT/topk_prof.txt
echo "== ncu launch list topk"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/topk_launches.csv python scripts/topk_prof.py 10 > $OUT/topk_under_ncu.log 2>&1
python scripts/summarize_launches.py $OUT/topk_launches.csv 2>&1 | tail -30 | tee $OUT/topk_launches_summary.txt
echo "== stages"; bash scripts/topk_stages.sh 10 1000 2>&1 | tee $OUT/topk_stages.txt; bash scripts/topk_stages.sh 100 1000 2>&1 | tee -a $OUT/topk_stages.txt
echo "== topk timing"; timeout 300 python scripts/topk_timing.py 2>&1 | head -100 | tee $OUT/topk_timeline.txt
echo "== ncu full (topk chain)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"blockmax_threshold|infer_topk" -s 8 -c 4 -o $OUT/prof_topk -f python scripts/topk_prof.py 10 > $OUT/ncu_topk.log 2>&1; ls -la $OUT | tail -3
