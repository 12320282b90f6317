#!/bin/bash
TAG=${1:-r2p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu all"; timeout 1800 python -m pytest tests -q -x -m gpu --timeout=900 2>&1 | tail -5 | tee $OUT/tests.txt
echo "== flip stress 100"; timeout 600 python scripts/flip_stress.py 100 2>&1 | tail -2 | cut -c1-250 | tee $OUT/flip_stress.txt
echo "== bench bnn"; timeout 600 python bench.py --leg bnn --steps 100 --warmup 5 2>&1 | tail -1 | tee $OUT/bench_bnn.json | cut -c1-700
echo "== ncu launch list (bnn)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file $OUT/bnn_launches.csv python bench.py --leg bnn --steps 10 --warmup 3 > $OUT/bnn_under_ncu.log 2>&1
python scripts/summarize_launches.py $OUT/bnn_launches.csv 2>&1 | tail -40 | tee $OUT/bnn_launches_summary.txt
echo "== topk timing"; timeout 300 python scripts/topk_timing.py 2>&1 | head -45 | tee $OUT/topk_timeline.txt
echo "== bench uspt shard N=1"; timeout 900 python bench.py --workload uspt --parallel shard --steps 50 --warmup 5 --no-cpu-baseline --no-extras 2>&1 | tail -1 | tee $OUT/bench_uspt_shard_n1.json | cut -c1-500
