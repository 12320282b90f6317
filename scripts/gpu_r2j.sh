#!/bin/bash
TAG=${1:-r2j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -x -m gpu --timeout=600 2>&1 | tail -8 | tee $OUT/tests.txt
echo "== tc2 timing"; timeout 600 python scripts/tc2_timing.py 2>&1 | head -24 | tee $OUT/tc2_timeline.txt
echo "== bench"; timeout 900 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-extras 2>&1 | tail -1 | tee $OUT/bench.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['avg_launch_ms'],d['roofline']['frac'],'infer',d['infer_topk']['value'],d['clocks'])"
echo "== ncu launch list (bench)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_under_ncu.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv 2>&1 | tail -30 | tee $OUT/launches_summary.txt
