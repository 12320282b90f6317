#!/bin/bash
TAG=${1:-r2g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== tc2 timing"; timeout 600 python scripts/tc2_timing.py 2>&1 | head -48 | tee $OUT/tc2_timeline.txt
echo "== pytest gpu (tc, shard, graph, topk)"; timeout 1500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_shard.py tests/test_gpu_graph.py tests/test_gpu_topk_fused.py tests/test_gpu_e2e.py tests/test_gpu_dp.py -q -x --timeout=600 2>&1 | tail -8 | tee $OUT/tests.txt
echo "== bench v2"; timeout 900 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-extras 2>&1 | tail -1 | tee $OUT/bench_v2.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['avg_launch_ms'],d['roofline']['frac'],'infer',d['infer_topk']['value'],d['clocks'])"
echo "== topk prof"; for K in 10; do timeout 300 python scripts/topk_prof.py $K 2>&1 | tail -2 | tee -a $OUT/topk_prof.txt; done
