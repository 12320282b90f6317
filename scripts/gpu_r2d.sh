#!/bin/bash
TAG=${1:-r2d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest tc"; timeout 1500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_e2e.py tests/test_gpu_graph.py tests/test_gpu_shard.py -q -x --timeout=600 2>&1 | tail -25 | tee $OUT/tc_tests.txt
echo "== bench v2"; timeout 900 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-extras 2>&1 | tail -1 | tee $OUT/bench_v2.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['avg_launch_ms'],d['roofline']['frac'],'infer',d['infer_topk']['value'],d['clocks'])"
echo "== bench v1"; NTF_TC_V1=1 timeout 900 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-extras 2>&1 | tail -1 | tee $OUT/bench_v1.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['avg_launch_ms'],d['roofline']['frac'],'infer',d['infer_topk']['value'],d['clocks'])"
