#!/bin/bash
# quick GPU loop: all GPU parity tests in one process (as the driver runs them) + one bench line.   usage: bash scripts/gpu_quick.sh [tag] [extra bench args]
TAG=${1:-quick}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -q -m gpu -x --timeout=120 2>&1 | tail -15 | tee $OUT/test_gpu_all.log
echo "== bench"; timeout 600 python bench.py --steps 1000 --warmup 10 --no-cpu-baseline "$@" 2>&1 | tail -1 | tee $OUT/bench.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['avg_launch_ms'],d['roofline']['frac'],'infer',d['infer_topk']['value'],d['clocks'])"
