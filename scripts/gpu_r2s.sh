#!/bin/bash
TAG=${1:-r2s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest topk"; timeout 1200 python -m pytest tests/test_gpu_topk_fused.py -q -x -m gpu --timeout=600 2>&1 | tail -3 | tee $OUT/tests.txt
echo "== stages"; bash scripts/topk_stages.sh 10 1000 2>&1 | tee $OUT/topk_stages.txt
echo "== timing"; python scripts/topk_timing.py 2>&1 | grep -A1 "^pass" | tee $OUT/topk_span.txt
