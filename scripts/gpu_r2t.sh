#!/bin/bash
TAG=${1:-r2t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu all"; timeout 1800 python -m pytest tests -q -x -m gpu --timeout=900 2>&1 | tail -5 | tee $OUT/tests.txt
echo "== stages"; bash scripts/topk_stages.sh 10 1000 2>&1 | tee $OUT/topk_stages.txt; bash scripts/topk_stages.sh 100 1000 2>&1 | tee -a $OUT/topk_stages.txt
for a in "1 1000" "10 4096" "10 128" "10 32768" "100 4096"; do python scripts/topk_prof.py $a 2>&1 | grep "fused=1" | tee -a $OUT/topk_prof.txt; done
