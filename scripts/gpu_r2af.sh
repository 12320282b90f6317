#!/bin/bash
OUT=gpurun_out/r2af; mkdir -p $OUT
echo "== pytest topk"; timeout 1200 python -m pytest tests/test_gpu_topk_fused.py tests/test_gpu_e2e.py tests/test_gpu_shard.py -q -x -m gpu --timeout=600 2>&1 | tail -8
for a in "1000 1000" "1000 4096" "1000 512" "100 1000" "10 1000"; do python scripts/topk_prof.py $a 2>&1 | grep "fused=" | tee -a $OUT/topk_prof.txt; done
bash scripts/topk_stages.sh 1000 1000 2>&1 | tee $OUT/stages_k1000.txt
