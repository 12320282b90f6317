#!/bin/bash
# Flipout tensor-core bug hunt: flip_stress3 under several experiment bits.   usage: bash scripts/gpu_flip3.sh [tag]
TAG=${1:-flip3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee $OUT/gpu.txt
for e in 0 32 16 48; do
  echo "== NTF_TC_EXP=$e"; NTF_TC_EXP=$e timeout 300 python scripts/flip_stress3.py 40 2>&1 | tail -40 | tee $OUT/stress3_exp$e.txt
done
echo "== mb_reduce"; timeout 120 ./scripts/mb/mb_reduce 2>&1 | tee $OUT/mb_reduce.txt
echo "== fused top-K tests"; timeout 600 python -m pytest tests/test_gpu_topk_fused.py -q -x --timeout=300 2>&1 | tail -25 | tee $OUT/topk_fused_tests.txt
echo "== bench (short)"; timeout 900 python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench.json | cut -c1-1500
