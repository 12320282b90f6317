#!/bin/bash
# all GPU tests in one process (as the driver runs them) + bench with and without CUDA graphs.   usage: bash scripts/gpu_graph.sh [tag]
TAG=${1:-graph}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 420 python -m pytest tests -q -m gpu -x --timeout=90 2>&1 | tail -30 | tee $OUT/test_gpu_all.log
echo "== bench graphs"; timeout 600 python bench.py --steps 1000 --warmup 10 --no-cpu-baseline 2>&1 | tail -3 | tee $OUT/bench_graph.json
echo "== bench no graphs"; NTF_GRAPHS=0 timeout 600 python bench.py --steps 1000 --warmup 10 --no-cpu-baseline 2>&1 | tail -3 | tee $OUT/bench_nograph.json
