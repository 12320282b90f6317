#!/bin/bash
# usage: bash scripts/gpu_dp2.sh [tag] [ngpus]   -- NCCL check of the public classes with full logs + the dense-input tests
TAG=${1:-dp2}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest dense/dp"; timeout 600 python -m pytest tests/test_gpu_dense_input.py tests/test_gpu_dp.py -q --timeout=180 2>&1 | tail -30 | tee $OUT/dense_tests.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== nccl check"; timeout 600 $TR scripts/shard_nccl_check.py > $OUT/nccl_check_full.txt 2>&1; grep -v "^W1\|^\*\*\*" $OUT/nccl_check_full.txt | grep -B30 -A3 "Error\|error\|OK\|max |" | tail -80
