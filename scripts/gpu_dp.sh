#!/bin/bash
# N-GPU check (gpurun --gpus N): data-parallel and expert-sharded modes against the single-GPU run through the public classes, then the
# bench at N=1 and N with each gradient-exchange mode.   usage: bash scripts/gpu_dp.sh [tag] [ngpus] [modes]
TAG=${1:-dp}; N=${2:-2}; MODES=${3:-"peer nccl torch"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/gpu.txt 2>&1
nvidia-smi topo -m >> $OUT/gpu.txt 2>&1
echo "== pytest dp"; timeout 900 python -m pytest tests/test_gpu_dp.py -q --timeout=240 2>&1 | tail -15 | tee $OUT/dp_tests.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== check of the public classes at N=$N"; timeout 600 $TR scripts/multi_gpu_check.py > $OUT/nccl_check_full.txt 2>&1; grep -v "^W1\|^\*\*\*" $OUT/nccl_check_full.txt | grep -B25 -A3 "Error\|error\|OK\|max |" | tail -60
echo "== bench N=1"; timeout 600 python bench.py --steps 500 --warmup 10 --no-extras --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_n1.json
for m in $MODES; do
  echo "== bench N=$N exchange=$m"; NTF_DP_EXCHANGE=$m timeout 600 $TR bench.py --gpus $N --steps 500 --warmup 10 --no-extras --no-cpu-baseline > $OUT/bench_n${N}_$m.log 2>&1; tail -1 $OUT/bench_n${N}_$m.log | tee $OUT/bench_n${N}_$m.json
done
echo "== bench bnn N=$N"; timeout 600 $TR bench.py --gpus $N --leg bnn --steps 100 --warmup 10 2>&1 | tail -1 | tee $OUT/bench_bnn_n${N}.json
