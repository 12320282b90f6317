#!/bin/bash
# usage: bash scripts/gpu_n4.sh [tag] [ngpus]: exchange micro-benchmark + bench (Fnn default exchange, Bnn leg) at N GPUs
TAG=${1:-n4}; N=${2:-4}
OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== peer_bench N=$N"; timeout 300 $TR scripts/peer_bench.py 2>&1 | grep -v "^W1\|^\*\*\*" | tail -3 | tee $OUT/peer_bench_n$N.json
echo "== bench N=$N"; timeout 600 $TR bench.py --gpus $N --steps 500 --warmup 10 --no-extras --no-cpu-baseline > $OUT/bench_n${N}_peer.log 2>&1; tail -1 $OUT/bench_n${N}_peer.log | tee $OUT/bench_n${N}_peer.json | cut -c1-330
echo "== bench bnn N=$N"; timeout 600 $TR bench.py --gpus $N --leg bnn --steps 100 --warmup 10 2>&1 | tail -1 | tee $OUT/bench_bnn_n${N}.json | cut -c1-500
