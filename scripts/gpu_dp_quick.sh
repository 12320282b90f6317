#!/bin/bash
# usage: bash scripts/gpu_dp_quick.sh [tag] [ngpus] -- kernel/graph/dp tests, multi-GPU check, exchange micro-benchmark, bench at N=1 and N
TAG=${1:-dpq}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_graph.py tests/test_gpu_dp.py tests/test_gpu_e2e.py -q --timeout=300 --tb=short 2>&1 | tail -12 | tee $OUT/tests.txt
echo "== peer_bench N=$N"; timeout 300 $TR scripts/peer_bench.py 2>&1 | grep -v "^W1\|^\*\*\*" | tail -1 | tee $OUT/peer_bench_n$N.json
echo "== bench N=1"; timeout 600 python bench.py --steps 500 --warmup 10 --no-extras --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_n1.json | cut -c1-330
echo "== bench N=$N"; timeout 600 $TR bench.py --gpus $N --steps 500 --warmup 10 --no-extras --no-cpu-baseline > $OUT/bench_n${N}_peer.log 2>&1; tail -1 $OUT/bench_n${N}_peer.log | tee $OUT/bench_n${N}_peer.json | cut -c1-330
