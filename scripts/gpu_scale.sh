#!/bin/bash
# one N-GPU data point (gpurun --gpus N): the exchange micro-benchmark, the Fnn bench with each requested gradient-exchange mode, the Bnn
# leg, and the multi-GPU check of the public classes.   usage: bash scripts/gpu_scale.sh [tag] [ngpus] [modes] [nocheck]
TAG=${1:-scale}; N=${2:-8}; MODES=${3:-"peer"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== peer_bench N=$N"; timeout 300 $TR scripts/peer_bench.py 2>&1 | grep -v "^W1\|^\*\*\*" | tail -1 | tee $OUT/peer_bench_n$N.json
for m in $MODES; do
  echo "== bench N=$N exchange=$m"; NTF_DP_EXCHANGE=$m timeout 600 $TR bench.py --gpus $N --steps 500 --warmup 10 --no-extras --no-cpu-baseline > $OUT/bench_n${N}_$m.log 2>&1; tail -1 $OUT/bench_n${N}_$m.log | tee $OUT/bench_n${N}_$m.json | cut -c1-600
done
echo "== bench bnn N=$N"; timeout 600 $TR bench.py --gpus $N --leg bnn --steps 100 --warmup 10 2>&1 | tail -1 | tee $OUT/bench_bnn_n${N}.json | cut -c1-500
if [ "$4" != "nocheck" ]; then
echo "== multi gpu check N=$N"; timeout 600 $TR scripts/multi_gpu_check.py 2>&1 | grep -v "^W1\|^\*\*\*\|UserWarning\|sparse_coo" | tail -45 > $OUT/multi_gpu_check_n${N}.txt; tail -3 $OUT/multi_gpu_check_n${N}.txt
fi
