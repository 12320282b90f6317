#!/bin/bash
# 8-GPU run of round 2: multi-GPU checks through the public classes, the expert-sharded layer at the USPT shape (BASELINE configs[3]) against the
# unsharded one, the sharded bench at 8 GPUs, the data-parallel bench at 8 GPUs.   usage: gpurun --gpus 8 -- bash scripts/gpu_r2_n8.sh [tag]
TAG=${1:-r2_n8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
N=${NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== multi_gpu_check"; timeout 600 $TR --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep -v "^W\|UserWarning\|warnings.warn" | tail -30 | tee $OUT/multi_gpu_check_n$N.txt
echo "== shard C4 check"; timeout 900 $TR --master-port 29512 scripts/shard_c4_check.py 2>&1 | grep -v "^W\|UserWarning\|warnings.warn" | tail -8 | tee $OUT/shard_c4_check_n$N.txt
echo "== bench uspt shard N=$N"; timeout 900 $TR --master-port 29513 bench.py --gpus $N --workload uspt --parallel shard --steps 100 --warmup 5 --no-cpu-baseline --no-extras 2>&1 | tail -1 | tee $OUT/bench_uspt_shard_n$N.json | cut -c1-600
echo "== bench dblp dp N=$N"; timeout 900 $TR --master-port 29514 bench.py --gpus $N --steps 300 --warmup 10 --no-cpu-baseline --extras 2>&1 | tail -1 | tee $OUT/bench_dblp_dp_n$N.json | cut -c1-600
echo "== bench bnn dp N=$N"; timeout 900 $TR --master-port 29515 bench.py --gpus $N --leg bnn --steps 100 --warmup 5 2>&1 | tail -1 | tee $OUT/bench_bnn_n$N.json | cut -c1-600
echo "== bench dblp dp N=$N (driver command)"; timeout 900 $TR --master-port 29516 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | tail -1 | tee $OUT/bench_dblp_dp_n${N}_driver.json | cut -c1-400
