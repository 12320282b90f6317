#!/bin/bash
# N-GPU check (gpurun --gpus N): data-parallel and expert-sharded modes over NCCL against the single-GPU run, then the bench at N=1 and N
# (gradient exchange inside the step vs through torch.distributed).   usage: bash scripts/gpu_dp.sh [tag] [ngpus]
TAG=${1:-dp}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/gpu.txt 2>&1
nvidia-smi topo -m >> $OUT/gpu.txt 2>&1
echo "== pytest dp/dense"; timeout 600 python -m pytest tests/test_gpu_dp.py tests/test_gpu_dense_input.py -q --timeout=180 2>&1 | tail -15 | tee $OUT/dp_tests.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== nccl check"; timeout 600 $TR scripts/shard_nccl_check.py > $OUT/nccl_check_full.txt 2>&1; grep -v "^W1\|^\*\*\*" $OUT/nccl_check_full.txt | grep -B25 -A3 "Error\|error\|OK\|max |" | tail -60
echo "== bench N=1"; timeout 600 python bench.py --steps 500 --warmup 10 --no-extras --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_n1.json
echo "== bench N=$N in-step exchange"; timeout 600 $TR bench.py --gpus $N --steps 500 --warmup 10 --no-extras --no-cpu-baseline > $OUT/bench_n${N}_instep.log 2>&1; tail -1 $OUT/bench_n${N}_instep.log | tee $OUT/bench_n${N}_instep.json
echo "== bench N=$N torch exchange"; NTF_DP_NCCL=0 timeout 600 $TR bench.py --gpus $N --steps 500 --warmup 10 --no-extras --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_n${N}_torch.json
