#!/bin/bash
# One gpurun call: GPU parity tests (one process per file so that a trapped kernel cannot poison the others), smoke, a
# short bench of both precisions, and the ncu launch list of the bench command.
# usage (from the repo root on the GPU box):  bash scripts/gpu_check.sh [tag] [quick]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
for f in test_gpu_kernels test_gpu_e2e test_gpu_tc test_gpu_bnn test_gpu_shard; do
  echo "== pytest $f"; timeout 900 python -m pytest tests/$f.py -q -m gpu 2>&1 | tail -60 > $OUT/$f.log; tail -25 $OUT/$f.log
done
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== bench fp32"; timeout 900 python bench.py --steps 100 --warmup 5 --precision fp32 --no-cpu-baseline 2>&1 | tail -3 | tee $OUT/bench_fp32.json
echo "== bench tf32"; timeout 900 python bench.py --steps 1000 --warmup 10 2>&1 | tail -3 | tee $OUT/bench_tf32.json
if [ "$2" != "quick" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1; tail -30 $OUT/launches_summary.txt
fi
