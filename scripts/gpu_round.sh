#!/bin/bash
# One gpurun call that mirrors the driver's round-end sequence: all GPU parity tests in one process, smoke, both bench arms,
# the ncu launch list of the bench command.   usage: bash scripts/gpu_round.sh [tag] [noncu]
TAG=${1:-round}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --timeout=180 2>&1 | tail -40 | tee $OUT/gpu_tests.txt | tail -12
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== bench (default)"; timeout 900 python bench.py 2>&1 | tail -1 | tee $OUT/bench_tf32.json
echo "== bench reference arm"; timeout 900 python bench.py --impl reference 2>&1 | tail -1 | tee $OUT/bench_reference.json
if [ "$2" != "noncu" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1; tail -30 $OUT/launches_summary.txt
fi
