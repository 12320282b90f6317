#!/bin/bash
# round-end run on ONE B200: GPU suite, smoke(), the driver's two bench commands, a 300-step bench, launch lists (Fnn step, Bnn step, top-K chain)
TAG=${1:-r2final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv | tee $OUT/gpu.txt
echo "== pytest gpu all"; timeout 1800 python -m pytest tests -q -x -m gpu --timeout=900 2>&1 | tail -4 | tee $OUT/tests.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench reference (driver command)"; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 > $OUT/bench_reference.json; cut -c1-200 $OUT/bench_reference.json
echo "== bench default (driver command)"; timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 2>$OUT/bench.err | tail -1 > $OUT/bench.json; cut -c1-300 $OUT/bench.json
echo "== bench 300 steps"; timeout 900 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > $OUT/bench_300.json; cut -c1-300 $OUT/bench_300.json
echo "== ncu launch list (Fnn step)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_under_ncu.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv 2>&1 | tail -32 | head -22 | tee $OUT/launches_summary.txt
echo "== ncu launch list (top-K chain)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/topk_launches.csv python scripts/topk_prof.py 10 > $OUT/topk_under_ncu.log 2>&1
python scripts/summarize_launches.py $OUT/topk_launches.csv 2>&1 | tail -20 | head -12 | tee $OUT/topk_launches_summary.txt
echo "== topk stages"; bash scripts/topk_stages.sh 10 1000 2>&1 | tee $OUT/topk_stages.txt
echo "== bnn"; timeout 600 python bench.py --leg bnn --steps 100 --warmup 5 2>&1 | tail -1 | tee $OUT/bench_bnn.json | cut -c1-400
echo "== uspt shard N=1"; timeout 900 python bench.py --workload uspt --parallel shard --steps 50 --warmup 5 --no-cpu-baseline --no-extras 2>&1 | tail -1 | tee $OUT/bench_uspt_shard_n1.json | cut -c1-300
