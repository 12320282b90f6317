#!/bin/bash
# quick loop for the tcgen05 kernel: its parity tests, tile stamps, bench lines.   usage: bash scripts/gpu_tc.sh [tag] [exp list]
TAG=${1:-tc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | tail -n 30 | tee $OUT/test_gpu_tc.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
timeout 300 python scripts/tc_timing.py > $OUT/tc_timing.txt 2>&1; grep -A 12 "train exp 0" $OUT/tc_timing.txt
for e in ${2:-0}; do
NTF_TC_EXP=$e timeout 600 python bench.py --steps 500 --warmup 10 --no-cpu-baseline 2>&1 | tail -n 1 | tee $OUT/bench_exp$e.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('exp $e teams/s', d['value'], 'ms/step', d['ms_per_step'], 'host enqueue ms', d['host_enqueue_ms_per_step'], 'out_tc ms', d['roofline']['avg_launch_ms'], 'e2e', d['e2e']['value'], 'infer', d['infer_topk']['value'])"
done
