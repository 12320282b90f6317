#!/bin/bash
TAG=${1:-r2w}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== tc tests (staggered)"; timeout 900 python -m pytest tests/test_gpu_tc.py -q -x -m gpu --timeout=600 2>&1 | tail -4
echo "== tc2 span staggered"; timeout 300 python scripts/tc2_timing.py 2>&1 | grep "^==" | tee $OUT/tc2_span.txt
echo "== tc2 span all-16"; NTF_TC2_STAG=0 timeout 300 python scripts/tc2_timing.py 2>&1 | grep "^==" | tee -a $OUT/tc2_span.txt
echo "== bench"; timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | tee $OUT/bench.json | cut -c1-300
timeout 300 python scripts/tc2_timing.py 2>&1 | head -60 > $OUT/tc2_timeline.txt
