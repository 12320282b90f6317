#!/bin/bash
TAG=${1:-r2v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== staging tests"; timeout 900 python -m pytest tests/test_gpu_staging.py -q -x -m gpu --timeout=600 2>&1 | tail -3
echo "== tc2 span, normal"; python scripts/tc2_timing.py 2>&1 | grep "^==" | tee $OUT/tc2_span.txt
echo "== tc2 span, no dA atomics"; NTF_TC2_EXP=1 python scripts/tc2_timing.py 2>&1 | grep "^==" | tee -a $OUT/tc2_span.txt
echo "== bench staging only"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().split('\n')[-1])
print(j['value'], j['ms_per_step']); s=j.get('staging',{})
for k,v in s.items(): print(k, {a:v[a] for a in ('ms','value','unit','frac') if a in v}, v.get('cpu_baseline',{}).get('value')) if isinstance(v,dict) else print(k,v)
"
