#!/bin/bash
TAG=${1:-r2y}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu all"; timeout 1800 python -m pytest tests -q -x -m gpu --timeout=900 2>&1 | tail -4 | tee $OUT/tests.txt
echo "== bench rows-split"; timeout 600 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | tee $OUT/bench.json | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['gpu_launches'])"
echo "== bench L1 flat"; NTF_ADAM_L1_FLAT=1 timeout 600 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['gpu_launches'])"
echo "== bench all flat"; NTF_ADAM_ROWS_OFF=1 timeout 600 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['gpu_launches'])"
