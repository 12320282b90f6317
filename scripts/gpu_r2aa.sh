#!/bin/bash
TAG=${1:-r2aa}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest e2e + graph"; timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_graph.py -q -x -m gpu --timeout=600 2>&1 | tail -3
for r in 1 2; do timeout 600 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('%.2f us/step  %.3f M teams/s  e2e %.3f M  host enqueue %.1f us' % (j['ms_per_step']*1e3, j['value']/1e6, j['e2e']['value']/1e6, j['host_enqueue_ms_per_step']*1e3))"; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('driver-style 20 steps: %.2f us/step  %.3f M teams/s  e2e %.3f M' % (j['ms_per_step']*1e3, j['value']/1e6, j['e2e']['value']/1e6))"
