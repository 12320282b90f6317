#!/bin/bash
# full single-GPU round-end rehearsal: the GPU suite, smoke(), the default bench line, the reference arm
TAG=${1:-r2x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu all"; timeout 1800 python -m pytest tests -q -x -m gpu --timeout=900 2>&1 | tail -4 | tee $OUT/tests.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench default"; timeout 1500 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; tail -c 300 $OUT/bench.err
python - $OUT/bench.json <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
print({k: j[k] for k in ('value','ms_per_step','gpu_launches') if k in j}); print('e2e', j.get('e2e',{}).get('value')); print('roofline', {k:j['roofline'][k] for k in ('achieved','frac','avg_launch_ms','share_of_step')})
print('infer_topk', j.get('infer_topk')); print('cpu', j.get('cpu_baseline',{}).get('value')); print('bnn', {k:j.get('bnn_train',{}).get(k) for k in ('value','ms_per_step')})
for r in j.get('kernel_rooflines', []): print(r)
sw=j.get('infer_topk_sweep')
if isinstance(sw, list):
    for r in sw: print(r)
else: print(sw)
PY
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tail -1 | tee $OUT/bench_reference.json | cut -c1-400
