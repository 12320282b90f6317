#!/bin/bash
TAG=${1:-r2u}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest staging + topk"; timeout 1200 python -m pytest tests/test_gpu_staging.py tests/test_gpu_topk_fused.py -q -x -m gpu --timeout=600 2>&1 | tail -15 | tee $OUT/tests.txt
echo "== bench default"; timeout 1500 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; tail -c 600 $OUT/bench.err; python - <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1] if len(sys.argv)>1 else 'gpurun_out/r2u/bench.json').read().strip().split('\n')[-1])
except Exception as e:
    print('no json', e); sys.exit(0)
print({k: j[k] for k in ('value','ms_per_step','gpu_launches') if k in j}); print('e2e', j.get('e2e',{}).get('value')); print('roofline', {k:j['roofline'][k] for k in ('achieved','frac','avg_launch_ms','share_of_step')})
print('infer_topk', j.get('infer_topk')); print('staging', json.dumps(j.get('staging'), indent=1)[:3000]); print('cpu', j.get('cpu_baseline'))
for r in j.get('infer_topk_sweep', {}).get('rows', [])[:60] if isinstance(j.get('infer_topk_sweep'), dict) else []: print(r)
PY
