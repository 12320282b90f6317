#!/bin/bash
# A/B of the step's optimiser arrangement on ONE box: rows-split Adam on/off x deferred reductions on their own stream / in front of the last layer's optimiser
TAG=${1:-r2z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { env "$@" timeout 600 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('%.2f us/step  %.3f M teams/s  e2e %.3f M' % (j['ms_per_step']*1e3, j['value']/1e6, j['e2e']['value']/1e6))"; }
for rep in 1 2; do
echo -n "rows-split + side2:   "; run A=1
echo -n "flat       + side2:   "; run NTF_ADAM_ROWS_OFF=1
echo -n "rows-split + side0:   "; run NTF_FINISH_SIDE0=1
echo -n "flat       + side0:   "; run NTF_ADAM_ROWS_OFF=1 NTF_FINISH_SIDE0=1
done | tee $OUT/ab.txt
