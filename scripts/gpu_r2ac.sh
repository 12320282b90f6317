#!/bin/bash
run() { env "$@" timeout 600 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('%.2f us/step  %.3f M teams/s  e2e %.3f M  out-layer %.1f us frac %.3f' % (j['ms_per_step']*1e3, j['value']/1e6, j['e2e']['value']/1e6, j['roofline']['avg_launch_ms']*1e3, j['roofline']['frac']))"; }
for rep in 1 2; do
echo -n "early pass 1 block/SM:  "; run NTF_ADAM_EARLY_BLOCKS=1
echo -n "early pass 2 blocks/SM: "; run NTF_ADAM_EARLY_BLOCKS=2
echo -n "early pass 16 blocks/SM:"; run NTF_ADAM_EARLY_BLOCKS=16
echo -n "no row splits:          "; run NTF_ADAM_ROWS_OFF=1
done
