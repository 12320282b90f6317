#!/bin/bash
# experiments on the step: tile stamps of the tcgen05 kernel + bench lines under env switches.  usage: bash scripts/gpu_exp.sh tag "ENV1=a ENV2=b" "ENV3=c" ...
TAG=${1:-exp}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== stamps"; timeout 300 python scripts/tc_timing.py > $OUT/tc_timing.txt 2>&1; sed -n '/== train exp 0/,/== train exp 8/p' $OUT/tc_timing.txt | head -16
i=0
for envs in "" "$@"; do
  echo "== bench [$envs]"
  env $envs timeout 600 python bench.py --steps 500 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_$i.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'out-layer ms',round(d['roofline']['avg_launch_ms'],4),'frac',round(d['roofline']['frac'],4))"
  i=$((i+1))
done
