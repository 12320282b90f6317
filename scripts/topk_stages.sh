#!/bin/bash
# device time of the fused top-K call cut after stage 1 (pass 1), 2 (+ select), 3 (+ pass 2), and whole: the differences are the in-pipeline
# (warm, back-to-back) costs of the stages, which the serialised cold-cache ncu launch list overstates.   usage: scripts/topk_stages.sh [K] [batch]
for s in 1 2 3 0; do echo -n "stop=$s: "; NTF_IT_STOP=$s python scripts/topk_prof.py ${1:-10} ${2:-1000} 2>&1 | grep "fused=1"; done
echo -n "no dependent launch, whole: "; NTF_IT_PDL=0 python scripts/topk_prof.py ${1:-10} ${2:-1000} 2>&1 | grep "fused=1"
