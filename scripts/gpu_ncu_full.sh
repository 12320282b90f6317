#!/bin/bash
# ncu --set full captures of the top kernels of one bench step (one GPU; numbers printed under ncu are never bench values)
TAG=${1:-ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"out_tc_kernel" -s 10 -c 2 -o $OUT/prof_out_tc -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_out_tc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"adam_kernel|bag_bwd_reduce_kernel|csr_bag_bwd_hot_kernel|topk_small_kernel|csr_bag_fwd_kernel|neg_sample_kernel" -s 24 -c 8 -o $OUT/prof_others -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_others.log 2>&1
ls -la $OUT
