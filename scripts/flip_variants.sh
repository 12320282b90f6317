#!/bin/bash
# The intermittent Flipout failure of round 1, reproduced and bisected: builds libntf_b200 variants that differ ONLY in out_tc.cu --
#   fixed  today's file;   early  today's file with -DNTF_FLIP_EARLY_RELEASE (the Q slot released right after its loads were issued: round 1's
#   hand-over);   r1  the file as round 1 left it (commit 2d04170) -- and runs scripts/flip_stress2.py on each (fresh engine per iteration).
# Round-2 record: profiles/r02a_flipout_bisect_variants.txt (first bisect, -DNTF_VARIANT bits) and profiles/r02b_flipout_fixed.txt.
#   usage: bash scripts/flip_variants.sh build   (CPU box)   |   run [iters]  (GPU box)
cd "$(dirname "$0")/.."
CS=opentf_b200/csrc
OUT=$CS/build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=default --expt-relaxed-constexpr -I include -I $CS"
if [ "$1" = build ]; then
  rm -rf $OUT; mkdir -p $OUT
  python -m opentf_b200.csrc.build > /dev/null
  OBJS=$(ls $CS/build/*.o | grep -v out_tc.o)
  nvcc $FLAGS -c $CS/out_tc.cu -o $OUT/out_tc_fixed.o && nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart static -o $OUT/libntf_fixed.so $OBJS $OUT/out_tc_fixed.o && echo built fixed
  nvcc $FLAGS -DNTF_FLIP_EARLY_RELEASE -c $CS/out_tc.cu -o $OUT/out_tc_early.o && nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart static -o $OUT/libntf_early.so $OBJS $OUT/out_tc_early.o && echo built early
  git show 2d04170:opentf_b200/csrc/out_tc.cu > $OUT/out_tc_r1.cu
  nvcc $FLAGS -c $OUT/out_tc_r1.cu -o $OUT/out_tc_r1.o && nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart static -o $OUT/libntf_r1.so $OBJS $OUT/out_tc_r1.o && echo built r1
else
  IT=${2:-40}
  mkdir -p gpurun_out/variants
  for v in r1 early fixed; do
    n=$IT; [ $v = fixed ] && n=$((IT * 6))
    echo "== variant $v ($n iterations)"; NTF_B200_LIB=$PWD/$OUT/libntf_$v.so timeout 900 python scripts/flip_stress2.py $n 2>&1 | grep -v "differs from iteration 0 in: nothing" | tail -4 | cut -c1-300 | tee gpurun_out/variants/stress2_$v.txt
  done
fi
