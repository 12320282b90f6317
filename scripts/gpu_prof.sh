#!/bin/bash
# One gpurun call: tile-level stamps of the tcgen05 kernel, bench variants with parts of the kernel switched off
# (NTF_TC_EXP), and ncu --set full captures of the top kernels.   usage: bash scripts/gpu_prof.sh [tag]
TAG=${1:-r01b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== stamps"; timeout 300 python scripts/tc_timing.py > $OUT/tc_timing.txt 2>&1; tail -n 40 $OUT/tc_timing.txt
for e in 0 1 2 3; do
  echo "== bench NTF_TC_EXP=$e"; NTF_TC_EXP=$e timeout 600 python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>&1 | tail -n 1 | tee $OUT/bench_exp$e.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['clocks'])"
done
echo "== ncu full: out_tc_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:out_tc_kernel -s 8 -c 2 -o $OUT/prof_out_tc -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_out_tc.log 2>&1
echo "== ncu full: csr_bag_bwd + topk + adam"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"csr_bag_bwd_cold_kernel|topk_select_kernel|adam_kernel|csr_bag_bwd_hot_kernel" -s 12 -c 5 -o $OUT/prof_others -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_others.log 2>&1
ls -la $OUT
