#!/bin/bash
OUT=gpurun_out/r2ag; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/topk1000_launches.csv python scripts/topk_prof.py 1000 > $OUT/under_ncu.log 2>&1
python scripts/summarize_launches.py $OUT/topk1000_launches.csv 2>&1 | tail -16 | head -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"infer_topk_final_kernel" -s 3 -c 1 -o $OUT/prof_final -f python scripts/topk_prof.py 1000 > $OUT/ncu_final.log 2>&1; ls -la $OUT | tail -3
