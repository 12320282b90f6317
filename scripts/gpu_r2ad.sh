#!/bin/bash
for rep in 1 2; do
echo "-- splits on";  python scripts/e2e_host_profile.py 2>&1 | grep "us per step\|per iteration"
echo "-- L1 flat";    NTF_ADAM_L1_FLAT=1 python scripts/e2e_host_profile.py 2>&1 | grep "us per step\|per iteration"
echo "-- all flat";   NTF_ADAM_ROWS_OFF=1 python scripts/e2e_host_profile.py 2>&1 | grep "us per step\|per iteration"
done
