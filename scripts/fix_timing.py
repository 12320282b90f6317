"""the output layer's training call in a loop (warm caches), for `ncu --cache-control none --metrics gpu__time_duration.sum` launch lists of
out_tc2_kernel / out_fix_kernel under NTF_FIX_EXP variants.   usage: python scripts/fix_timing.py [zipf]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, torch
from opentf_b200 import ops
from test_gpu_tc import run_tc, make_case
ws = ops.Workspace(torch.device('cuda:0'))
A, W, b, Y, negs = make_case(1000, 40000, 1)
if len(sys.argv) > 1:  # hot experts: every team has expert 7 as a member and expert 9 as a negative
    negs[:, 0] = 9
for rep in range(6): run_tc(ops, ws, A, W, b, Y, negs, 10.0, 1.0, train=True)
torch.cuda.synchronize()
