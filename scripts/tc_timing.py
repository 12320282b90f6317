"""diagnostic: clock64 stamps of CTA 0 of the tcgen05 kernels (where does a tile's time go?)"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, torch
from opentf_b200 import ops
from test_gpu_tc import run_tc, make_case
ws = ops.Workspace(torch.device('cuda:0'))
B, E = 1000, 40000
A, W, b, Y, negs = make_case(B, E, 1)
names = ['mma:bwd_start', 'mma:A_full', 'mma:Z_empty', 'mma:fwd_done', 'epi:Z_full', 'epi:ld_done', 'epi:tile_done', 'mma:bwd_done']
for mode in ('infer', 'train'):
    tim = torch.zeros(64 * 8, dtype=torch.int64, device='cuda:0')
    os.environ['NTF_TC_TIMING'] = str(tim.data_ptr())
    for rep in range(2):
        tim.zero_()
        if mode == 'infer':
            P = torch.empty(B, E, device='cuda:0')
            ops.infer_scores(1, A.cuda(), W.cuda(), b.cuda(), B, 128, E, P, ws)
        else:
            run_tc(ops, ws, A, W, b, Y, negs, 10.0, 1.0, train=(mode == 'train'))
        torch.cuda.synchronize()
    os.environ.pop('NTF_TC_TIMING')
    t = tim.cpu().numpy().reshape(64, 8)[:16]
    t0 = t[t > 0].min()
    print('==', mode, '(cycles since first stamp)')
    print('tile ' + ' '.join(f'{n:>15s}' for n in names))
    for i in range(16): print(f'{i:4d} ' + ' '.join(f'{(v - t0) if v > 0 else -1:15d}' for v in t[i]))
