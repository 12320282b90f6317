"""diagnostic: clock64 stamps of CTA 0 of the tcgen05 kernels (where does a tile's time go?) and the start/end of every CTA"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, torch
from opentf_b200 import ops
from test_gpu_tc import run_tc, make_case
ws = ops.Workspace(torch.device('cuda:0'))
B, E = 1000, 40000
A, W, b, Y, negs = make_case(B, E, 1)
names = ['mma:bwd_start', 'mma:A_full', 'mma:Z_empty', 'dA:drained', 'epi:Z_full', 'epi:ld_done', 'epi:tile_done', 'mma:dA_free']
nct = (E + 127) // 128 + 256  # (+ the CTAs of a split last wave)
for mode, exp in (('infer', '0'), ('train', '0'), ('train', '1'), ('train', '2'), ('train', '8')):
    tim = torch.zeros(128 + 3 * nct + 8, dtype=torch.int64, device='cuda:0')
    os.environ['NTF_TC_TIMING'] = str(tim.data_ptr()); os.environ['NTF_TC_EXP'] = exp
    for rep in range(2):
        tim.zero_()
        if mode == 'infer':
            P = torch.empty(B, E, device='cuda:0')
            ops.infer_scores(1, A.cuda(), W.cuda(), b.cuda(), B, 128, E, P, ws)
        else:
            run_tc(ops, ws, A, W, b, Y, negs, 10.0, 1.0, train=(mode == 'train'))
        torch.cuda.synchronize()
    os.environ.pop('NTF_TC_TIMING'); os.environ.pop('NTF_TC_EXP')
    raw = tim.cpu().numpy()
    t = raw[:128].reshape(16, 8)
    t0 = t[:15][t[:15] > 0].min()
    print('==', mode, 'exp', exp, '(cycles since first stamp; exp 4 = serialised pipeline, 8 = no split of the last wave)')
    print('CTA 0: entry, setup done, W image ready, products complete, drained:', [int(v - t[15][0]) for v in t[15][:5]])
    print('tile ' + ' '.join(f'{n:>15s}' for n in names))
    for i in range(8): print(f'{i:4d} ' + ' '.join(f'{(v - t0) if v > 0 else -1:15d}' for v in t[i]))
    c = raw[128:128 + 3 * nct].reshape(nct, 3)
    c = c[c[:, 0] > 0]
    g0 = c[:, 0].min()
    dur = (c[:, 1] - c[:, 0]) / 1e3
    print(f'CTAs: {len(c)}; kernel span {(c[:, 1].max() - g0) / 1e3:.1f} us; CTA duration us: min {dur.min():.1f} median {np.median(dur):.1f} max {dur.max():.1f}')
    order = np.argsort(c[:, 0])
    print('start us of CTAs (sorted, every 20th):', [round(float(c[i, 0] - g0) / 1e3, 1) for i in order[::20]])
    print('SMs used:', len(np.unique(c[:, 2])), ' CTAs on the busiest SM:', np.bincount(c[:, 2].astype(int)).max())
