"""diagnostic: clock64 stamps of the persistent output-layer kernel (out_tc2.cu): per CTA start / end, per tile when its logits were ready and
when its epilogue finished, per expert tile the dW drain.   usage: python scripts/tc2_timing.py [B] [E]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, torch
from opentf_b200 import ops
from test_gpu_tc import run_tc, make_case
ws = ops.Workspace(torch.device('cuda:0'))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
E = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
A, W, b, Y, negs = make_case(B, E, 1)
G = 1024
for mode in ('train', 'valid'):
    tim = torch.zeros(512 * G, dtype=torch.int64, device='cuda:0')
    os.environ['NTF_TC_TIMING'] = str(tim.data_ptr())
    for rep in range(2):
        tim.zero_()
        run_tc(ops, ws, A, W, b, Y, negs, 10.0, 1.0, train=(mode == 'train'))
        torch.cuda.synchronize()
    os.environ.pop('NTF_TC_TIMING')
    raw = tim.cpu().numpy().reshape(G, 512)
    c = raw[raw[:, 0] > 0]
    g0 = c[:, 0].min()
    dur = (c[:, 1] - c[:, 0]) / 1e3
    cyc = c[:, 4] - c[:, 3]
    print(f'== {mode}: {len(c)} CTAs; kernel span {(c[:, 1].max() - g0) / 1e3:.1f} us; CTA duration us: min {dur.min():.1f} median {np.median(dur):.1f} max {dur.max():.1f}; '
          f'cycles median {int(np.median(cyc))} -> {np.median(cyc) / np.median(dur) / 1e3:.2f} GHz; SMs used {len(np.unique(c[:, 2]))}')
    for cta in (0, len(c) // 2, len(c) - 1):
        r = c[cta]
        t0 = r[3]
        st = r[8:8 + 480].reshape(60, 8)
        n = int((st[:, 0] > 0).sum())
        print(f'  CTA {cta} (SM {r[2]}): {n} tiles, total {r[4] - t0} cycles')
        print('    tile: Z ready / tile done (+ drain start / end at the end of an expert tile), cycles since CTA start')
        prev = 0
        for i in range(n):
            z, d, ds, de, b0, b1, da, dwf = (int(v - t0) if v > 0 else -1 for v in st[i])
            extra = f'   drain {ds} -> dW complete {dwf} -> {de} ({de - ds})' if ds > 0 else ''
            print(f'    {i:3d}: {z:7d} {d:7d}  (epilogue {d - z:5d}, since previous done {z - prev:5d}) bwd issue {b0:7d}..{b1:7d} dA staged {da:7d}{extra}')
            prev = de if de > 0 else d
