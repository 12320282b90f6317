"""diagnostic: clock64 stamps of the fused top-K kernel (infer_topk.cu), one pass at a time (NTF_IT_TIMING_PASS): per product when
its W tile had landed, when it was issued, when its logits were ready and when the epilogue was done.   usage: python scripts/topk_timing.py [B] [E] [K]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from opentf_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
E = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = 'cuda:0'
torch.manual_seed(0)
A, W, b = torch.randn(B, 128, device=dev).abs(), torch.randn(E, 128, device=dev) * 0.2, torch.randn(E, device=dev) * 0.5
W16 = torch.empty(E * 128, dtype=torch.float16, device=dev); ops.to_half(W, E * 128, W16)
ws = ops.Workspace(torch.device(dev))
vals, idx = torch.empty(B, K, device=dev), torch.empty(B, K, dtype=torch.int32, device=dev)
G = 1024
for tpass in (1, 2):
    tim = torch.zeros(256 * G, dtype=torch.int64, device=dev)
    os.environ['NTF_IT_TIMING_PASS'] = str(tpass)
    for rep in range(3):
        if rep == 2: os.environ['NTF_IT_TIMING'] = str(tim.data_ptr())
        ops.infer_topk(A, W16, b, B, 128, E, K, vals, idx, ws)
        torch.cuda.synchronize()
    os.environ.pop('NTF_IT_TIMING')
    raw = tim.cpu().numpy().reshape(G, 256)
    c = raw[raw[:, 0] > 0]
    cyc = c[:, 1] - c[:, 0]
    print(f'pass {tpass}: {len(c)} CTAs; cycles per CTA: min {cyc.min()} median {int(np.median(cyc))} max {cyc.max()}')
    g0, g1 = c[:, 2], c[:, 3]
    print(f'  globaltimer: first CTA start -> last CTA end {(g1.max() - g0.min()) / 1e3:.2f} us; CTA starts spread over {(g0.max() - g0.min()) / 1e3:.2f} us; '
          f'CTA ends spread over {(g1.max() - g1.min()) / 1e3:.2f} us; CTA lifetime median {np.median(g1 - g0) / 1e3:.2f} us max {(g1 - g0).max() / 1e3:.2f} us')
    for cta in (0, len(c) // 2):
        r = c[cta]; t0 = r[0]
        st = r[8:8 + 240].reshape(60, 4)
        n = int((st[:, 2] > 0).sum())
        print(f'  CTA {cta}: {n} products, total {r[1] - t0} cycles;  per product: W landed / issued / logits ready / epilogue done')
        for i in range(n):
            print('   %3d: %7d %7d %7d %7d' % ((i,) + tuple(int(v - t0) if v > 0 else -1 for v in st[i])))
