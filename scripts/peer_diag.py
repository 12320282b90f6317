"""diagnostics for tests/test_gpu_dp.py's peer-memory tests: where do the one-rank parameters differ; which barrier do two ranks on one device miss"""
import os, sys
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, torch
from opentf_b200 import synth
from test_gpu_graph import _engine

tv = synth.make_teamsvecs('toy', seed=2)
print('toy', tv['skill'].shape, tv['member'].shape)
B = 128
res = []
for peers in (False, True):
    eng = _engine(tv, 'fp32', False, B, nsd='unigram_b', h=128)
    if peers:
        eng.world, eng.rank = 2, 0
        eng.attach_peers(local=[eng])
    sp = eng.split(np.arange(0, 4 * B))
    eng.step(sp, 0, B, True, lr=1e-2, loss_slot=0)
    torch.cuda.synchronize()
    res.append((eng.params.cpu().clone(), eng.adam_m.cpu().clone(), eng.adam_v.cpu().clone(), eng.grads.cpu().clone(), dict(eng.views)))
(p0, m0, v0, g0, views), (p1, m1, v1, g1, _) = res
for name, a, b in (('grads', g0, g1), ('m', m0, m1), ('v', v0, v1), ('params', p0, p1)):
    d = (a != b).nonzero().flatten()
    print(name, 'differing elements:', d.numel(), 'first', d[:5].tolist(), 'max abs', (a - b).abs().max().item())
    if d.numel():
        i = int(d[0]); print('   ', a[i].item(), b[i].item(), [k for k, (o, s) in views.items() if o <= i < o + int(np.prod(s))])

G = 2
Bg, n_rows = 32 * G, 3 * 32 * G + 1
rows = np.arange(n_rows)
ranks = [_engine(tv, 'fp32', False, Bg, nsd='unigram_b', h=128) for _ in range(G)]
streams = [torch.cuda.Stream() for _ in range(G)]
sps = [e.split(rows) for e in ranks]
for e, sp in zip(ranks, sps): e.step(sp, 0, 32, True, lr=1e-2, loss_slot=0)
torch.cuda.synchronize()
for r, e in enumerate(ranks): e.world, e.rank = G, r
for e in ranks: e.attach_peers(local=ranks)
torch.cuda.synchronize()
import time
for b0 in range(0, n_rows, Bg):
    Bb = min(Bg, n_rows - b0)
    p_ = -(-Bb // G)
    t0 = time.time()
    for r, e in enumerate(ranks):
        lo, hi = min(Bb, r * p_), min(Bb, (r + 1) * p_)
        with torch.cuda.stream(streams[r]):
            if hi > lo: e.step(sps[r], b0 + lo, hi - lo, True, lr=1e-2, loss_slot=0, loss_scale=1.0 / Bb, gbatch=(b0, Bb))
            else: e.idle_step(True, 1e-2)
        print('  enqueued rank', r, 'batch', b0, 'host s', round(time.time() - t0, 3), flush=True)
    torch.cuda.synchronize()
    print('batch', b0, 'B', Bb, 'wall', round(time.time() - t0, 2), 's', flush=True)
    for r, e in enumerate(ranks): print('  rank', r, 'flags', e._peer_flags[:40].cpu().tolist(), flush=True)
    if any(e.peer_error() for e in ranks): break
