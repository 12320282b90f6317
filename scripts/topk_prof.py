"""top-K inference on the C2 shape: the fused path (ntf_infer_topk) and the unfused one, a few calls each -- run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel launch list, or alone for event timings.   usage: python scripts/topk_prof.py [K] [batch]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from bench import workload
from opentf_b200.engine import Engine

K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
tv, splits = workload('dblp')
N, S = tv['skill'].shape; E = tv['member'].shape[1]
dev = torch.device('cuda:0')
eng = Engine(S, [128], E, dev, precision='tf32', nsd='unigram_b', ns=5, max_batch=b)
eng.stage(tv['skill'], tv['member'])
torch.manual_seed(0)
lin = [torch.nn.Linear(S, 128), torch.nn.Linear(128, E)]
for m in lin: torch.nn.init.xavier_uniform_(m.weight)
eng.load_state_dict({f'layers.{i}.{n}': getattr(m, n).detach() for i, m in enumerate(lin) for n in ('weight', 'bias')})
sp = eng.split(np.asarray(splits['test']))
ib = min(b, sp.n)
scores = torch.empty(ib, E, device=dev)
vals, idx = torch.empty(ib, K, device=dev), torch.empty(ib, K, dtype=torch.int32, device=dev)
for mode in ('1', '0'):
    os.environ['NTF_FUSED_TOPK'] = mode
    for _ in range(3): eng.topk(sp, 0, ib, K, scores, vals, idx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    h0 = time.perf_counter()
    e0.record()
    for r in range(reps): eng.topk(sp, (r * ib) % max(1, sp.n - ib + 1), ib, K, scores, vals, idx)
    e1.record()
    host = (time.perf_counter() - h0) / reps * 1e6
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f'fused={mode} K={K} batch={ib}: {us:.1f} us per call on the device ({ib / us:.2f} M teams/s), host enqueue {host:.1f} us per call', flush=True)
