"""where the host's time goes in the end-to-end loop of bench.py (HostBatches.run): per iteration the wait for the loader thread's block, the
pack of the next batch, Engine.step_host (H2D + step enqueue) and the wait for the previous loss, then a cProfile of 100 iterations.   usage: python scripts/e2e_host_profile.py"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from bench import workload, HostBatches
from opentf_b200.engine import Engine

tv, splits = workload('dblp')
N, S = tv['skill'].shape; E = tv['member'].shape[1]
dev = torch.device('cuda:0')
b = 1000
eng = Engine(S, [128], E, dev, precision='tf32', nsd='unigram_b', ns=5, max_batch=b)
eng.stage(tv['skill'], tv['member'])
torch.manual_seed(0)
lin = [torch.nn.Linear(S, 128), torch.nn.Linear(128, E)]
eng.load_state_dict({f'layers.{i}.{n}': getattr(m, n).detach() for i, m in enumerate(lin) for n in ('weight', 'bias')})
train_rows = np.asarray(splits['folds'][0]['train'])
host = HostBatches(tv, train_rows, b, 0, 1)
host.run(eng, 0, 6)
torch.cuda.synchronize()
steps = 200
T = np.zeros((steps, 3))
gB = b; lo, hi = host._slice()
t_all = time.perf_counter()
eng.step_host(host.packer.pack(host._rows(100), lo, hi), gB, host.cap_s, host.cap_m, 0, 1, lr=1e-3, sync=False, slot=0)
for i in range(1, steps):
    t0 = time.perf_counter(); blk = host.packer.pack(host._rows(100 + i), lo, hi)
    t1 = time.perf_counter()
    eng.step_host(blk, gB, host.cap_s, host.cap_m, 0, 1, lr=1e-3, sync=False, slot=i & 1)
    t2 = time.perf_counter()
    eng.step_host_loss((i - 1) & 1)
    t3 = time.perf_counter()
    T[i] = (t1 - t0, t2 - t1, t3 - t2)
eng.step_host_loss((steps - 1) & 1)
tot = (time.perf_counter() - t_all) / steps * 1e6
m = T[20:].mean(0) * 1e6
print(f'per iteration {tot:.1f} us: pack {m[0]:.1f}, step_host enqueue {m[1]:.1f}, wait for the previous loss {m[2]:.1f}')
def timed(label, fn, n=200):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(n); e1.record(); torch.cuda.synchronize()
    print(f'{label}: {e0.elapsed_time(e1) / n * 1e3:.1f} us per step on the device clock')

blk0 = host.packer.pack(host._rows(100), lo, hi)
def same_block(n):
    for i in range(n): eng.step_host(blk0, gB, host.cap_s, host.cap_m, 0, 1, lr=1e-3, sync=False, slot=i & 1)
timed('streaming, one block again and again (no pack)', same_block)
st = eng._hstream['stage'][0]
def resident_on_stage(n):
    for i in range(n): eng.step(st, 0, b, True, lr=1e-3, loss_slot=0, loss_scale=1.0 / b, gbatch=(0, b))
timed('resident steps on the staged block (no copies, no events)', resident_on_stage)
nbb = min(64, len(train_rows) // b); sp = eng.split(train_rows[:nbb * b])
def resident(n):
    for i in range(n): eng.step(sp, (i % nbb) * b, b, True, lr=1e-3, loss_slot=0)
resident(70)
timed('resident steps on the device split', resident)
