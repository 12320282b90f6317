"""Expert-sharded output layer at the USPT shape (BASELINE configs[3]: S=213317, E=394187) on N GPUs against the unsharded layer on one:
the same seeded batches of b=1000 through Engine(shard=(rank, N)) -- dA exchanged inside the step over peer memory -- and, on rank 0, through
an unsharded engine.  Losses (summed over the shards), the replicated first layer and this rank's rows of the output layer must agree
(tensor-core mode: 2e-3 on the losses, 1e-2 relative in norm on the weights after a few Adam steps).   usage: torchrun --nproc-per-node N scripts/shard_c4_check.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch, torch.distributed as dist
from bench import workload
from opentf_b200.engine import Engine

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device(f'cuda:{local}')
if world > 1: dist.init_process_group('nccl', device_id=dev)
tv, splits = workload('uspt')
N, S = tv['skill'].shape; E = tv['member'].shape[1]
b, steps = 1000, 4
torch.manual_seed(0)
lin = [torch.nn.Linear(S, 128), torch.nn.Linear(128, E)]
for m in lin: torch.nn.init.xavier_uniform_(m.weight)
sd = {f'layers.{i}.{n}': getattr(m, n).detach() for i, m in enumerate(lin) for n in ('weight', 'bias')}
rows = np.asarray(splits['folds'][0]['train'])[:steps * b]


def run(shard):
    eng = Engine(S, [128], E, dev, precision='tf32', tpw=10, tnw=1, nsd='unigram_b', ns=5, seed=0, max_batch=b, shard=shard)
    if shard is not None:
        eng.world, eng.rank = world, rank
        if world > 1: eng.attach_shard_peers()
    eng.stage(tv['skill'], tv['member'])
    eng.load_state_dict(sd)
    sp = eng.split(rows)
    for i in range(steps): eng.step(sp, i * b, b, True, lr=1e-3, loss_slot=i, loss_scale=1.0 / b, gbatch=(i * b, b))
    torch.cuda.synchronize()
    assert eng.peer_error() == 0
    return eng


sh = run((rank, world))
loss = sh.loss_buf[:steps].clone()
if world > 1: dist.all_reduce(loss)
ok = True
if rank == 0:
    full = run(None)
    lf = full.loss_buf[:steps]
    print(f'N={world}: E={E} S={S} b={b}; losses sharded {loss.tolist()} unsharded {lf.tolist()}')
    ok &= bool(((loss - lf).abs() <= 2e-3 * lf.abs()).all())
    w1s, w1f = sh.view('layers.0.weight'), full.view('layers.0.weight')
    w2s, w2f = sh.view('layers.1.weight'), full.view('layers.1.weight')[sh.e_lo:sh.e_hi]
    d1 = ((w1s - w1f).norm() / (w1f - torch.as_tensor(sd['layers.0.weight']).t().to(dev)).norm()).item()  # relative to how far the weights moved
    d2 = ((w2s - w2f).norm() / (w2f - torch.as_tensor(sd['layers.1.weight'])[sh.e_lo:sh.e_hi].to(dev)).norm()).item()
    print(f'after {steps} Adam steps: layer 0 (replicated) differs by {d1:.3g} of its update, layer 1 rows [{sh.e_lo},{sh.e_hi}) by {d2:.3g}')
    ok &= d1 < 5e-2 and d2 < 5e-2
    print('SHARD C4 CHECK OK' if ok else 'SHARD C4 CHECK FAILED')
if world > 1: dist.destroy_process_group()
sys.exit(0 if ok else 1)
