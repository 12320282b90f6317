"""N-GPU check of the multi-GPU modes through the public class (torchrun, one process per GPU): Fnn.learn/test with parallel='dp' (every
gradient-exchange mode) and parallel='shard' must write the same checkpoint / predictions as the single-GPU run of the same seed
(fp32 mode; weights to 1e-4 absolute, losses to 1e-4 relative).  b=7 leaves a last batch of ONE team: all ranks but the first idle in it.
usage: torchrun --nproc-per-node 2 scripts/multi_gpu_check.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, torch, torch.distributed as dist
from conftest import load_toy
from opentf_b200.fnn import Fnn

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
skill, member, splits, _ = load_toy('gith')
tv = {'skill': skill.tolil(), 'member': member.tolil()}
one = {'test': splits['test'], 'folds': {0: splits['folds'][0]}}
root = '/tmp/ntf_multi_gpu_check'
os.makedirs(root, exist_ok=True)


def run(tag, b, **over):
    cfg = dict(b=b, e=4, ns=5, lr=0.01, es=10, h=[32], spe=0, l='bce', tpw=10, tnw=1, nsd='unigram_b', precision='fp32')
    cfg.update(over)
    if cfg.get('parallel') == 'none' and rank != 0: return None
    m = Fnn(f'{root}/{tag}', f'cuda:{local}', 0, cfg)
    m.learn(tv, one, None)
    m.test(tv, one, dict(on_train=False, per_epoch=False, topK=5))
    if cfg.get('parallel') != 'none': dist.barrier()
    if rank != 0: return None
    ck = torch.load(f'{m.output}/f0.pt', weights_only=False)['model_state_dict']
    pr = torch.load(f'{m.output}/f0.test.pred', weights_only=False)['y_pred']
    return ck, pr, m.last_history[0], (m.engine.peer_error() if m.engine.peers is not None else 0)


def same(name, a, ref):
    for k in ref[0]:
        d = (a[0][k] - ref[0][k]).abs().max().item()
        print(f'{name}: {k} {tuple(ref[0][k].shape)} max |d| vs single GPU = {d:.3g}')
        assert d < 1e-4, (name, k, d)
    print(f'{name}: losses {a[2][-1]} single {ref[2][-1]}  peer_error {a[3]}')
    assert abs(a[2][-1][0] - ref[2][-1][0]) < 1e-4 * abs(ref[2][-1][0]) and a[3] == 0
    assert a[1].is_sparse and torch.equal(a[1].indices(), ref[1].indices()), name  # same top-5 experts per test team


n_train = len(one['folds'][0]['train'])
for b in (8, 7):
    if rank == 0: print(f'--- b={b}: {n_train} train rows, last batch of {n_train % b or b}')
    ref = run(f'none_b{b}', b, parallel='none')
    dist.barrier()
    for tag, over in (('dp_peer', dict(parallel='dp', exchange='peer')), ('dp_nccl', dict(parallel='dp', exchange='nccl')), ('shard_torch', dict(parallel='shard', exchange='torch')),
                      ('dp_torch', dict(parallel='dp', exchange='torch')), ('shard', dict(parallel='shard'))):
        got = run(f'{tag}_b{b}', b, **over)
        if rank == 0: same(f'{tag} b={b} N={world}', got, ref)
        dist.barrier()
if rank == 0: print('MULTI-GPU CHECK OK')
dist.destroy_process_group()
