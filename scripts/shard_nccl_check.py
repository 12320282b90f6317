"""2-GPU check of the expert-sharded mode through the public class (torchrun, NCCL): Fnn.learn/test with parallel='shard' must write the
same checkpoint / predictions as the single-GPU run of the same seed (fp32 mode).  usage: torchrun --nproc-per-node 2 scripts/shard_nccl_check.py"""
import os, sys, tempfile
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
root = '/tmp/ntf_shard_check'
os.makedirs(root, exist_ok=True)
res = {}
for mode in ('shard', 'dp'):
    cfg = dict(b=8, e=4, ns=5, lr=0.01, es=10, h=[32], spe=0, l='bce', tpw=10, tnw=1, nsd='unigram_b', precision='fp32', parallel=mode)
    m = Fnn(f'{root}/{mode}', f'cuda:{local}', 0, cfg)
    m.learn(tv, one, None)
    m.test(tv, one, dict(on_train=False, per_epoch=False, topK=5))
    dist.barrier()
    if rank == 0:
        ck = torch.load(f'{m.output}/f0.pt', weights_only=False)
        pr = torch.load(f'{m.output}/f0.test.pred', weights_only=False)['y_pred']
        res[mode] = (ck, pr, m.last_history[0])
if rank == 0:
    # single-process reference of the same seed: the whole batch on one GPU
    cfg = dict(b=8, e=4, ns=5, lr=0.01, es=10, h=[32], spe=0, l='bce', tpw=10, tnw=1, nsd='unigram_b', precision='fp32', parallel='none')
    m = Fnn(f'{root}/none', f'cuda:{local}', 0, cfg)
    m.learn(tv, one, None)
    ck1 = torch.load(f'{m.output}/f0.pt', weights_only=False)['model_state_dict']
    for k, v in ck1.items():
        d = (v - res['dp'][0]['model_state_dict'][k]).abs().max().item()
        print(k, 'max |single - dp| =', d)
        assert d < 1e-4, k
    print('losses single', m.last_history[0][-1], 'dp', res['dp'][2][-1])
    assert abs(m.last_history[0][-1][0] - res['dp'][2][-1][0]) < 1e-4 * abs(res['dp'][2][-1][0])
dist.barrier()
if rank == 0:
    a, b = res['shard'], res['dp']
    for k in a[0]['model_state_dict']:
        d = (a[0]['model_state_dict'][k] - b[0]['model_state_dict'][k]).abs().max().item()
        print(k, tuple(a[0]['model_state_dict'][k].shape), 'max |shard - dp| =', d)
        assert d < 1e-4, k
    print('losses shard', a[2][-1], 'dp', b[2][-1])
    assert abs(a[2][-1][0] - b[2][-1][0]) < 1e-3 * abs(b[2][-1][0])
    assert a[1].is_sparse and tuple(a[1].shape) == tuple(b[1].shape)
    print('SHARD CHECK OK')
dist.destroy_process_group()
