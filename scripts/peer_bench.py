"""micro-benchmark of the data-parallel exchange at the C2 arena size (8.96 M parameters), one process per GPU (torchrun):
the fused peer-memory pass (whole arena / the step's two segments / barriers only) against ncclAllReduce + Adam.  Times are CUDA events
on the launching stream, max over ranks, after warm-up."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, torch.distributed as dist
from opentf_b200 import ops
from opentf_b200.engine import Engine

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device(f'cuda:{local}')
dist.init_process_group('nccl', device_id=dev)
eng = Engine(30000, [128], 40000, dev, precision='fp32', nsd='uniform', max_batch=8)
eng.world, eng.rank = world, rank
eng.attach_peers()
eng.grads.normal_(); eng.params.normal_()
n = eng.n_params
split = eng.views['layers.1.weight'][0]


def timed(fn, iters=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t.item()), 1)


step = [0]
def ex(off, cnt, ch):
    step[0] += 1
    ops.peer_exchange_adam(eng.dev_index, eng.peers, eng.adam_m, eng.adam_v, off, cnt, 1e-3, 0.9, 0.999, 1e-8, step[0], ch)

res = {'world': world, 'n_params': n, 'arena_MB': round(n * 4 / 1e6, 1)}
res['peer_whole_arena_us'] = timed(lambda: ex(0, n, 0))
res['peer_two_segments_us'] = timed(lambda: (ex(split, n - split, 1), ex(0, split, 0)))
res['peer_out_layer_segment_us'] = timed(lambda: ex(split, n - split, 1))
res['peer_rest_segment_us'] = timed(lambda: ex(0, split, 0))
res['peer_barriers_only_us'] = timed(lambda: ex(0, 0, 0) if False else ops.peer_exchange_adam(eng.dev_index, eng.peers, eng.adam_m, eng.adam_v, 0, 4 * world, 1e-3, 0.9, 0.999, 1e-8, 1, 0))
def nccl_adam():
    dist.all_reduce(eng.grads)
    ops.adam_step(eng.params, eng.grads, eng.adam_m, eng.adam_v, n, 1e-3, 0.9, 0.999, 1e-8, 1)
    eng.grads.mul_(1.0 / world)  # (keep the values finite over the iterations)
res['nccl_allreduce_plus_adam_us'] = timed(nccl_adam)
res['adam_alone_us'] = timed(lambda: ops.adam_step(eng.params, eng.grads, eng.adam_m, eng.adam_v, n, 1e-3, 0.9, 0.999, 1e-8, 1))
res['peer_error'] = eng.peer_error()
per_rank_bytes = 4 * n * (world - 1) / world
res['link_GBps_each_way_whole_arena'] = round(per_rank_bytes / (res['peer_whole_arena_us'] * 1e-6) / 1e9, 1)
if rank == 0: print(json.dumps(res))
dist.destroy_process_group()
