"""Data-parallel step: the overlapped gradient exchange -- output layer's arena segment summed and stepped on a side
stream while the hidden layers' backward runs -- must equal the plain sequence (whole step, one all-reduce, one Adam) bit for bit
given the same sums.  One GPU: the exchange is replaced by an in-process stand-in (every rank contributed the same gradient);
the real 2-rank NCCL run is scripts/multi_gpu_check.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('precision,graphs', [('fp32', False), ('fp32', True), ('tf32', True)])
def test_overlapped_exchange_equals_one_allreduce_after_the_step(precision, graphs):
    from opentf_b200 import synth, _lib
    from test_gpu_graph import _engine
    tv = synth.make_teamsvecs('toy', seed=2)
    B = 128
    if precision == 'tf32' and not _lib.lib().ntf_tc_supported(B, 128, tv['member'].shape[1], 0): pytest.skip('shape not on the tensor-core path')
    res = []
    for overlap in (False, True):
        eng = _engine(tv, precision, graphs, B, nsd='unigram_b', h=128)
        eng.world, eng.rank, eng.dp_overlap = 2, 0, overlap
        calls = []
        eng.allreduce = lambda t, _c=calls: (_c.append(t.numel()), t.mul_(2.0))[1]
        sp = eng.split(np.arange(0, 4 * B))
        for e in range(3):
            for bi in range(4): eng.step(sp, bi * B, B, True, lr=1e-2, loss_slot=bi, loss_scale=0.5 / B)
        torch.cuda.synchronize()
        res.append((eng.loss_buf[:4].cpu().clone(), eng.params.cpu().clone(), eng.adam_m.cpu().clone(), calls, eng.adam_t))
    (l0, p0, m0, c0, t0), (l1, p1, m1, c1, t1) = res
    assert t0 == t1 == 12 and len(c0) == 12 and len(c1) == 24 and sum(c0) == sum(c1)  # same floats exchanged, in two segments
    if precision == 'fp32':
        assert torch.equal(l0, l1) and torch.equal(p0, p1) and torch.equal(m0, m1)
    else:  # dA is summed by L2 in arrival order on the tensor-core path
        assert torch.allclose(l0, l1, rtol=1e-4) and (p0 - p1).norm() <= 1e-3 * p0.norm()


@pytest.mark.parametrize('graphs', [False, True])
def test_exchange_inside_the_step_over_a_raw_nccl_communicator(graphs):
    """ntf_fnn_step_args.comm: the library all-reduces the two arena segments itself (ncclAllReduce on its communication stream, inside
    the captured step).  One GPU: a communicator of ONE rank (the sum is the identity), so the run must equal the plain single-GPU
    run bit for bit -- this checks the ctypes plumbing, the stream choreography and that the NCCL calls capture into the step's graph."""
    from opentf_b200 import synth, nccl
    from test_gpu_graph import _engine
    tv = synth.make_teamsvecs('toy', seed=2)
    B = 128
    comm = nccl.Comm(0, 1, 'cuda:0')
    t = torch.arange(8, dtype=torch.float32, device='cuda')
    assert torch.equal(comm.allreduce(t.clone()), t)
    res = []
    for use_comm in (False, True):
        eng = _engine(tv, 'fp32', graphs, B, nsd='unigram_b', h=128)
        if use_comm:
            eng.world, eng.rank = 2, 0  # (takes the data-parallel branch; the communicator itself has one rank)
            eng.attach_comm(comm)
        sp = eng.split(np.arange(0, 4 * B))
        for e in range(3):
            for bi in range(4): eng.step(sp, bi * B, B, True, lr=1e-2, loss_slot=bi)
            eng.step(sp, 0, B, False, loss_slot=4)  # a validation step in between (no exchange)
        torch.cuda.synchronize()
        res.append((eng.loss_buf[:5].cpu().clone(), eng.params.cpu().clone(), eng.adam_t, len(eng._graphs)))
    (l0, p0, t0, g0), (l1, p1, t1, g1) = res
    assert t0 == t1 == 12 and g0 == g1
    assert torch.equal(l0, l1) and torch.equal(p0, p1)
    comm.destroy()


# ---------------------------------------------------------------------------------------------- exchange + Adam over peer memory (csrc/peer.cu)
@pytest.mark.parametrize('graphs', [False, True])
def test_peer_exchange_with_one_rank_is_plain_adam(graphs):
    """a table of ONE rank: the fused pass reads its own gradients and writes its own parameters -- bit-identical to the single-GPU step"""
    from opentf_b200 import synth
    from test_gpu_graph import _engine
    tv = synth.make_teamsvecs('toy', seed=2)
    B = 128
    res = []
    for peers in (False, True):
        eng = _engine(tv, 'fp32', graphs, B, nsd='unigram_b', h=128)
        if peers:
            eng.world, eng.rank = 2, 0  # (takes the data-parallel branch; the table itself has one rank)
            eng.attach_peers(local=[eng])
            assert eng.peers is not None and eng.peers.world == 1
        sp = eng.split(np.arange(0, 4 * B))
        for e in range(3):
            for bi in range(4): eng.step(sp, bi * B, B, True, lr=1e-2, loss_slot=bi)
            eng.step(sp, 0, B, False, loss_slot=4)
        torch.cuda.synchronize()
        assert eng.peer_error() == 0
        res.append((eng.loss_buf[:5].cpu().clone(), eng.params.cpu().clone(), eng.adam_m.cpu().clone(), eng.adam_t))
    (l0, p0, m0, t0), (l1, p1, m1, t1) = res
    assert t0 == t1 == 12
    assert torch.equal(l0, l1) and torch.equal(p0, p1) and torch.equal(m0, m1)


def test_public_classes_on_two_gpus_equal_one_gpu():
    """the real thing (needs >= 2 GPUs, skipped on a single-GPU box): torchrun scripts/multi_gpu_check.py -- Fnn.learn/test through the
    public classes, data-parallel (peer-memory exchange, incl. a short last batch that leaves a rank idle) and expert-sharded, must
    reproduce the single-GPU run of the same seed (fp32 mode: weights to 1e-4 absolute; measured 2e-7).  Ranks are processes with a GPU
    each: several ranks on ONE device would have to wait for each other inside kernels that the device may not co-schedule."""
    import os, subprocess, sys
    if torch.cuda.device_count() < 2: pytest.skip('needs 2 GPUs (profiles/*multi_gpu_check* holds the output of the last 2-GPU run)')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', '29533', os.path.join(root, 'scripts', 'multi_gpu_check.py')], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'MULTI-GPU CHECK OK' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
