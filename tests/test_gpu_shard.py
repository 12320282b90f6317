"""Expert-sharded output layer (BASELINE config 4) on ONE GPU: two engines own the two halves of the expert axis,
the test plays the collectives (sum of dA, sum of the loss partials, gather of the top-k lists) in-process, and the result must
equal the unsharded engine's.  fp32 mode: equality up to summation order."""
import numpy as np
import pytest
import torch

from oracle import fnn_oracle as O
from test_gpu_kernels import DEV, dense, rand_csr, rel_err

pytestmark = pytest.mark.gpu


def build(S, hidden, E, B, skill, member, sd, shard, precision='fp32', nsd='unigram_b'):
    from opentf_b200.engine import Engine
    eng = Engine(S, hidden, E, DEV, precision=precision, tpw=10, tnw=1, nsd=nsd, ns=5, seed=11, max_batch=B, shard=shard)
    eng.stage(skill, member)
    eng.load_state_dict(sd)
    return eng


@pytest.mark.parametrize('precision,E,hidden', [('fp32', 301, [24]), ('fp32', 200, [16, 8]), ('tf32', 1000, [128])])
def test_two_shards_equal_the_unsharded_layer(precision, E, hidden):
    rng = np.random.default_rng(E)
    torch.manual_seed(E)
    B, S = 96, 40
    skill, member = rand_csr(rng, B, S, 1, 5), rand_csr(rng, B, E, 1, 4)
    layers = O.init_params(S, hidden, E)
    sd = {f'layers.{i}.{n}': t for i, (W, b) in enumerate(layers) for n, t in (('weight', W), ('bias', b))}
    full = build(S, hidden, E, B, skill, member, sd, None, precision)
    shards = [build(S, hidden, E, B, skill, member, sd, (i, 2), precision) for i in range(2)]
    assert shards[0].E + shards[1].E == E and shards[1].e_lo == shards[0].e_hi
    sp = full.split(np.arange(B))
    sps = [e.split(np.arange(B)) for e in shards]
    tol = 1e-5 if precision == 'fp32' else 3e-3
    for step in range(3):
        full.step(sp, 0, B, True, lr=1e-2, loss_slot=step)
        # phase 1 on both shards, then the all-reduce of dA played by hand, then phase 2 (Engine.step does exactly this over NCCL)
        pending = []
        for e, s_ in zip(shards, sps):
            e.allreduce = lambda t, _p=pending: _p.append(t)
        import threading
        # run the two shards' steps interleaved: phase 1 of both must finish before either adds the other's dA
        barrier = threading.Barrier(2)
        def run(e, s_):
            def ar(t):
                barrier.wait()               # both shards have produced their partial dA
                torch.cuda.synchronize()
                other = shards[1 - shards.index(e)].dact[-1][:B]
                e._sum = t + other           # read both partials ...
                barrier.wait()               # ... before either is overwritten
                t.copy_(e._sum)
            e.allreduce = ar
            e.step(s_, 0, B, True, lr=1e-2, loss_slot=step)
        th = [threading.Thread(target=run, args=(e, s_)) for e, s_ in zip(shards, sps)]
        [t.start() for t in th]; [t.join() for t in th]
        torch.cuda.synchronize()
        loss_full = full.loss_buf[step].item()
        loss_sh = sum(e.loss_buf[step].item() for e in shards)   # the epoch's loss all-reduce (fnn.Fnn.learn)
        assert abs(loss_sh - loss_full) <= tol * abs(loss_full), (step, loss_sh, loss_full)
    sdf = full.state_dict()
    last = len(hidden)
    for name, t in sdf.items():
        if name.startswith(f'layers.{last}.'):
            got = torch.cat([e.state_dict(gather=False)[name] for e in shards])
        else:
            got = shards[0].state_dict(gather=False)[name]
            assert torch.equal(got, shards[1].state_dict(gather=False)[name])  # replicated layers stay bit-identical across shards
        assert rel_err(got, t) <= (2e-5 if precision == 'fp32' else 2e-2), name
    # inference: local top-k lists with global ids, merged
    K = 7
    scores = torch.empty(B, E, device=DEV); vf = torch.empty(B, K, device=DEV); jf = torch.empty(B, K, dtype=torch.int32, device=DEV)
    full.topk(sp, 0, B, K, scores, vf, jf)
    lists = []
    for e, s_ in zip(shards, sps):
        sc = torch.empty(B, e.E, device=DEV)
        e.scores(s_, 0, B, sc)
        v = torch.empty(B, K, device=DEV); j = torch.empty(B, K, dtype=torch.int32, device=DEV)
        from opentf_b200 import ops
        ops.topk_select(sc, B, e.E, K, 1.0, v, j)
        lists.append((v, j + e.e_lo))
    from opentf_b200 import ops
    gv, gi = torch.stack([l[0] for l in lists]).contiguous(), torch.stack([l[1] for l in lists]).contiguous()
    vm, jm = torch.empty(B, K, device=DEV), torch.empty(B, K, dtype=torch.int32, device=DEV)
    ops.topk_merge(gv, gi, 2, B, K, vm, jm)
    if precision == 'fp32':
        assert (vm - vf).abs().max().item() <= 1e-6
        same = (jm == jf) | ((vm - vf).abs() <= 1e-6)  # ids may differ only inside a tie
        assert bool(same.all())
    else:
        assert (vm - vf).abs().max().item() <= 2e-3


def test_special_planes_of_a_shard_cover_its_expert_range_only():
    from opentf_b200 import ops
    from test_gpu_kernels import dev_csr
    rng = np.random.default_rng(3)
    B, E, ns = 40, 300, 5
    Y = rand_csr(rng, B, E, 1, 6)
    indptr, indices = dev_csr(Y)
    negs = rng.integers(-1, E, (B, ns)).astype(np.int32)
    e_lo, El = 100, 130
    pitch = (El + 31) // 32
    plane = torch.zeros(B, pitch, dtype=torch.int32, device=DEV)
    ops.special_bits(1, B, indptr.data_ptr(), indices, torch.from_numpy(negs).to(DEV), ns, El, plane, pitch, e_lo=e_lo)
    ref = np.asarray(Y.todense()).astype(bool)
    for n in range(B):
        for j in negs[n]:
            if j >= 0: ref[n, j] = True
    words = plane.cpu().numpy().view(np.uint32)
    got = ((words[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(B, -1)[:, :El].astype(bool)
    assert (got == ref[:, e_lo:e_lo + El]).all()


@pytest.mark.parametrize('precision,graphs', [('fp32', False), ('tf32', True)])
def test_sharded_step_in_one_call_equals_the_two_phase_step(precision, graphs, monkeypatch):
    """ntf_fnn_step with a peer table on an expert shard: the dA exchange (ntf_peer_allreduce) runs inside the call.  One GPU: a table of ONE
    rank (the sum over shards is this shard's own partial), against the two-phase step with an identity all-reduce -- same arithmetic, so the
    parameters must agree bit for bit in fp32 mode; this checks the plumbing (exchange block, flags, the phase-3 sequence, graph capture)."""
    monkeypatch.setenv('NTF_GRAPHS', '1' if graphs else '0')
    rng = np.random.default_rng(17)
    torch.manual_seed(17)
    B, S, E, hidden = 128, 40, 1000, [128]
    skill, member = rand_csr(rng, 4 * B, S, 1, 5), rand_csr(rng, 4 * B, E, 1, 4)
    layers = O.init_params(S, hidden, E)
    sd = {f'layers.{i}.{n}': t for i, (W, b) in enumerate(layers) for n, t in (('weight', W), ('bias', b))}
    res = []
    for one_call in (False, True):
        eng = build(S, hidden, E, B, skill, member, sd, (1, 2), precision)
        if one_call: eng.attach_shard_peers(local=[eng])
        else: eng.allreduce = lambda t: t
        sp = eng.split(np.arange(4 * B))
        for e in range(3):
            for bi in range(4): eng.step(sp, bi * B, B, True, lr=1e-2, loss_slot=bi)
        torch.cuda.synchronize()
        assert eng.peer_error() == 0
        res.append((eng.loss_buf[:4].cpu().clone(), eng.params.cpu().clone()))
    (l0, p0), (l1, p1) = res
    if precision == 'fp32': assert torch.equal(l0, l1) and torch.equal(p0, p1)
    else: assert torch.allclose(l0, l1, rtol=1e-4) and (p0 - p1).norm() <= 1e-3 * p0.norm()  # dA is summed by L2 in arrival order
