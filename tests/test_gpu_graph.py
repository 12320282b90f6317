"""CUDA-graph replay of a step (ntf_graph_* + ntf_dyn, include/ntf_b200.h) against the same steps enqueued launch by launch:
the graph is a host-side optimisation only, so in the deterministic fp32 mode every loss and every parameter must come out
bit-identical, across epochs (re-gathered data under the same pointers, advancing RNG counter, Adam step count, changing lr)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _engine(tv, precision, graphs, B, nsd='unigram_b', h=128, seed=3):
    from opentf_b200.engine import Engine
    from oracle import fnn_oracle as O
    S, E = tv['skill'].shape[1], tv['member'].shape[1]
    eng = Engine(S, [h], E, 'cuda:0', precision=precision, tpw=10, tnw=1, nsd=nsd, ns=5, seed=seed, max_batch=B)
    eng.use_graphs = graphs
    eng.stage(tv['skill'], tv['member'])
    torch.manual_seed(0)
    layers = O.init_params(S, [h], E)
    eng.load_state_dict({f'layers.{i}.{n}': t for i, (W, b) in enumerate(layers) for n, t in (('weight', W), ('bias', b))})
    return eng


def _run_epochs(eng, rows, vrows, B, epochs):
    sp, vsp = eng.split(rows), eng.split(vrows)
    nb, nbv = -(-sp.n // B), -(-vsp.n // B)
    rng = np.random.default_rng(5)
    out = []
    for e in range(epochs):
        sp.regather(rng.permutation(sp.n))
        lr = 1e-2 * (0.1 ** (e // 2))  # the plateau scheduler changes lr between epochs
        for bi in range(nb):
            eng.step(sp, bi * B, min(B, sp.n - bi * B), True, lr=lr, loss_slot=bi)
        for bi in range(nbv):  # validation batches sample negatives too (fnn.py:148) and share the step counter
            eng.step(vsp, bi * B, min(B, vsp.n - bi * B), False, loss_slot=nb + bi)
        out.append(eng.loss_buf[:nb + nbv].cpu().numpy().copy())
    return np.stack(out), eng.params.cpu().clone()


@pytest.mark.parametrize('nsd', ['unigram_b', 'uniform', 'unigram'])
def test_graph_replay_is_bit_identical_to_direct_launches_fp32(nsd):
    from opentf_b200 import synth, _lib
    tv = synth.make_teamsvecs('toy', seed=2)
    rows, vrows = np.arange(0, 300), np.arange(300, 420)  # 300 = 4 full batches of 64 + a short one
    res = []
    for graphs in (False, True):
        eng = _engine(tv, 'fp32', graphs, 64, nsd=nsd, h=32)
        if nsd == 'unigram': eng.set_global_unigram()
        _lib.lib().ntf_launch_count(1)
        res.append(_run_epochs(eng, rows, vrows, 64, 4) + (int(_lib.lib().ntf_launch_count(0)), len(eng._graphs)))
    (l0, p0, n0, g0), (l1, p1, n1, g1) = res
    assert g0 == 0 and g1 == 5 + 2  # one graph per distinct batch, captured in epoch 0 and replayed in epochs 1..3
    assert np.array_equal(l0, l1)
    assert torch.equal(p0, p1)
    assert n1 == n0 + 4 * 7  # same kernels, plus one ntf_dyn_update per step
    assert l0[-1, :5].mean() < l0[0, :5].mean()


def test_graph_replay_tensor_core_mode():
    """the tcgen05 path sums dA in L2 arrival order, so two runs agree to round-off, not bit for bit"""
    from opentf_b200 import synth, _lib
    tv = synth.make_teamsvecs('toy', seed=2, n_teams=1200, n_experts=700)
    rows, vrows = np.arange(0, 1000), np.arange(1000, 1200)
    res = []
    for graphs in (False, True):
        eng = _engine(tv, 'tf32', graphs, 256)
        if eng.precision != _lib.NTF_TF32: pytest.skip('no tensor-core kernel for this shape')
        res.append(_run_epochs(eng, rows, vrows, 256, 3))
    (l0, p0), (l1, p1) = res
    assert np.allclose(l0, l1, rtol=2e-4, atol=0)
    assert (p0 - p1).abs().max().item() <= 5e-3 * p0.abs().max().item()


def test_step_host_replays_one_graph_for_every_batch_of_a_layout():
    """host batches packed to fixed capacities land at the same device addresses: one captured graph serves them all, and the
    losses / parameters equal the resident path's"""
    from opentf_b200 import synth
    from opentf_b200.engine import pack_host_batch, to_csr
    tv = synth.make_teamsvecs('toy', seed=4)
    B = 48
    batches = [np.arange(i * B, (i + 1) * B) for i in range(5)]
    out = []
    for mode in ('resident', 'host'):
        eng = _engine(tv, 'fp32', True, B, h=32)
        if mode == 'resident':
            eng.use_graphs = False
            for i in range(5): eng.step(eng.split(batches[i]), 0, B, True, lr=1e-2, loss_slot=i)  # (row 0 of its split, like a host batch)
            losses = eng.loss_buf[:5].cpu().numpy().copy()
        else:
            packs = []
            for r in batches:
                (sp_, si, _), (mp_, mi, _) = to_csr(tv['skill'][r]), to_csr(tv['member'][r])
                packs.append((sp_, si, mp_, mi))
            cap_s, cap_m = max(len(p[1]) for p in packs), max(len(p[3]) for p in packs)
            losses = np.array([eng.step_host(*pack_host_batch(*p, cap_s=cap_s, cap_m=cap_m), lr=1e-2) for p in packs], dtype=np.float32)
            assert len(eng._graphs) == 1
        out.append((losses, eng.params.cpu().clone()))
    assert np.array_equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])


@pytest.mark.parametrize('precision', ['fp32', 'tf32'])
def test_bnn_step_graph_replay_equals_direct_launches(precision):
    """the Bnn (Flipout) step is ~45 host-driven calls; captured per batch (ntf_set_dyn feeds the RNG counter and Adam's constants from
    the device block) it must reproduce the eager run: bit for bit in fp32 mode, to round-off on the tensor-core path (dA / split-tile
    sums arrive in L2 order).  Device-side noise: same counters, same draws."""
    from opentf_b200 import synth, _lib
    from opentf_b200.engine import Engine
    tv = synth.make_teamsvecs('toy', seed=2)
    S, E, B, h = tv['skill'].shape[1], tv['member'].shape[1], 128, 128
    if precision == 'tf32' and not _lib.lib().ntf_tc_supported(B, h, E, 1): pytest.skip('shape not on the tensor-core path')
    g = torch.Generator().manual_seed(0)
    sd = {}
    for i, (fin, fout) in enumerate(((S, h), (h, E))):
        sd[f'layers.{i}.mu_weight'] = torch.empty(fout, fin).normal_(0.0, 0.1, generator=g); sd[f'layers.{i}.rho_weight'] = torch.empty(fout, fin).normal_(-3.0, 0.1, generator=g)
        sd[f'layers.{i}.mu_bias'] = torch.empty(fout).normal_(0.0, 0.1, generator=g); sd[f'layers.{i}.rho_bias'] = torch.empty(fout).normal_(-3.0, 0.1, generator=g)
    res = []
    for graphs in (False, True):
        eng = Engine(S, [h], E, 'cuda:0', bayesian=True, precision=precision, tpw=10, tnw=1, nsd='unigram_b', ns=5, seed=3, max_batch=B)
        eng.use_graphs = graphs
        eng.stage(tv['skill'], tv['member'])
        eng.load_state_dict(sd)
        sp, vsp = eng.split(np.arange(0, 300)), eng.split(np.arange(300, 420))
        out = []
        for e in range(3):
            lr = 1e-2 * (0.1 ** e)
            for bi in range(3): eng.step(sp, bi * 100, 100, True, lr=lr, loss_slot=bi)
            eng.step(vsp, 0, 120, False, loss_slot=3)
            out.append(eng.loss_buf[:4].cpu().clone())
        torch.cuda.synchronize()
        res.append((torch.stack(out), eng.params.cpu().clone(), eng.adam_t, eng.global_step, len(eng._graphs)))
    (l0, p0, t0, s0, g0), (l1, p1, t1, s1, g1) = res
    assert t0 == t1 == 9 and s0 == s1 == 12
    assert g0 == 0 and g1 == 4  # 3 train batches + 1 validation batch, captured at their first occurrence after the warm-up step
    if precision == 'fp32': assert torch.equal(l0, l1) and torch.equal(p0, p1)
    else: assert torch.allclose(l0, l1, rtol=1e-4) and (p0 - p1).norm() <= 1e-3 * p0.norm()
    assert l0[-1, :3].mean() < l0[0, :3].mean()
