"""Pin the CPU oracle (oracle/fnn_oracle.py) against the reference's golden vectors."""
import os, random

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import fnn_oracle as O

TOYS = ['dblp', 'imdb', 'gith', 'uspt']


def ckpt_layers(z, tag):
    n = len([k for k in z.files if k.startswith(f'ckpt/{tag}/layers.') and k.endswith('.weight')])
    return [(torch.from_numpy(z[f'ckpt/{tag}/layers.{i}.weight']), torch.from_numpy(z[f'ckpt/{tag}/layers.{i}.bias'])) for i in range(n)]


@pytest.mark.parametrize('key', TOYS)
def test_g1_checkpoint_to_prediction(toy, key):
    """every committed (f{k}[.eN].pt, f{k}.test[.eN].pred) pair: 43 pairs over the 4 toy datasets."""
    skill, member, splits, z = toy(key)
    pairs = list(z['pairs'])
    assert len(pairs) == {'dblp': 9, 'imdb': 10, 'gith': 15, 'uspt': 9}[key]
    for tag in pairs:
        p = O.predict(ckpt_layers(z, tag), skill, splits['test'], 1000).numpy()
        assert p.shape == z[f'pred/{tag}'].shape
        assert np.abs(p - z[f'pred/{tag}']).max() <= 5e-7, tag


def _seed(s):
    random.seed(s); np.random.seed(s); torch.manual_seed(s)


@pytest.mark.parametrize('name,key,over', [('unigram_b', 'dblp', {}), ('uniform', 'dblp', {}), ('unigram', 'dblp', {}),
                                            ('unigram_b_small', 'imdb', dict(b=4, h=[16, 8], e=6))])
def test_g2_free_running_trajectory(toy, name, key, over):
    """seed 0, the oracle draws its own randomness in the reference's order and must retrace the
    recorded reference run step by step (bit-equal negatives, losses to fp32 round-off)."""
    skill, member, splits, _ = toy(key)
    t = np.load(os.path.join(GOLDEN, f'traj_{key}_{name}.npz'))
    nsd = str(t['nsd'])
    cfg = dict(b=1000, e=100, ns=5, lr=0.001, es=5, h=[128], spe=10, tpw=10, tnw=1, nsd=nsd); cfg.update(over)
    _seed(0)
    uni = O.global_unigram(member) if nsd == 'unigram' else None
    for k in range(3):
        trace = []
        r = O.learn_fold(skill, member, splits['folds'][k]['train'], splits['folds'][k]['valid'], cfg, trace=trace, unigram=uni)
        assert r['e'] == int(t[f'f{k}/e'])
        assert len(trace) == len(t[f'f{k}/loss'])
        ptr = t[f'f{k}/rows_ptr']
        for i, (e, ph, rows, neg, loss) in enumerate(trace):
            assert (rows == t[f'f{k}/rows'][ptr[i]:ptr[i + 1]]).all()
            assert (neg.numpy() == t[f'f{k}/neg'][ptr[i]:ptr[i + 1]]).all()
            assert abs(loss - t[f'f{k}/loss'][i]) <= 2e-5 * abs(loss), (k, i)
        assert abs(r['t_loss'] - float(t[f'f{k}/t_loss'])) < 1e-4 and abs(r['v_loss'] - float(t[f'f{k}/v_loss'])) < 1e-4
        for i, (W, b) in enumerate(r['layers']):
            assert np.abs(W.numpy() - t[f'f{k}/final/layers.{i}.weight']).max() < 2e-5
            assert np.abs(b.numpy() - t[f'f{k}/final/layers.{i}.bias']).max() < 2e-5
        assert [e for e, _ in r['ckpts']] == list(t[f'f{k}/ckpt_epochs'])


def iter_feed(t, k):
    ptr, has_neg = t[f'f{k}/rows_ptr'], f'f{k}/neg' in t.files
    for i in range(len(ptr) - 1):
        rows = t[f'f{k}/rows'][ptr[i]:ptr[i + 1]]
        neg = t[f'f{k}/neg'][ptr[i]:ptr[i + 1]] if has_neg else None
        yield ('train' if t[f'f{k}/phase'][i] == 0 else 'valid'), rows, neg


@pytest.mark.parametrize('name,key,over', [('unigram_b', 'dblp', {}), ('unigram_b_small', 'imdb', dict(b=4, h=[16, 8], e=6))])
def test_g2_replay_with_fed_indices(toy, name, key, over):
    """host-supplied batches + negatives (the contract the CUDA path is tested under) give the same run."""
    skill, member, splits, _ = toy(key)
    t = np.load(os.path.join(GOLDEN, f'traj_{key}_{name}.npz'))
    cfg = dict(b=1000, e=100, ns=5, lr=0.001, es=5, h=[128], spe=10, tpw=10, tnw=1, nsd='unigram_b'); cfg.update(over)
    for k in range(3):
        nl = len([n for n in t.files if n.startswith(f'f{k}/init/') and n.endswith('.weight')])
        cfg['init_layers'] = [(torch.from_numpy(t[f'f{k}/init/layers.{i}.weight']), torch.from_numpy(t[f'f{k}/init/layers.{i}.bias'])) for i in range(nl)]
        r = O.learn_fold(skill, member, splits['folds'][k]['train'], splits['folds'][k]['valid'], cfg, feed=iter_feed(t, k))
        assert r['e'] == int(t[f'f{k}/e'])
        assert abs(r['t_loss'] - float(t[f'f{k}/t_loss'])) < 1e-4 and abs(r['v_loss'] - float(t[f'f{k}/v_loss'])) < 1e-4
        for i, (W, b) in enumerate(r['layers']):
            assert np.abs(W.numpy() - t[f'f{k}/final/layers.{i}.weight']).max() < 2e-5


def test_handwritten_gradients_match_autograd():
    torch.manual_seed(1)
    B, S, E = 7, 11, 19
    layers = O.init_params(S, [8, 6], E)
    X = (torch.rand(B, S) < 0.3).float(); y = (torch.rand(B, E) < 0.2).float()
    neg = torch.randint(0, E, (B, 3))
    w = O.loss_weights(y, neg, 10, 1)
    logits, acts, pre = O.forward(layers, X)
    grads = O.backward(layers, acts, pre, y, w)
    ps = [(W.clone().requires_grad_(), b.clone().requires_grad_()) for W, b in layers]
    a = X
    for W, b in ps: a = torch.nn.functional.leaky_relu(torch.nn.functional.linear(a, W, b))
    loss = torch.nn.functional.binary_cross_entropy_with_logits(a, y, w, reduction='none').sum(1).mean()
    assert abs(loss.item() - O.bce_with_logits(logits, y, w).sum(1).mean().item()) < 1e-5
    loss.backward()
    for (gW, gb), (W, b) in zip(grads, ps):
        assert torch.allclose(gW, W.grad, atol=1e-6) and torch.allclose(gb, b.grad, atol=1e-6)


def test_adam_restatement_matches_torch():
    torch.manual_seed(2)
    p0 = [torch.randn(5, 4), torch.randn(4)]
    mine = [t.clone() for t in p0]; theirs = [t.clone().requires_grad_() for t in p0]
    a, b = O.Adam(mine, 1e-3), torch.optim.Adam(theirs, lr=1e-3)
    for _ in range(25):
        g = [torch.randn_like(t) for t in p0]
        a.step(g)
        for t, gg in zip(theirs, g): t.grad = gg.clone()
        b.step()
    for m, t in zip(mine, theirs): assert torch.equal(m, t.detach())


def test_flipout_gradients_match_autograd():
    torch.manual_seed(3)
    B, S, E = 6, 9, 14
    L = O.init_flipout_params(S, [5], E)
    X = (torch.rand(B, S) < 0.4).float(); y = (torch.rand(B, E) < 0.2).float()
    w = O.loss_weights(y, torch.randint(0, E, (B, 2)), 10, 1)
    Nz = O.draw_flipout_noise(L, B)
    logits, acts, pre = O.flipout_forward(L, Nz, X)
    grads = O.flipout_backward(L, Nz, acts, pre, y, w)
    P = [{k: v.clone().requires_grad_() for k, v in l.items()} for l in L]
    lg, _, _ = O.flipout_forward(P, Nz, X)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(lg, y, w, reduction='none').sum(1).mean() + O.flipout_kl(P) / B
    loss.backward()
    for g, p in zip(grads, P):
        for k in g: assert torch.allclose(g[k], p[k].grad, atol=2e-6), k


def test_g4_bnn_layout_and_statistics(toy):
    _, _, _, z = toy('dblp')
    assert list(z['bnn/names']) == [f'layers.{i}.{n}' for i in (0, 1) for n in ('mu_weight', 'rho_weight', 'mu_bias', 'rho_bias')]
    torch.manual_seed(0)
    L = O.init_flipout_params(10, [128], 13)
    for i, l in enumerate(L):
        for mine, theirs in (('mu_w', 'mu_weight'), ('rho_w', 'rho_weight'), ('mu_b', 'mu_bias'), ('rho_b', 'rho_bias')):
            assert tuple(l[mine].shape) == tuple(z[f'bnn/f0/layers.{i}.{theirs}'].shape)
    # KL scale: a sum of 4 tensor MEANS, about 2.5 each at init
    assert 9.0 < float(O.flipout_kl(L)) < 11.0
    # uncertainty bookkeeping: p ~ 0.5 over E=13 experts -> predictive entropy ~ 13*0.5*ln2 = 4.5
    mc = np.stack([z['bnn/f0/y_pred']] * 3)
    assert np.allclose(O.predictive_entropy(mc), z['bnn/f0/unc_pred'], atol=0.05)


def test_g3_metrics_match_committed_eval_csv(toy):
    skill, member, splits, z = toy('dblp')
    for k in range(3):
        cols = list(z[f'eval/f{k}/columns']); vals = z[f'eval/f{k}/values']
        m = O.trec_metrics(member[splits['test']], z[f'pred/f{k}'], ks=(2, 5, 10), topK=1000)
        for name, arr in m.items():
            assert np.abs(arr - vals[:, cols.index(name)]).max() < 1e-5, (k, name)
        names = list(z[f'eval/f{k}/mean_names'])
        for name, arr in m.items():
            assert abs(arr.mean() - z[f'eval/f{k}/mean_values'][names.index(name)]) < 1e-9


# ---------------------------------------------------------------------------------------------- dense (embedded) skill input
def test_g2_dense_skill_input_trajectory(toy):
    """SURVEY 8f-2 / ntf.py:24: teamsvecs['skill'] as a dense [N,d] matrix of skill embeddings.  The recording is the UNMODIFIED
    reference run on seeded dense vectors (tests/golden/make_golden.py dense); the oracle must retrace it free-running
    (bit-equal negatives) and replayed with fed indices."""
    _, member, splits, _ = toy('gith')
    t = np.load(os.path.join(GOLDEN, 'traj_gith_dense16.npz'))
    X = t['dense_x']
    assert X.shape == (member.shape[0], int(t['dense_d']))
    cfg = dict(b=32, e=8, ns=5, lr=0.01, es=5, h=[24], spe=10, tpw=10, tnw=1, nsd='unigram_b')
    _seed(0)
    for k in range(3):
        trace = []
        r = O.learn_fold(X, member, splits['folds'][k]['train'], splits['folds'][k]['valid'], cfg, trace=trace)
        assert r['e'] == int(t[f'f{k}/e']) and len(trace) == len(t[f'f{k}/loss'])
        ptr = t[f'f{k}/rows_ptr']
        for i, (e, ph, rows, neg, loss) in enumerate(trace):
            assert (rows == t[f'f{k}/rows'][ptr[i]:ptr[i + 1]]).all()
            assert (neg.numpy() == t[f'f{k}/neg'][ptr[i]:ptr[i + 1]]).all()
            assert abs(loss - t[f'f{k}/loss'][i]) <= 2e-5 * abs(loss), (k, i)
        for i, (W, b) in enumerate(r['layers']):
            assert np.abs(W.numpy() - t[f'f{k}/final/layers.{i}.weight']).max() < 2e-5
    k = 0
    cfg['init_layers'] = [(torch.from_numpy(t[f'f{k}/init/layers.{i}.weight']), torch.from_numpy(t[f'f{k}/init/layers.{i}.bias'])) for i in range(2)]
    r = O.learn_fold(X, member, splits['folds'][k]['train'], splits['folds'][k]['valid'], cfg, feed=iter_feed(t, k))
    assert abs(r['t_loss'] - float(t[f'f{k}/t_loss'])) < 1e-4 * r['t_loss']
    for i, (W, b) in enumerate(r['layers']):
        assert np.abs(W.numpy() - t[f'f{k}/final/layers.{i}.weight']).max() < 2e-5


# ---------------------------------------------------------------------------------------------- tNtf: year-by-year warm starts
def test_tntf_chain_recorded_from_the_reference(toy):
    """tntf.py:17-38 around fnn.py:99-101: every interval trains k folds starting from the previous interval's checkpoints.  The recording
    (tests/golden/tntf_gith.npz) is the reference's own tNtf(Fnn) run, nsd unset, seed 0; the oracle retraces it from the seed alone."""
    skill, member, _, _ = toy('gith')
    t = np.load(os.path.join(GOLDEN, 'tntf_gith.npz'))
    cfg = dict(b=8, e=5, ns=5, lr=0.01, es=5, h=[16], spe=0, tpw=10, tnw=1, nsd=None)
    _seed(0)
    prev = {}
    for _, year in t['year_idx'][:-1]:
        for k in range(2):
            c = dict(cfg, prev_layers=prev.get(k))
            r = O.learn_fold(skill, member, t[f'{year}/f{k}/train'], t[f'{year}/f{k}/valid'], c)
            assert r['e'] == int(t[f'{year}/f{k}/e'])
            assert abs(r['t_loss'] - float(t[f'{year}/f{k}/t_loss'])) < 1e-5 * r['t_loss']
            for i, (W, b) in enumerate(r['layers']):
                assert np.abs(W.numpy() - t[f'{year}/f{k}/layers.{i}.weight']).max() < 2e-5, (year, k, i)
            prev[k] = r['layers']


def test_tntf_wrapper_host_logic(tmp_path):
    """opentf_b200.tntf.tNtf: interval row ranges, per-year KFold, run directories, splits.pkl, prev_model chaining and the resume rule,
    with a recording stand-in for the wrapped model -- and the folds must be the ones the reference's tNtf produced."""
    import pickle
    from opentf_b200.tntf import tNtf
    t = np.load(os.path.join(GOLDEN, 'tntf_gith.npz'))
    year_idx = [tuple(int(v) for v in r) for r in t['year_idx']]

    class Inner:
        def __init__(self, out): self.output, self.calls = out, []
        def learn(self, tv, splits, prev): self.calls.append((self.output, {k: (v['train'].copy(), v['valid'].copy()) for k, v in splits['folds'].items()}, prev))

    inner = Inner(str(tmp_path))
    w = tNtf(str(tmp_path), 'cpu', 0, dict(tfolds=2, step_ahead=1), inner, year_idx)
    splits = {'test': np.arange(year_idx[-1][0], 40), 'folds': {k: {'train': None, 'valid': None} for k in range(2)}}
    w.learn({}, splits, None)
    assert [os.path.basename(c[0]) for c in inner.calls] == ['2000', '2001', '2002']
    for (out, folds, prev), (_, year) in zip(inner.calls, year_idx[:-1]):
        for k in range(2):
            assert (folds[k][0] == t[f'{year}/f{k}/train']).all() and (folds[k][1] == t[f'{year}/f{k}/valid']).all()
        assert os.path.exists(f'{out}/splits.pkl') and set(pickle.load(open(f'{out}/splits.pkl', 'rb'))['folds']) == {0, 1}
    assert inner.calls[0][2] is None and inner.calls[1][2] == {0: f'{tmp_path}/2000/f0.pt', 1: f'{tmp_path}/2000/f1.pt'}
    # resume (tntf.py:19-27): with the directories of 2000..2002 present, the first two intervals are skipped
    inner2 = Inner(str(tmp_path))
    tNtf(str(tmp_path), 'cpu', 0, dict(tfolds=2, step_ahead=1), inner2, year_idx).learn({}, splits, None)
    assert [os.path.basename(c[0]) for c in inner2.calls] == ['2002']
