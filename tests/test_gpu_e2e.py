"""GPU parity tests of the whole path through the reference-facing classes (opentf_b200.fnn.Fnn -> engine -> C ABI)
against the reference's golden vectors and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import fnn_oracle as O

pytestmark = pytest.mark.gpu
TOYS = ['dblp', 'imdb', 'gith', 'uspt']


def teamsvecs(toy, key):
    skill, member, splits, z = toy(key)
    return {'skill': skill.tolil(), 'member': member.tolil()}, splits, z  # lil uint8, as the reference hands them over


def base_cfg(**over):
    c = dict(b=1000, e=100, ns=5, lr=0.001, es=5, h=[128], spe=10, l='bce', tpw=10, tnw=1, nsd='unigram_b', precision='fp32')
    c.update(over)
    return c


def make(tmp_path, cfg, seed=0):
    from opentf_b200.fnn import Fnn
    return Fnn(str(tmp_path), 'cuda:0', seed, cfg)


# ---------------------------------------------------------------------------------------------- G1
@pytest.mark.parametrize('key', TOYS)
def test_g1_committed_checkpoints_reproduce_committed_predictions(toy, tmp_path, key):
    """every committed (checkpoint, prediction) pair of the reference, through Fnn.test(): <= 1e-6 (fp32 mode)."""
    tv, splits, z = teamsvecs(toy, key)
    m = make(tmp_path, base_cfg())
    for tag in z['pairs']:
        tag = str(tag)
        fold = int(tag[1:].split('.')[0])
        sd = {k.split('/', 2)[2]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f'ckpt/{tag}/layers.')}
        torch.save({'model_state_dict': sd, 'cfg': None, 'f': fold, 'e': 0, 't_loss': 0.0, 'v_loss': 0.0}, f'{m.output}/f{fold}.pt')
        m.test(tv, {'test': splits['test'], 'folds': {fold: splits['folds'][fold]}}, dict(on_train=False, per_epoch=False, topK=1000))
        y = torch.load(f'{m.output}/f{fold}.test.pred', weights_only=False)['y_pred']
        assert not y.is_sparse and y.shape == z[f'pred/{tag}'].shape  # topK=1000 >= E -> dense, like the committed files
        assert np.abs(y.numpy() - z[f'pred/{tag}']).max() <= 1e-6, tag


def test_sparse_pred_file_matches_reference_format(toy, tmp_path):
    tv, splits, z = teamsvecs(toy, 'gith')
    m = make(tmp_path, base_cfg())
    sd = {k.split('/', 2)[2]: torch.from_numpy(z[k]) for k in z.files if k.startswith('ckpt/f0/layers.')}
    torch.save({'model_state_dict': sd}, f'{m.output}/f0.pt')
    m.test(tv, {'test': splits['test'], 'folds': {0: splits['folds'][0]}}, dict(on_train=True, per_epoch=False, topK=7))
    layers = [(sd[f'layers.{i}.weight'], sd[f'layers.{i}.bias']) for i in range(2)]
    for pred_set, rows in (('test', splits['test']), ('train', splits['folds'][0]['train']), ('valid', splits['folds'][0]['valid'])):
        y = torch.load(f'{m.output}/f0.{pred_set}.pred', weights_only=False)['y_pred']
        assert y.is_sparse and y.is_coalesced() and tuple(y.shape) == (len(rows), 346)
        ref = O.topk_sparse(O.predict(layers, toy('gith')[0], rows, 1000), 7)  # pkgmgr.topk_sparse on the reference's scores
        assert torch.equal(y.indices(), ref.indices())
        assert (y.values() - ref.values()).abs().max() <= 1e-6


# ---------------------------------------------------------------------------------------------- G2
class Replay:
    """feeds a trajectory recorded from the unmodified reference (tests/golden/traj_*.npz) into Fnn.learn"""

    def __init__(self, t, splits, b):
        self.t, self.splits, self.b, self.cursor = t, splits, b, {}

    def init(self, fold):
        return {k.split('/', 2)[2]: torch.from_numpy(self.t[k]) for k in self.t.files if k.startswith(f'f{fold}/init/')}

    def _steps(self, fold, epoch, phase):
        tr, va = len(self.splits['folds'][fold]['train']), len(self.splits['folds'][fold]['valid'])
        nb_t, nb_v = -(-tr // self.b), -(-va // self.b)
        first = epoch * (nb_t + nb_v) + (0 if phase == 'train' else nb_t)
        return first, (nb_t if phase == 'train' else nb_v)

    def order(self, fold, epoch, phase, n):
        first, nb = self._steps(fold, epoch, phase)
        ptr = self.t[f'f{fold}/rows_ptr']
        rows = self.t[f'f{fold}/rows'][ptr[first]:ptr[first + nb]]
        base = self.splits['folds'][fold][phase]
        pos = {int(r): i for i, r in enumerate(base)}
        return np.array([pos[int(r)] for r in rows])

    def neg(self, fold, epoch, phase, bi, team_rows):
        first, _ = self._steps(fold, epoch, phase)
        ptr = self.t[f'f{fold}/rows_ptr']
        assert (self.t[f'f{fold}/rows'][ptr[first + bi]:ptr[first + bi + 1]] == team_rows).all()
        return self.t[f'f{fold}/neg'][ptr[first + bi]:ptr[first + bi + 1]]


@pytest.mark.parametrize('name,key,over', [('unigram_b', 'dblp', {}), ('uniform', 'dblp', {}), ('unigram', 'dblp', {}),
                                            ('unigram_b_small', 'imdb', dict(b=4, h=[16, 8], e=6))])
def test_g2_recorded_reference_trajectory_is_retraced(toy, tmp_path, name, key, over):
    """Fnn.learn, fed the reference's recorded initial weights, batch order and sampled negatives, must reproduce the
    reference's run: same stop epoch, epoch losses to 1e-5 relative, final weights to 2e-5 absolute (fp32 mode).
    For dblp/unigram_b these are the numbers of the checkpoints COMMITTED in the reference repository."""
    tv, splits, _ = teamsvecs(toy, key)
    t = np.load(os.path.join(GOLDEN, f'traj_{key}_{name}.npz'))
    cfg = base_cfg(nsd=str(t['nsd']), **over)
    m = make(tmp_path, cfg)
    m.replay = Replay(t, splits, cfg['b'])
    m.learn(tv, splits, None)
    for k in range(3):
        ck = torch.load(f'{m.output}/f{k}.pt', weights_only=False)
        assert ck['e'] == int(t[f'f{k}/e'])
        assert abs(ck['t_loss'] - float(t[f'f{k}/t_loss'])) <= 1e-5 * ck['t_loss']
        assert abs(ck['v_loss'] - float(t[f'f{k}/v_loss'])) <= 1e-5 * ck['v_loss']
        for n_, w in ck['model_state_dict'].items():
            assert np.abs(w.numpy() - t[f'f{k}/final/{n_}']).max() < 2e-5, (k, n_)
        saved = sorted(int(f.split('.e')[1].split('.')[0]) for f in os.listdir(m.output) if f.startswith(f'f{k}.e'))
        assert saved == list(t[f'f{k}/ckpt_epochs'])  # checkpoint cadence of fnn.py:158
        # per-step losses of the last epoch against the recording
        hist = m.last_history[k]
        assert len(hist) == int(t[f'f{k}/e']) + 1


# ---------------------------------------------------------------------------------------------- free-running, no sampler
def test_learn_without_sampling_retraces_the_oracle_from_the_seed_alone(toy, tmp_path):
    """nsd unset: the reference draws nothing in bxe (fnn.py:39), so seed -> init -> shuffles line up and the whole run is
    comparable without feeding anything."""
    tv, splits, _ = teamsvecs(toy, 'gith')
    cfg = base_cfg(nsd=None, b=8, h=[32, 16], e=4, es=10, spe=2, tpw=3, tnw=1)
    m = make(tmp_path, cfg, seed=7)
    m.learn(tv, splits, None)
    import random
    random.seed(7); np.random.seed(7); torch.manual_seed(7)
    skill, member = toy('gith')[0], toy('gith')[1]
    for k in range(3):
        r = O.learn_fold(skill, member, splits['folds'][k]['train'], splits['folds'][k]['valid'], cfg)
        ck = torch.load(f'{m.output}/f{k}.pt', weights_only=False)
        assert ck['e'] == r['e']
        for (t_m, v_m), (t_o, v_o) in zip(m.last_history[k], r['history']):
            assert abs(t_m - t_o) <= 1e-5 * t_o and abs(v_m - v_o) <= 1e-5 * v_o
        for i, (W, b) in enumerate(r['layers']):
            assert (ck['model_state_dict'][f'layers.{i}.weight'] - W).abs().max() < 2e-5
            assert (ck['model_state_dict'][f'layers.{i}.bias'] - b).abs().max() < 2e-5


@pytest.mark.parametrize('nsd', ['uniform', 'unigram', 'unigram_b'])
def test_learn_with_device_sampler_is_reproducible_and_learns(toy, tmp_path, nsd):
    tv, splits, _ = teamsvecs(toy, 'gith')
    cfg = base_cfg(nsd=nsd, b=8, e=6, es=10, spe=0, h=[64])
    runs = []
    for rep in range(2):
        m = make(tmp_path / f'r{rep}', cfg, seed=3)
        m.learn(tv, {'test': splits['test'], 'folds': {0: splits['folds'][0]}}, None)
        runs.append(m.last_history[0])
    assert runs[0] == runs[1]  # same seed -> bit-identical losses (counter RNG, deterministic kernels)
    assert runs[0][-1][0] < runs[0][0][0]  # it trains


def test_evaluate_writes_reference_csvs(toy, tmp_path):
    tv, splits, z = teamsvecs(toy, 'dblp')
    m = make(tmp_path, base_cfg())
    for k in range(3):
        sd = {n.split('/', 2)[2]: torch.from_numpy(z[n]) for n in z.files if n.startswith(f'ckpt/f{k}/layers.')}
        torch.save({'model_state_dict': sd}, f'{m.output}/f{k}.pt')
    m.test(tv, splits, dict(on_train=False, per_epoch=False, topK=1000))
    evalcfg = dict(topK=1000, on_train=False, per_epoch=False, per_instance=True,
                   metrics=dict(trec=['P_2,5,10', 'recall_2,5,10', 'ndcg_cut_2,5,10', 'map_cut_2,5,10', 'success_2,5,10'], other=['aucroc']))
    m.evaluate(tv, splits, evalcfg)
    import pandas as pd
    for k in range(3):
        mean = pd.read_csv(f'{m.output}/f{k}.test.pred.eval.mean.csv', index_col=0)
        names = list(z[f'eval/f{k}/mean_names'])
        for name in mean.index:
            assert abs(mean.loc[name, 'mean'] - z[f'eval/f{k}/mean_values'][names.index(name)]) < 1e-6, (k, name)
    assert os.path.exists(f'{m.output}/test.pred.eval.mean.csv')


def test_step_host_is_the_same_step_as_the_resident_path(toy):
    """Engine.step_host (host batch -> one H2D block -> ntf_fnn_step -> loss back) against Engine.step on the staged split"""
    from opentf_b200.engine import Engine, pack_host_batch, to_csr
    skill, member, splits, _ = toy('gith')
    rows = np.asarray(splits['folds'][0]['train'])[:16]
    torch.manual_seed(0)
    layers = O.init_params(skill.shape[1], [32], member.shape[1])
    sd = {f'layers.{i}.{n}': t for i, (W, b) in enumerate(layers) for n, t in (('weight', W), ('bias', b))}
    out = []
    for mode in ('resident', 'host'):
        eng = Engine(skill.shape[1], [32], member.shape[1], 'cuda:0', precision='fp32', nsd='unigram_b', ns=5, seed=7, max_batch=16)
        eng.stage(skill, member)
        eng.load_state_dict(sd)
        if mode == 'resident':
            sp = eng.split(rows)
            for i in range(3): eng.step(sp, 0, 16, True, lr=1e-2, loss_slot=i)
            losses = eng.loss_buf[:3].cpu().numpy().copy()
        else:
            (sp_, si, _), (mp_, mi, _) = to_csr(skill[rows]), to_csr(member[rows])
            packed = pack_host_batch(sp_, si, mp_, mi)
            losses = np.array([eng.step_host(*packed, lr=1e-2) for _ in range(3)], dtype=np.float32)
        out.append((losses, eng.params.cpu().clone()))
    assert np.array_equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    assert out[0][0][2] < out[0][0][0]


def test_step_host_with_two_steps_in_flight_returns_every_loss(toy):
    """the streaming loop of bench.py's end-to-end leg: step i+1 is enqueued before the host reads loss i (pinned slot + its own event);
    losses and final parameters are those of the synchronous loop"""
    from opentf_b200.engine import Engine, HostPacker, to_csr
    skill, member, splits, _ = toy('gith')
    train = np.asarray(splits['folds'][0]['train'])
    n, steps = 16, 7
    torch.manual_seed(0)
    layers = O.init_params(skill.shape[1], [32], member.shape[1])
    sd = {f'layers.{i}.{nm}': t for i, (W, b) in enumerate(layers) for nm, t in (('weight', W), ('bias', b))}
    s_csr, m_csr = to_csr(skill), to_csr(member)
    out = []
    for mode in ('sync', 'pipelined'):
        eng = Engine(skill.shape[1], [32], member.shape[1], 'cuda:0', precision='fp32', nsd='unigram_b', ns=5, seed=7, max_batch=n)
        eng.stage(skill, member)
        eng.load_state_dict(sd)
        packer = HostPacker(s_csr, m_csr, n, 512, 512)
        rows = lambda i: train[(i * n) % (len(train) - n):][:n]
        losses = []
        if mode == 'sync':
            for i in range(steps): losses.append(eng.step_host(packer.pack(rows(i)), n, 512, 512, lr=1e-2))
        else:
            eng.step_host(packer.pack(rows(0)), n, 512, 512, lr=1e-2, sync=False, slot=0)
            for i in range(1, steps):
                eng.step_host(packer.pack(rows(i)), n, 512, 512, lr=1e-2, sync=False, slot=i & 1)
                losses.append(eng.step_host_loss((i - 1) & 1))
            losses.append(eng.step_host_loss((steps - 1) & 1))
        out.append((np.asarray(losses, dtype=np.float32), eng.params.cpu().clone()))
    assert np.array_equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    assert len(set(out[0][0].tolist())) == steps  # (seven different batches, seven different losses: no slot was read twice)


# ---------------------------------------------------------------------------------------------- tNtf: year-by-year fine-tuning
def test_tntf_chain_retraces_the_reference_run(toy, tmp_path):
    """opentf_b200.tntf.tNtf around opentf_b200.fnn.Fnn, free-running from seed 0 (nsd unset: nothing is drawn in bxe), against the run of
    the reference's OWN tNtf(Fnn) recorded in tests/golden/tntf_gith.npz: per year and fold the same stop epoch, epoch losses to 1e-5
    relative and final weights to 2e-5 -- every year starts from the previous year's checkpoint (fnn.py:101) and the host generator is
    consumed in the reference's order across the whole chain."""
    from opentf_b200.tntf import tNtf
    tv, _, _ = teamsvecs(toy, 'gith')
    t = np.load(os.path.join(GOLDEN, 'tntf_gith.npz'))
    year_idx = [tuple(int(v) for v in r) for r in t['year_idx']]
    cfg = base_cfg(nsd=None, b=8, e=5, h=[16], lr=0.01, es=5, spe=0)
    inner = make(tmp_path, cfg)
    w = tNtf(str(tmp_path), 'cuda:0', 0, dict(tfolds=2, step_ahead=1), inner, year_idx)
    splits = {'test': np.arange(year_idx[-1][0], tv['skill'].shape[0]), 'folds': {k: {'train': None, 'valid': None} for k in range(2)}}
    w.learn(tv, splits, None)
    for _, year in year_idx[:-1]:
        for k in range(2):
            ck = torch.load(f'{w.output}/{year}/f{k}.pt', weights_only=False)
            assert ck['e'] == int(t[f'{year}/f{k}/e'])
            assert abs(ck['t_loss'] - float(t[f'{year}/f{k}/t_loss'])) <= 1e-5 * ck['t_loss']
            assert abs(ck['v_loss'] - float(t[f'{year}/f{k}/v_loss'])) <= 1e-5 * ck['v_loss']
            for n_, wt in ck['model_state_dict'].items():
                assert np.abs(wt.numpy() - t[f'{year}/f{k}/{n_}']).max() < 2e-5, (year, k, n_)
    w.test(tv, splits, dict(on_train=False, per_epoch=False, topK=None))  # the wrapped model tests from the LAST year's directory
    assert os.path.exists(f'{w.output}/{year_idx[-2][1]}/f0.test.pred')
