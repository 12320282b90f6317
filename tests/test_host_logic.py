"""CPU tests of the host-side logic that mirrors the reference (no GPU, no kernels)."""
import math, os, random

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from opentf_b200 import metric, synth, util
from opentf_b200.earlystopping import EarlyStopping
from opentf_b200.fnn import Fnn, _Plateau, loader_order
from oracle import fnn_oracle as O
from oracle import sampler_oracle as SO


def test_loader_order_matches_a_real_dataloader():
    for n, shuffle in ((17, True), (9, False), (100, True)):
        torch.manual_seed(5)
        mine = loader_order(torch, n, shuffle)
        after_mine = torch.rand(1).item()
        torch.manual_seed(5)
        dl = torch.utils.data.DataLoader(list(range(n)), batch_size=4, shuffle=shuffle)
        theirs = [int(i) for b in dl for i in b]
        after_theirs = torch.rand(1).item()
        assert (list(range(n)) if mine is None else list(mine)) == theirs
        assert after_mine == after_theirs  # generator left in the same state


def test_host_init_matches_reference_module_order():
    cfg = dict(b=4, e=1, ns=2, lr=1e-3, es=1, h=[8, 6], spe=0, l='bce', tpw=10, tnw=1, nsd='uniform')
    f = Fnn.__new__(Fnn); f.cfg = cfg
    import opentf_b200.ntf as ntf; ntf.Ntf.torch = torch
    torch.manual_seed(3); sd = f._host_init(11, 19)
    torch.manual_seed(3); ref = O.init_params(11, [8, 6], 19)
    for i, (W, b) in enumerate(ref):
        assert torch.equal(sd[f'layers.{i}.weight'], W) and torch.equal(sd[f'layers.{i}.bias'], b)


def test_plateau_matches_torch():
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1e-3)
    ref = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, factor=0.1, patience=2)
    mine, lr = _Plateau(), 1e-3
    rng = random.Random(0)
    v = 10.0
    for _ in range(60):
        v = v * (0.97 if rng.random() < 0.4 else 1.0 + 1e-5 * rng.random())
        ref.step(v); lr = mine.step(v, lr)
        assert math.isclose(lr, opt.param_groups[0]['lr'], rel_tol=1e-12)


def test_earlystopping_matches_oracle_rule():
    rng = random.Random(1)
    for trial in range(20):
        a, b = EarlyStopping(patience=3, delta=0.001, trace_func=lambda s: None), O.EarlyStop(3, 0.001)
        for _ in range(40):
            v = 50 + rng.random()
            assert a(v).early_stop == b(v)
            if b.stop: break


def test_cfg2str_is_the_reference_run_directory_name():
    cfg = dict(b=1000, e=100, ns=5, lr=0.001, es=5, h=[128], spe=10, l='bce', tpw=10, tnw=1, nsd='unigram_b')
    assert util.cfg2str(cfg) == 'b1000.e100.ns5.lr0.001.es5.h[128].spe10.lbce.tpw10.tnw1.nsdunigram_b'


def test_topk_to_sparse_equals_reference_helper():
    torch.manual_seed(0)
    P = torch.rand(7, 50)
    v, i = torch.topk(P, 5, dim=1)
    mine = util.topk_to_sparse(torch, v, i.to(torch.int32), 50)
    ref = O.topk_sparse(P, 5)
    assert torch.equal(mine.indices(), ref.indices()) and torch.equal(mine.values(), ref.values())
    sc = util.torch_sparse_2_scipy_sparse(mine, 'csr')
    assert sc.shape == (7, 50) and sc.nnz == 35


def test_metrics_match_committed_eval_csvs(toy):
    skill, member, splits, z = toy('dblp')
    trec = ['P_2,5,10', 'recall_2,5,10', 'ndcg_cut_2,5,10', 'map_cut_2,5,10', 'success_2,5,10']
    for k in range(3):
        df, mean = metric.calculate_metrics(member[splits['test']], z[f'pred/f{k}'], 1000, True, trec)
        cols = list(z[f'eval/f{k}/columns'])
        for name in df.columns: assert np.abs(df[name].values - z[f'eval/f{k}/values'][:, cols.index(name)]).max() < 1e-5
        names = list(z[f'eval/f{k}/mean_names'])
        for name in df.columns: assert abs(mean.loc[name, 'mean'] - z[f'eval/f{k}/mean_values'][names.index(name)]) < 1e-9
        auc, _ = metric.calculate_auc_roc(member[splits['test']], z[f'pred/f{k}'])
        assert abs(auc - z[f'eval/f{k}/mean_values'][names.index('aucroc')]) < 1e-9


def test_metrics_agree_with_oracle_on_sparse_predictions():
    rng = np.random.default_rng(0)
    import scipy.sparse as sp
    Y = sp.csr_matrix((rng.random((20, 60)) < 0.08).astype(np.uint8))
    P = torch.from_numpy(rng.random((20, 60)).astype(np.float32).round(2))  # rounded -> plenty of exact ties
    sparse = util.torch_sparse_2_scipy_sparse(O.topk_sparse(P, 12), 'csr')
    df, _ = metric.calculate_metrics(Y, sparse, 1000, True, ['P_2,5', 'ndcg_cut_2,5', 'map_cut_2,5', 'recall_2,5', 'success_2,5'])
    ref = O.trec_metrics(Y, sparse, ks=(2, 5), topK=1000)
    for name in df.columns: assert np.allclose(df[name].values, ref[name])


def test_synth_shapes_and_splits():
    tv = synth.make_teamsvecs('toy', seed=1)
    N, S, E = synth.SHAPES['toy'][:3]
    assert tv['skill'].shape == (N, S) and tv['member'].shape == (N, E)
    assert (np.diff(tv['skill'].tocsc().indptr) > 0).all() and (np.diff(tv['member'].tocsc().indptr) > 0).all()
    assert (np.diff(tv['member'].indptr) >= 2).all() and tv['skill'].data.max() == 1
    sp_ = synth.make_splits(N, seed=1)
    allrows = np.concatenate([sp_['test'], sp_['folds'][0]['train'], sp_['folds'][0]['valid']])
    assert sorted(allrows.tolist()) == list(range(N))


def test_sampler_oracle_distribution_matches_reference_sampler():
    """the successive-draw sampler has the distribution of the reference's multinomial-without-replacement / uniform top-ns
    (chi-square on the marginal inclusion counts, small expert set)."""
    E, ns, trials = 12, 3, 4000
    y = torch.zeros(1, E); y[0, [2, 7]] = 1
    members = [[0, 1, 2, 3], [2, 7], [3, 3 + 1, 5], [0, 5, 9], [0, 2]]  # batch rows -> counts
    counts, cdf = SO.expert_cdf(members, E)
    pool = np.concatenate([np.asarray(m) for m in members])
    uni = torch.tensor(counts / len(members), dtype=torch.float32).unsqueeze(0)
    for nsd in ('uniform', 'unigram', 'unigram_b'):
        torch.manual_seed(0)
        ref = np.zeros(E); mine = np.zeros(E)
        for t in range(trials):
            r = O.ns_uniform(y, ns) if nsd == 'uniform' else O.ns_unigram(y, uni, ns)
            ref[r.numpy().ravel()] += 1
            m = SO.sample_row(nsd, 1234, t, 0, [2, 7], E, ns, cdf, pool)
            mine[[j for j in m if j >= 0]] += 1
        assert mine[2] == 0 and mine[7] == 0 and ref[2] == 0 and ref[7] == 0
        live = ref > 0
        assert ((mine > 0) == live).all()
        chi2 = (((mine - ref) ** 2) / (mine + ref))[live].sum()
        assert chi2 < 30, (nsd, chi2, mine, ref)  # 9-10 dof; p ~ 1e-3 at 30


def test_sampler_oracle_edge_cases():
    # no candidate mass outside the team's own members -> uniform over ALL experts (fnn.py:67-69)
    out = SO.sample_row('unigram', 7, 0, 0, [1, 2], 6, 3, np.cumsum([0, 1, 1, 0, 0, 0]))
    assert len(set(out)) == 3 and all(0 <= j < 6 for j in out)
    out = SO.sample_row('unigram_b', 7, 0, 0, [1, 2], 6, 3, None, [1, 2, 2, 1])
    assert len(set(out)) == 3 and all(0 <= j < 6 for j in out)
    # fewer weighted candidates than ns -> the rest is topped up uniformly from the non-members
    out = SO.sample_row('unigram', 7, 0, 0, [1], 6, 3, np.cumsum([0, 1, 0, 0, 2, 0]))
    assert 4 in out and 1 not in out and len(set(out)) == 3
    out = SO.sample_row('unigram_b', 7, 0, 0, [1], 6, 3, None, [1, 4, 4])
    assert 4 in out and 1 not in out and len(set(out)) == 3
    # fewer negatives than ns -> padded with -1
    out = SO.sample_row('uniform', 7, 0, 0, [0, 1, 2], 5, 4, None)
    assert sorted(out) == [-1, -1, 3, 4]


def test_evaluate_hands_skill_coverage_to_the_device_function_and_writes_the_reference_csvs(toy, tmp_path, monkeypatch):
    """Ntf.evaluate with metrics.other = ['skill_coverage_2,5,10'] (ntf.py:72-78): the wiring around opentf_b200.staging.calculate_skill_coverage --
    which rows of the skill matrix, which prediction file, which co-occurrence matrix, the cut-offs, the device -- and the CSVs that come out.  The
    device function itself is replaced by the oracle here (this test runs without a GPU; tests/test_gpu_staging.py checks the real one)."""
    import pandas as pd
    import scipy.sparse as sp
    from oracle import staging_oracle as SO
    from opentf_b200 import staging
    from opentf_b200.ntf import Ntf
    skill, member, splits, z = toy('dblp')
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'staging_dblp.npz'))
    co = sp.csr_matrix((g['co/data'].astype(np.uint8), g['co/indices'], g['co/indptr']), shape=(member.shape[1], skill.shape[1]))
    seen = {}

    def fake(X, Y_, expertskillvecs, per_instance=False, topks='2,5,10', device='cuda:0'):
        seen.update(n=X.shape[0], topks=topks, device=device, same_co=(expertskillvecs != co).nnz == 0)
        cov = SO.skill_coverage(X, Y_, expertskillvecs, topks)
        df = pd.DataFrame({f'skill_coverage_{k}': v for k, v in cov.items()})
        return df, df.mean().to_frame('mean').rename_axis('metrics')
    monkeypatch.setattr(staging, 'calculate_skill_coverage', fake)
    m = Ntf(str(tmp_path), 'cuda:0', 0, {})
    torch.save({'y_pred': torch.as_tensor(z['pred/f0']), 'uncertainty': None}, f'{m.output}/f0.test.pred')
    tv = {'skill': skill.tolil(), 'member': member.tolil(), 'skillcoverage': co}
    one_fold = {'test': splits['test'], 'folds': {0: splits['folds'][0]}}
    m.evaluate(tv, one_fold, {'on_train': False, 'per_instance': True, 'per_epoch': False, 'topK': None, 'metrics': {'trec': [], 'other': ['skill_coverage_2,5,10']}})
    assert seen == {'n': len(splits['test']), 'topks': '2,5,10', 'device': 'cuda:0', 'same_co': True}
    inst = pd.read_csv(f'{m.output}/f0.test.pred.eval.instance.csv')
    for k in (2, 5, 10): assert np.allclose(inst[f'skill_coverage_{k}'].to_numpy(), g[f'pred/skill_coverage_{k}'], atol=5e-6)  # (the reference's own frame, %.5f in the file)
    mean = pd.read_csv(f'{m.output}/test.pred.eval.mean.csv', index_col=0)
    assert abs(mean.loc['skill_coverage_10', 'mean'] - g['pred/skill_coverage_10'].mean()) < 1e-9
