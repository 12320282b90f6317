"""ntf_eval_ranked (on-device trec metrics) against the host restatement of metric.py / trec_eval
(`opentf_b200.metric.calculate_metrics`, itself pinned on the reference's committed eval CSVs by test_oracle_golden G3)."""
import numpy as np
import pytest
import scipy.sparse as sp

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from opentf_b200 import metric  # noqa: E402

TREC = ['P_2,5,10', 'recall_2,5,10', 'ndcg_cut_2,5,10', 'map_cut_2,5,10', 'success_2,5,10']


def _members(rng, N, E, lo=1, hi=6):
    rows, cols = [], []
    for i in range(N):
        c = rng.choice(E, size=rng.integers(lo, hi), replace=False)
        rows += [i] * len(c); cols += c.tolist()
    return sp.csr_matrix((np.ones(len(rows), np.uint8), (rows, cols)), shape=(N, E))


def _sparse_pred(rng, N, E, K, Y, ties):
    """top-K style sparse predictions: K distinct columns per row (fewer on some rows), the true members mixed in, values quantised when
    `ties` so that equal scores (ranked by document-id STRING) are common"""
    rows, cols, vals = [], [], []
    for i in range(N):
        k = K if i % 7 else max(1, K // 3)
        truth = Y.indices[Y.indptr[i]:Y.indptr[i + 1]]
        c = np.unique(np.concatenate([rng.choice(E, size=k, replace=False), truth[:rng.integers(0, len(truth) + 1)]]))[:k]
        v = rng.random(len(c)).astype(np.float32)
        if ties: v = (np.floor(v * 4) / 4 + 0.125).astype(np.float32)
        rows += [i] * len(c); cols += c.tolist(); vals += v.tolist()
    return sp.csr_matrix((np.asarray(vals, np.float32), (rows, cols)), shape=(N, E))


@pytest.mark.parametrize('N,E,K,ties', [(64, 50, 12, False), (200, 123457, 100, True), (33, 2000, 1000, True), (5, 9, 3, False)])
def test_sparse_candidates_match_host_metrics(N, E, K, ties):
    rng = np.random.default_rng(N + K)
    Y = _members(rng, N, E)
    Y_ = _sparse_pred(rng, N, E, K, Y, ties)
    df_h, mean_h = metric.calculate_metrics(Y, Y_, 1000, True, TREC)
    df_d, mean_d = metric.calculate_metrics_device(Y, Y_, 1000, True, TREC, device='cuda:0', chunk=37)
    assert list(df_h.columns) == list(df_d.columns)
    assert np.abs(df_h.values - df_d.values).max() < 1e-12
    assert np.abs(mean_h.values - mean_d.values).max() < 1e-12
    assert df_h['success_10'].sum() > 0  # the case is not vacuous


def test_dense_predictions_first_stage_on_device():
    """dense .pred (fnn.py:218 with topK >= E or None) wider than the kernel's candidate list: ntf_topk_select does metric.py:12's first stage"""
    rng = np.random.default_rng(5)
    N, E = 40, 3000
    Y = _members(rng, N, E)
    Y_ = rng.random((N, E)).astype(np.float32)
    for i in range(N): Y_[i, Y.indices[Y.indptr[i]:Y.indptr[i + 1]]] += 0.5 * (i % 3)
    for topK in (None, 1000, 50):
        df_h, _ = metric.calculate_metrics(Y, Y_, topK, True, TREC)
        df_d, _ = metric.calculate_metrics_device(Y, Y_, topK, True, TREC, device='cuda:0')
        assert np.abs(df_h.values - df_d.values).max() < 1e-12, topK


def test_toy_goldens_through_the_device_path(toy):
    """G3: the reference's committed f{k}.test.pred.eval.instance.csv from its committed predictions"""
    skill, member, splits, z = toy('dblp')
    Y = member[splits['test']]
    for k in range(3):
        if f'pred/f{k}' not in z.files: pytest.skip('golden file holds no predictions')
        Y_ = z[f'pred/f{k}']
        df, mean = metric.calculate_metrics_device(Y, Y_, 1000, True, TREC, device='cuda:0')
        cols = list(z[f'eval/f{k}/columns']); vals = z[f'eval/f{k}/values']
        for name in df.columns:
            assert np.abs(df[name].values - vals[:, cols.index(name)]).max() < 5.1e-6, (k, name)  # the CSVs are written with %.5f


def test_bad_arguments_are_refused():
    from opentf_b200 import ops, _lib
    dev = 'cuda:0'
    idx = torch.zeros(2, 4, dtype=torch.int32, device=dev); vals = torch.zeros(2, 4, device=dev)
    ptr = torch.tensor([0, 1, 2], dtype=torch.int32, device=dev); mi = torch.tensor([0, 1], dtype=torch.int32, device=dev)
    with pytest.raises(_lib.NtfError): ops.eval_ranked(idx, vals, ptr, mi, [5, 2], torch.empty(2, 5, 2, dtype=torch.float64, device=dev))
    big = torch.zeros(1, 2000, device=dev)
    with pytest.raises(_lib.NtfError):
        ops.eval_ranked(torch.zeros(1, 2000, dtype=torch.int32, device=dev), big, ptr[:2].contiguous(), mi, [2], torch.empty(1, 5, 1, dtype=torch.float64, device=dev))
