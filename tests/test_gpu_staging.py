"""GPU parity tests of SURVEY 8(f-4) (csrc/staging.cu through opentf_b200.staging / ops): id lists -> multi-hot CSR rows (team.py:148-173),
member^T . skill (team.py:302-341) and the skill-coverage loop (metric.py:44-73).  Integer work: bit-exact against the oracle and against the
fixtures recorded from the unmodified reference; the coverage ratios are the same fp64 division of the same two integers."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from conftest import GOLDEN
from oracle import staging_oracle as SO
from test_gpu_kernels import DEV
from test_staging_oracle import KEYS, golden, messy_lists, same_csr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def staging():
    from opentf_b200 import staging as _s
    return _s


def rand_multihot(rng, n, n_cols, lo, hi, zipf=False):
    rows = []
    for _ in range(n):
        k = int(rng.integers(lo, hi + 1))
        c = (np.minimum(rng.zipf(1.3, size=k) - 1, n_cols - 1) if zipf else rng.integers(0, n_cols, size=k))
        rows.append(np.unique(c))
    indptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])])
    ind = np.concatenate(rows) if indptr[-1] else np.zeros(0, np.int64)
    return sp.csr_matrix((np.ones(len(ind), np.uint8), ind, indptr), shape=(n, n_cols))


@pytest.mark.parametrize('key', KEYS)
def test_lists_to_rows_rebuild_the_committed_teamsvecs(staging, toy, key):
    skill, member, _, _ = toy(key)
    rng = np.random.default_rng(3)
    for M in (skill, member):
        indptr, ids = messy_lists(M, rng)
        R = staging.csr_from_lists(indptr, ids, M.shape[1], DEV)
        assert R.dtype == np.uint8 and R.has_sorted_indices and (R != sp.csr_matrix(M)).nnz == 0


@pytest.mark.parametrize('n,n_cols,lo,hi', [(1, 5, 0, 0), (257, 40, 0, 70), (3000, 30000, 0, 12), (500, 213317, 20, 400), (64, 1500000, 0, 3000)])
def test_lists_to_rows_match_the_oracle_incl_empty_duplicated_and_long_rows(staging, n, n_cols, lo, hi):
    rng = np.random.default_rng(n + n_cols)
    M = rand_multihot(rng, n, n_cols, lo, hi)
    indptr, ids = messy_lists(M, rng)
    R = staging.csr_from_lists(indptr, ids, n_cols, DEV)
    assert np.array_equal(R.indptr, M.indptr) and np.array_equal(R.indices, M.indices)
    if n * n_cols <= 1 << 22:  # the literal restatement (a dense row per team) where it is affordable
        assert (R != SO.rows_from_lists(indptr, ids, n_cols)).nnz == 0


def test_lists_with_an_id_out_of_range_are_refused(staging):
    with pytest.raises(RuntimeError, match='outside'):
        staging.csr_from_lists([0, 2], [1, 9], 9, DEV)


@pytest.mark.parametrize('key', KEYS)
def test_cooccurrence_is_the_reference_matrix(staging, toy, key):
    skill, member, splits, _ = toy(key)
    g = golden(key)
    co = staging.cooccurrence(member, skill, skipteams=splits['test'], device=DEV)
    assert str(co.dtype) == str(g['co/dtype']) and same_csr(co, g['co/indptr'], g['co/indices'], g['co/data'])
    assert same_csr(staging.cooccurrence(member, skill, device=DEV), g['co_all/indptr'], g['co_all/indices'], g['co_all/data'])


@pytest.mark.parametrize('T,E,S,zipf', [(1, 3, 4, False), (400, 60, 33, False), (5000, 2000, 1500, True), (20000, 44774, 30000, True), (3000, 500, 213317, False)])
def test_cooccurrence_matches_the_oracle_incl_skipped_teams_hot_pairs_and_idle_experts(staging, T, E, S, zipf):
    rng = np.random.default_rng(T + E)
    member, skill = rand_multihot(rng, T, E, 0, 6, zipf), rand_multihot(rng, T, S, 1, 12, zipf)
    skip = rng.choice(T, size=T // 7, replace=False) if T > 7 else None
    want = SO.cooccurrence(member, skill, skip)
    got = staging.cooccurrence(member, skill, skip, DEV)
    assert got.dtype == want.dtype and got.has_sorted_indices
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices) and np.array_equal(got.data, want.data)


def test_cooccurrence_counts_wrap_like_the_reference_uint8_product(staging):
    T = 300
    m, s = sp.lil_matrix((T, 2), dtype='u1'), sp.lil_matrix((T, 3), dtype='u1')
    for t in range(256): m[t, 0] = 1; s[t, 1] = 1
    for t in range(300): m[t, 1] = 1; s[t, 2] = 1
    co = staging.cooccurrence(m, s, device=DEV)
    assert co.dtype == np.uint8 and co.nnz == 1 and co[1, 2] == 44
    # the library itself counts in int32
    from opentf_b200 import ops
    M, S_ = sp.csr_matrix(m), sp.csr_matrix(s)
    i32 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32), device=DEV)
    ptr, idx, val = ops.cooccur(i32(M.indptr), i32(M.indices), i32(S_.indptr), i32(S_.indices), 2, 3, None, ops.Workspace(torch.device(DEV)))
    assert ptr.tolist() == [0, 2, 4] and idx.tolist() == [1, 2, 1, 2] and val.tolist() == [256, 256, 256, 300]  # (teams 0..255 hold both experts and both skills)


def test_cooccurrence_row_sums_at_the_dblp_shape(staging):
    """size-independent property at BASELINE configs[1]'s shape (no oracle run: the scipy product of 100 000 teams takes a while): every row of
    member^T . skill sums to the total number of skills of the expert's teams, and the pattern is symmetric with skill^T . member"""
    from opentf_b200 import ops
    rng = np.random.default_rng(0)
    T, E, S = 100000, 40000, 30000
    member, skill = rand_multihot(rng, T, E, 1, 5, True), rand_multihot(rng, T, S, 1, 10, True)
    i32 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32), device=DEV)
    ws = ops.Workspace(torch.device(DEV))
    ptr, idx, val = ops.cooccur(i32(member.indptr), i32(member.indices), i32(skill.indptr), i32(skill.indices), E, S, None, ws)
    co = sp.csr_matrix((val.cpu().numpy().astype(np.int64), idx.cpu().numpy(), ptr.cpu().numpy()), shape=(E, S))
    n_s = np.diff(skill.indptr).astype(np.int64)
    want_rows = np.asarray(member.astype(np.int64).T @ n_s).ravel()
    assert np.array_equal(np.asarray(co.sum(axis=1)).ravel(), want_rows) and co.has_sorted_indices and (co.data > 0).all()
    ptr2, idx2, val2 = ops.cooccur(i32(skill.indptr), i32(skill.indices), i32(member.indptr), i32(member.indices), S, E, None, ws)
    co_t = sp.csr_matrix((val2.cpu().numpy().astype(np.int64), idx2.cpu().numpy(), ptr2.cpu().numpy()), shape=(S, E))
    assert (co_t.T.tocsr() != co).nnz == 0


@pytest.mark.parametrize('key', KEYS)
def test_skill_coverage_is_the_reference_frame(staging, toy, key):
    skill, member, splits, z = toy(key)
    g = golden(key)
    co = sp.csr_matrix((g['co/data'].astype(np.uint8), g['co/indices'], g['co/indptr']), shape=(member.shape[1], skill.shape[1]))
    X = skill[splits['test']]
    df, df_mean = staging.calculate_skill_coverage(X, g['random/Y_'], co, True, '2,5,10', DEV)
    for k in (2, 5, 10): assert np.array_equal(df[f'skill_coverage_{k}'].to_numpy(), g[f'random/skill_coverage_{k}']), (key, k)
    assert list(df_mean.index) == ['skill_coverage_2', 'skill_coverage_5', 'skill_coverage_10'] and df_mean.index.name == 'metrics'
    if 'pred/skill_coverage_2' in g.files:
        df, _ = staging.calculate_skill_coverage(X, z['pred/f0'], co, True, '2,5,10', DEV)
        for k in (2, 5, 10): assert np.array_equal(df[f'skill_coverage_{k}'].to_numpy(), g[f'pred/skill_coverage_{k}']), (key, k)


@pytest.mark.parametrize('sparse_pred', [False, True])
def test_skill_coverage_matches_the_oracle_on_random_teams(staging, sparse_pred):
    rng = np.random.default_rng(17)
    T, E, S, n = 3000, 800, 300, 257
    member, skill = rand_multihot(rng, T, E, 1, 5, True), rand_multihot(rng, T, S, 1, 40, True)
    co = SO.cooccurrence(member, skill, np.arange(n))
    X = skill[:n]
    Y_ = (rng.permutation(n * E).reshape(n, E).astype(np.float32) + 1) / (n * E + 1)  # distinct positive scores
    if sparse_pred:  # a top-20 .pred file (fnn.py:218): the cut-offs stay inside the stored entries
        keep = np.argsort(-Y_, axis=1)[:, :20]
        Yd = np.zeros_like(Y_); np.put_along_axis(Yd, keep, np.take_along_axis(Y_, keep, axis=1), axis=1)
        Y_ = sp.csr_matrix(Yd)
    want = SO.skill_coverage(X, Y_, co, '1,2,5,10,20')
    df, _ = staging.calculate_skill_coverage(X, Y_, co, True, '1,2,5,10,20', DEV)
    for k in (1, 2, 5, 10, 20): assert np.array_equal(df[f'skill_coverage_{k}'].to_numpy(), want[k]), k
    assert df['skill_coverage_20'].mean() > df['skill_coverage_1'].mean()  # (the union only grows with k)
