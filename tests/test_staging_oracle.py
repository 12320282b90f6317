"""CPU tests of SURVEY 8(f-4): the staging oracle (oracle/staging_oracle.py) against fixtures recorded from the UNMODIFIED reference
(tests/golden/staging_*.npz: Team.gen_skill_coverage, team.py:302-341, and calculate_skill_coverage, metric.py:44-73, on the committed toy
teamsvecs) and against the committed teamsvecs themselves (the output of Team.bucketing, team.py:148-173)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN
from oracle import staging_oracle as SO

KEYS = ['dblp', 'imdb', 'gith', 'uspt']


def golden(key):
    return np.load(os.path.join(GOLDEN, f'staging_{key}.npz'))


def same_csr(a, indptr, indices, data):
    a = sp.csr_matrix(a); a.sort_indices()
    return np.array_equal(a.indptr, indptr) and np.array_equal(a.indices, indices) and np.array_equal(a.data.astype(np.int64), data)


def messy_lists(M, rng):
    """every row of the multi-hot matrix as an id list in random order with random repeats"""
    M = sp.csr_matrix(M)
    indptr, ids = [0], []
    for i in range(M.shape[0]):
        row = M.indices[M.indptr[i]:M.indptr[i + 1]]
        row = np.concatenate([row, rng.choice(row, size=rng.integers(0, 4))]) if len(row) else row
        ids.append(rng.permutation(row)); indptr.append(indptr[-1] + len(row))
    return np.asarray(indptr), (np.concatenate(ids) if ids else np.zeros(0, np.int64))


@pytest.mark.parametrize('key', KEYS)
def test_oracle_cooccurrence_is_the_reference_matrix(toy, key):
    skill, member, splits, _ = toy(key)
    g = golden(key)
    co = SO.cooccurrence(member, skill, skipteams=splits['test'])
    assert str(co.dtype) == str(g['co/dtype']) and same_csr(co, g['co/indptr'], g['co/indices'], g['co/data'])
    assert same_csr(SO.cooccurrence(member, skill), g['co_all/indptr'], g['co_all/indices'], g['co_all/data'])


@pytest.mark.parametrize('key', KEYS)
def test_oracle_skill_coverage_is_the_reference_frame(toy, key):
    skill, member, splits, z = toy(key)
    g = golden(key)
    co = sp.csr_matrix((g['co/data'].astype(np.uint8), g['co/indices'], g['co/indptr']), shape=(member.shape[1], skill.shape[1]))
    X = skill[splits['test']]
    cov = SO.skill_coverage(X, g['random/Y_'], co, '2,5,10')
    for k in (2, 5, 10): assert np.array_equal(cov[k], g[f'random/skill_coverage_{k}']), (key, k)
    if 'pred/skill_coverage_2' in g.files:  # the reference's own committed prediction file of the toy run
        cov = SO.skill_coverage(X, z['pred/f0'], co, '2,5,10')
        for k in (2, 5, 10): assert np.array_equal(cov[k], g[f'pred/skill_coverage_{k}']), (key, k)


@pytest.mark.parametrize('key', KEYS)
def test_oracle_one_hot_rows_rebuild_the_committed_teamsvecs(toy, key):
    skill, member, _, _ = toy(key)
    rng = np.random.default_rng(3)
    for M in (skill, member):
        indptr, ids = messy_lists(M, rng)
        R = SO.rows_from_lists(indptr, ids, M.shape[1])
        assert (R != sp.csr_matrix(M)).nnz == 0 and R.dtype == np.uint8


def test_oracle_cooccurrence_wraps_like_the_reference_uint8_product():
    """256 co-occurrences store nothing, 300 store 44 (checked against Team.gen_skill_coverage when the golden file was made)"""
    T = 300
    m, s = sp.lil_matrix((T, 2), dtype='u1'), sp.lil_matrix((T, 3), dtype='u1')
    for t in range(256): m[t, 0] = 1; s[t, 1] = 1
    for t in range(300): m[t, 1] = 1; s[t, 2] = 1
    co = SO.cooccurrence(m, s)
    assert co.dtype == np.uint8 and co.nnz == 1 and co[1, 2] == 44


def test_product_staging_has_no_host_path():
    """opentf_b200.staging is the device implementation only: on a CPU device (or without a GPU) every entry point raises instead of computing"""
    from opentf_b200 import staging
    m, s = sp.csr_matrix(np.eye(3, dtype=np.uint8)), sp.csr_matrix(np.eye(3, dtype=np.uint8))
    for call in (lambda: staging.csr_from_lists([0, 1], [0], 3, device='cpu'), lambda: staging.cooccurrence(m, s, device='cpu'),
                 lambda: staging.calculate_skill_coverage(s, np.eye(3, dtype=np.float32), m, topks='1', device='cpu')):
        with pytest.raises(RuntimeError, match='CUDA'):
            call()
