"""Synthetic teamsvecs of the shapes the reference's datasets have, for benchmarks and scale tests.

Rows: n_skill ~ 1 + Poisson(lam_s - 1), n_member ~ m_min + Poisson(lam_m - m_min); ids drawn without replacement per
team from a Zipf(alpha) popularity over a fixed random permutation of the columns (real data is heavy-tailed);
every column appears at least once (mirrors Team.validate, team.py:185-201).  Output: scipy csr uint8 (the
reference hands lil uint8, team.py:154; both are accepted at the boundary) and splits made exactly like
main.py:24-33 (train_test_split + KFold with the run seed).
"""
import numpy as np
import scipy.sparse as sp

SHAPES = {  # name: (N teams, S skills, E experts, lam_s, lam_m, m_min)   -- SURVEY.md 8d C2..C4
    'dblp': (100_000, 30_000, 40_000, 8.57, 3.06, 2),
    'imdb': (186_385, 27, 44_774, 1.54, 2.6, 2),
    'uspt': (262_144, 213_317, 394_187, 6.29, 2.51, 2),
    'toy': (512, 64, 96, 3.0, 2.5, 2),
}


def _draw_rows(rng, k, ncols, alpha):
    """k[i] distinct column ids per row i, successive draws from Zipf(alpha) with repeats skipped."""
    N = len(k)
    k = np.minimum(k, ncols)
    w = 1.0 / np.arange(1, ncols + 1) ** alpha
    cum = np.cumsum(w)
    perm = rng.permutation(ncols)
    have = np.zeros(0, dtype=np.int64)
    need = k.astype(np.int64).copy()
    while need.sum() > 0:
        rows = np.repeat(np.arange(N, dtype=np.int64), need)
        cols = perm[np.minimum(np.searchsorted(cum, rng.random(len(rows)) * cum[-1]), ncols - 1)]
        keys = np.unique(rows * ncols + cols)
        if len(have): keys = keys[~np.isin(keys, have, assume_unique=True)]
        have = np.union1d(have, keys)
        need = k - np.bincount(have // ncols, minlength=N)
    rows, cols = have // ncols, have % ncols
    missing = np.setdiff1d(np.arange(ncols), cols, assume_unique=False)
    if len(missing):  # force every column to appear once
        rows = np.concatenate([rows, rng.integers(0, N, len(missing))])
        cols = np.concatenate([cols, missing])
    m = sp.csr_matrix((np.ones(len(rows), dtype=np.uint8), (rows, cols)), shape=(N, ncols))
    m.sum_duplicates()
    m.data[:] = 1
    m.sort_indices()
    return m


def make_teamsvecs(name='dblp', seed=0, n_teams=None, n_skills=None, n_experts=None, alpha=1.1):
    N, S, E, lam_s, lam_m, m_min = SHAPES[name]
    N, S, E = n_teams or N, n_skills or S, n_experts or E
    rng = np.random.default_rng(seed)
    ks = 1 + rng.poisson(max(lam_s - 1, 0.0), N)
    km = m_min + rng.poisson(max(lam_m - m_min, 0.0), N)
    return {'skill': _draw_rows(rng, ks, S, alpha), 'member': _draw_rows(rng, km, E, alpha)}


def make_splits(n_sample, n_folds=3, train_ratio=0.85, seed=0):
    """main.py:24-33."""
    from sklearn.model_selection import KFold, train_test_split
    train, test = train_test_split(np.arange(n_sample), train_size=train_ratio, random_state=seed, shuffle=True)
    splits = {'test': test, 'folds': {}}
    for k, (tr, va) in enumerate(KFold(n_splits=n_folds, random_state=seed, shuffle=True).split(train)):
        splits['folds'][k] = {'train': train[tr], 'valid': train[va]}
    return splits
