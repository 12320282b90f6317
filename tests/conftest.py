import os, sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path: sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run by the driver with -m gpu)')


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) on a box without a GPU even if someone forgets -m "not gpu"."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu: return
    skip = pytest.mark.skip(reason='no CUDA device')
    for it in items:
        if 'gpu' in it.keywords: it.add_marker(skip)


def load_toy(key):
    """-> (skill csr, member csr, splits dict, raw npz) from tests/golden/toy_{key}.npz"""
    import scipy.sparse as sp
    z = np.load(os.path.join(GOLDEN, f'toy_{key}.npz'))
    def mat(name):
        ind = z[f'{name}_indices']
        return sp.csr_matrix((np.ones(len(ind), dtype=np.uint8), ind, z[f'{name}_indptr']), shape=tuple(z[f'{name}_shape']))
    splits = {'test': z['test'], 'folds': {k: {'train': z[f'fold{k}_train'], 'valid': z[f'fold{k}_valid']} for k in range(3)}}
    return mat('skill'), mat('member'), splits, z


@pytest.fixture(scope='session')
def toy():
    cache = {}
    def get(key):
        if key not in cache: cache[key] = load_toy(key)
        return cache[key]
    return get
