"""Host helpers that mirror the reference's `pkgmgr` module for this path (src/pkgmgr.py:81-95,106-114,125-143)."""
import importlib
import random

import numpy as np

textcolor = {'blue': '\033[94m', 'green': '\033[92m', 'yellow': '\033[93m', 'red': '\033[91m', 'magenta': '\033[95m', 'cyan': '\033[96m',
             'reset': '\033[0m'}


def set_seed(seed, torch=None):
    """pkgmgr.py:81-93."""
    if seed is None: return
    random.seed(seed)
    np.random.seed(seed)
    if torch:
        torch.manual_seed(seed)
        if torch.cuda.is_available():
            torch.cuda.manual_seed(seed)
            torch.cuda.manual_seed_all(seed)
            torch.backends.cudnn.deterministic = True
            torch.backends.cudnn.benchmark = False


def cfg_items(cfg):
    """(key, value) pairs of an OmegaConf node, dict, or attribute object, in declaration order."""
    if cfg is None: return []
    try:
        from omegaconf import OmegaConf  # present when hosted inside OpeNTF
        if OmegaConf.is_config(cfg): return list(OmegaConf.to_container(cfg, resolve=True).items())
    except ImportError:
        pass
    if isinstance(cfg, dict): return list(cfg.items())
    return list(vars(cfg).items())


def cfg2str(cfg):
    """pkgmgr.py:95: 'b1000.e100.ns5...' -- the run directory name is part of the contract (ntf.py:27)."""
    return '.'.join(f'{k}{v}' for k, v in cfg_items(cfg))


def cfg_get(cfg, key, default=None):
    if cfg is None: return default
    if isinstance(cfg, dict): return cfg.get(key, default)
    try:
        v = getattr(cfg, key)
    except (AttributeError, KeyError):
        return default
    return default if v is None and default is not None else v


def summary_writer():
    """tensorboardX.SummaryWriter when installed (ntf.py:13), else torch's, else a no-op with the same calls."""
    for mod in ('tensorboardX', 'torch.utils.tensorboard'):
        try:
            return getattr(importlib.import_module(mod), 'SummaryWriter')
        except Exception:
            continue

    class _Null:
        def __init__(self, *a, **k): pass
        def add_scalar(self, *a, **k): pass
        def close(self): pass
    return _Null


def topk_to_sparse(torch, vals, idx, n_cols):
    """[N,K] (values, expert ids) -> the coalesced sparse COO tensor `pkgmgr.topk_sparse` stores in *.pred files
    (pkgmgr.py:125-134: sorted by row, then column)."""
    N, K = idx.shape
    rows = torch.arange(N).unsqueeze(1).expand(-1, K)
    ind = torch.stack([rows, idx.to(torch.int64)], dim=0).reshape(2, -1)
    return torch.sparse_coo_tensor(ind, vals.reshape(-1), size=(N, n_cols)).coalesce()


def torch_sparse_2_scipy_sparse(Y_, type='coo'):
    """pkgmgr.py:136-143."""
    import scipy.sparse
    Y_ = Y_.coalesce()
    i, v = Y_.indices().cpu().numpy(), Y_.values().cpu().numpy()
    cls = scipy.sparse.coo_matrix if type == 'coo' else scipy.sparse.csr_matrix
    return cls((v, (i[0], i[1])), shape=tuple(Y_.size()))


def first_device(device):
    """'cuda' / 'cuda:N' / the reference's 'cuda:0,1,...' list (nmt.py:74) -> this process's 'cuda:N'."""
    import os
    d = str(device)
    if ',' in d:
        ids = d.split(':')[1].split(',')
        return f'cuda:{ids[int(os.environ.get("LOCAL_RANK", 0)) % len(ids)]}'
    return d if ':' in d else 'cuda:0'
