"""Import the UNMODIFIED reference hot path (`/root/reference/src/mdl/fnn.py`) in this container.

TEST INFRASTRUCTURE ONLY.  This module exists to (a) pin the oracle restatement
(`oracle/fnn_oracle.py`) against the real reference and (b) generate the golden
fixtures under `tests/golden/` (see `tests/golden/make_golden.py`).  It is never
imported by the product (`opentf_b200/`).  On the GPU box `/root/reference` does not exist:
there the same unmodified files are found under `oracle/_ref/src` (placed by `oracle/make_ref.py`
at build time, git-ignored) and serve as the CPU baseline of `bench.py`.

Five environment shims, none of which touches arithmetic:
  1. a stub `pkgmgr` module (the real one imports omegaconf and opens
     ../requirements.txt at import time: src/pkgmgr.py:3,79);
  2. `np.Inf` (removed in numpy 2; src/mdl/earlystopping.py:21);
  3. `ReduceLROnPlateau(verbose=...)` (removed in torch 2.11; src/mdl/fnn.py:105);
  4. `torch.load(weights_only=False)` (src/mdl/fnn.py:187, src/mdl/ntf.py:52);
  5. a plain-object cfg instead of an OmegaConf node.
"""
import importlib, os, random, sys, types

_LOCAL = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'src')
REF_SRC = '/root/reference/src' if os.path.isdir('/root/reference/src/mdl') else _LOCAL


def available():
    return os.path.isfile(os.path.join(REF_SRC, 'mdl', 'fnn.py'))


class Cfg(dict):
    """attribute-style dict standing in for the OmegaConf node (shim 5)."""
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


def _install_shims():
    import numpy as np, torch
    if 'pkgmgr' in sys.modules and getattr(sys.modules['pkgmgr'], '_ntf_b200_shim', False):
        return
    if not hasattr(np, 'Inf'):
        np.Inf = np.inf  # shim 2

    _orig_sched = torch.optim.lr_scheduler.ReduceLROnPlateau
    if not getattr(_orig_sched, '_ntf_b200_shim', False):
        class _Sched(_orig_sched):  # shim 3
            _ntf_b200_shim = True

            def __init__(self, *a, verbose=None, **k):
                super().__init__(*a, **k)
        torch.optim.lr_scheduler.ReduceLROnPlateau = _Sched

    _orig_load = torch.load
    if not getattr(_orig_load, '_ntf_b200_shim', False):
        def _load(*a, **k):  # shim 4
            k.setdefault('weights_only', False)
            return _orig_load(*a, **k)
        _load._ntf_b200_shim = True
        torch.load = _load

    m = types.ModuleType('pkgmgr')  # shim 1
    m._ntf_b200_shim = True

    class _NullWriter:
        def __init__(self, *a, **k): pass
        def add_scalar(self, *a, **k): pass
        def close(self): pass

    def install_import(pkg_name, import_path=None, from_module=None):
        if pkg_name == 'tensorboardX':
            return _NullWriter
        mod = importlib.import_module(import_path or pkg_name)
        return getattr(mod, from_module) if from_module else mod

    def set_seed(seed, torch=None):  # mirrors src/pkgmgr.py:81-93 on a CPU-only box
        if seed is None: return
        random.seed(seed)
        np.random.seed(seed)
        if torch: torch.manual_seed(seed)

    def cfg2str(cfg):
        return '.'.join(f'{k}{v}' for k, v in cfg.items()) if cfg else ''

    def topk_sparse(torch, probs, k, type='coo'):
        # the reference's helper is 6 lines of torch calls (src/pkgmgr.py:125-134); the stub has to
        # provide the same result because fnn.py:218 calls it through this module.
        v, i = torch.topk(probs, k, dim=1)
        r = torch.arange(probs.shape[0]).unsqueeze(1).expand(-1, k)
        return torch.sparse_coo_tensor(torch.stack([r, i], 0).reshape(2, -1), v.reshape(-1), size=probs.shape).coalesce()

    m.install_import, m.set_seed, m.cfg2str, m.topk_sparse = install_import, set_seed, cfg2str, topk_sparse
    m.textcolor = {k: '' for k in ('blue', 'green', 'yellow', 'red', 'magenta', 'cyan', 'reset')}
    sys.modules['pkgmgr'] = m


def load_reference_fnn():
    """-> the reference's `mdl.fnn.Fnn` class, unmodified."""
    if not available():
        raise RuntimeError('reference tree not mounted; the golden fixtures under tests/golden are the travelling copy')
    _install_shims()
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    from mdl.fnn import Fnn
    return Fnn


def default_cfg(**over):
    """the `fnn:` block of src/mdl/__config__.yaml:18-31 with the root defaults resolved."""
    c = Cfg(b=1000, e=100, ns=5, lr=0.001, es=5, h=[128], spe=10, l='bce', tpw=10, tnw=1, nsd='unigram_b')
    c.update(over)
    return c
