"""Generate the golden fixtures in this directory from the reference tree (run in the build container,
where `/root/reference` is mounted; the fixtures are what travels to the GPU box).

    python tests/golden/make_golden.py

Writes (all numpy .npz, no pickled objects):
  toy_{dblp,imdb,gith,uspt}.npz   teamsvecs as CSR arrays + the committed splits + every committed
                                  (checkpoint, prediction) pair of the reference's fnn run  (G1), the
                                  committed per-instance eval CSV of f{k}.test.pred (G3) and, for the
                                  bnn run, parameter names/shapes/statistics (G4)
  traj_dblp_{unigram_b,uniform,unigram}.npz
                                  a seed-0 `Fnn.learn` trajectory recorded from the UNMODIFIED reference
                                  (imported through oracle/ref_shim.py): initial weights, per-step
                                  (fold, epoch, phase, row ids, sampled negatives, loss), final weights and
                                  epoch losses (G2).  For unigram_b the final numbers are also asserted
                                  against the checkpoints committed in the reference (made with torch 2.4.1).
"""
import csv, glob, os, pickle, re, sys, tempfile, types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_shim  # noqa: E402

REF_OUT = '/root/reference/output'
TOYS = {'dblp': 'dblp/toy.dblp.v12.json', 'imdb': 'imdb/toy.title.basics.tsv', 'gith': 'gith/toy.repos.csv', 'uspt': 'uspt/toy.patent.tsv'}
RUN = 'splits.f3.r0.85/fnn.b1000.e100.ns5.lr0.001.es5.h[128].spe10.lbce.tpw10.tnw1.nsdunigram_b'
BRUN = 'splits.f3.r0.85/bnn.b1000.e100.ns5.lr0.001.es5.h[128].spe10.lbce.tpw10.tnw1.nsdunigram_b.nmc10'


def _stub_omegaconf():
    """the committed .pt files pickle their `cfg` as OmegaConf nodes; omegaconf is not installed here."""
    if 'omegaconf' in sys.modules: return
    class _Any:
        def __init__(self, *a, **k): pass
        def __setstate__(self, s): self.__dict__['_s'] = s
    for name, classes in {'omegaconf': [], 'omegaconf.base': ['Metadata', 'ContainerMetadata'], 'omegaconf.listconfig': ['ListConfig'],
                          'omegaconf.dictconfig': ['DictConfig'], 'omegaconf.nodes': ['AnyNode']}.items():
        m = types.ModuleType(name)
        for c in classes: setattr(m, c, type(c, (_Any,), {}))
        sys.modules[name] = m


def csr_arrays(mat):
    c = mat.tocsr(); c.sort_indices()
    return c.indptr.astype(np.int64), c.indices.astype(np.int32), np.array(c.shape, dtype=np.int64)


def load_toy(key):
    root = f'{REF_OUT}/{TOYS[key]}'
    tv = pickle.load(open(f'{root}/teamsvecs.pkl', 'rb'))
    sp_ = pickle.load(open(f'{root}/splits.f3.r0.85.pkl', 'rb'))
    return root, tv, sp_


def dump_toy(key):
    _stub_omegaconf()
    root, tv, sp_ = load_toy(key)
    out = {}
    for name in ('skill', 'member'):
        out[f'{name}_indptr'], out[f'{name}_indices'], out[f'{name}_shape'] = csr_arrays(tv[name])
    out['test'] = np.asarray(sp_['test'], dtype=np.int64)
    for k, f in sp_['folds'].items():
        out[f'fold{k}_train'] = np.asarray(f['train'], dtype=np.int64)
        out[f'fold{k}_valid'] = np.asarray(f['valid'], dtype=np.int64)
    rundir = f'{root}/{RUN}'
    pairs = []
    for ck in sorted(os.listdir(rundir)):
        m = re.fullmatch(r'f(\d+)(\.e\d+)?\.pt', ck)
        if not m: continue
        pred = f'{rundir}/f{m.group(1)}.test{m.group(2) or ""}.pred'
        if not os.path.exists(pred): continue
        tag = ck[:-3]
        d = torch.load(f'{rundir}/{ck}', map_location='cpu', weights_only=False)
        for name, t in d['model_state_dict'].items(): out[f'ckpt/{tag}/{name}'] = t.numpy()
        out[f'ckpt/{tag}/e'] = np.int64(d['e'])
        out[f'ckpt/{tag}/t_loss'] = np.float64(d['t_loss']); out[f'ckpt/{tag}/v_loss'] = np.float64(d['v_loss'])
        p = torch.load(pred, map_location='cpu', weights_only=False)['y_pred']
        out[f'pred/{tag}'] = (p.to_dense() if p.is_sparse else p).numpy()
        pairs.append(tag)
    out['pairs'] = np.array(pairs)
    # G3: per-instance metrics of the final fold models + the fold-mean file
    for k in sp_['folds'].keys():
        rows = list(csv.reader(open(f'{rundir}/f{k}.test.pred.eval.instance.csv')))
        if rows[0][0] == '':  # older layout committed for some toys: metrics as rows, queries q0.. as columns
            out[f'eval/f{k}/columns'] = np.array([r[0] for r in rows[1:]]); out[f'eval/f{k}/values'] = np.array([r[1:] for r in rows[1:]], dtype=np.float64).T
        else:
            out[f'eval/f{k}/columns'] = np.array(rows[0]); out[f'eval/f{k}/values'] = np.array(rows[1:], dtype=np.float64)
        rows = list(csv.reader(open(f'{rundir}/f{k}.test.pred.eval.mean.csv')))
        out[f'eval/f{k}/mean_names'] = np.array([r[0] for r in rows[1:]]); out[f'eval/f{k}/mean_values'] = np.array([float(r[1]) for r in rows[1:]])
    # G4: Bnn layout + statistics + the committed MC prediction (distribution-level pin only)
    bdir = f'{root}/{BRUN}'
    if os.path.isdir(bdir):
        d = torch.load(f'{bdir}/f0.pt', map_location='cpu', weights_only=False)
        names = list(d['model_state_dict'].keys())
        out['bnn/names'] = np.array(names)
        out['bnn/shapes'] = np.array([list(d['model_state_dict'][n].shape) + [0] * (2 - d['model_state_dict'][n].dim()) for n in names])
        out['bnn/mean_std'] = np.array([[float(d['model_state_dict'][n].mean()), float(d['model_state_dict'][n].std())] for n in names])
        for n in names: out[f'bnn/f0/{n}'] = d['model_state_dict'][n].numpy()
        out['bnn/f0/t_loss'] = np.float64(d['t_loss']); out['bnn/f0/v_loss'] = np.float64(d['v_loss']); out['bnn/f0/e'] = np.int64(d['e'])
        p = torch.load(f'{bdir}/f0.test.pred', map_location='cpu', weights_only=False)
        yp = p['y_pred']; out['bnn/f0/y_pred'] = (yp.to_dense() if yp.is_sparse else yp).numpy()
        out['bnn/f0/unc_pred'] = np.asarray(p['uncertainty']['pred'][0]); out['bnn/f0/unc_model'] = np.asarray(p['uncertainty']['model'][0])
    np.savez_compressed(f'{HERE}/toy_{key}.npz', **out)
    print(key, 'pairs:', len(pairs), 'bytes:', os.path.getsize(f'{HERE}/toy_{key}.npz'))


def dense_skill_vectors(n, d, seed=7):
    """stand-in for the d2v/gnn skill embeddings main.py:148-153 puts in teamsvecs['skill'] (one fp32 d-vector per team)"""
    return np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)


def record_trajectory(nsd, key='dblp', seed=0, tag=None, dense_d=0, **cfg_over):
    """run the verbatim reference and tap (without editing it) the dataset index order, the sampler
    output and the per-step loss.  dense_d > 0: teamsvecs['skill'] is replaced by dense vectors (the ntf.py:24 branch)."""
    Fnn = ref_shim.load_reference_fnn()
    root, tv, sp_ = load_toy(key)
    if dense_d: tv = dict(tv, skill=dense_skill_vectors(tv['skill'].shape[0], dense_d))
    outdir = tempfile.mkdtemp()
    cfg = ref_shim.default_cfg(nsd=nsd, **cfg_over)
    mdl = Fnn(outdir, 'cpu', seed, cfg)
    steps, fetched, inits, state = [], [], [], {'neg': None}

    DS = type(mdl).dataset  # Ntf.dataset (ntf.py:25)
    orig_getitem = DS.__getitem__
    def tapped_getitem(self, index):
        fetched.append(int(index)); return orig_getitem(self, index)
    DS.__getitem__ = tapped_getitem

    for name in ('ns_uniform', 'ns_unigram'):
        orig = getattr(mdl, name)
        def tapped(y, _orig=orig):
            r = _orig(y); state['neg'] = r.clone(); return r
        setattr(mdl, name, tapped)

    orig_init = mdl.init
    def tapped_init(*a, **k):
        m = orig_init(*a, **k)
        inits.append({n: t.detach().clone().numpy() for n, t in m.state_dict().items()})
        return m
    mdl.init = tapped_init

    orig_bxe = mdl.bxe
    def tapped_bxe(y_, y):
        state['neg'] = None
        r = orig_bxe(y_, y)
        steps.append(dict(idx=list(fetched), neg=None if state['neg'] is None else state['neg'].numpy().copy(),
                          loss=float(r.sum(dim=1).mean().item()), train=torch.is_grad_enabled()))
        fetched.clear()
        return r
    mdl.bxe = tapped_bxe

    mdl.learn(tv, sp_, None)
    DS.__getitem__ = orig_getitem

    out = {'seed': np.int64(seed), 'nsd': np.array(nsd), 'dense_d': np.int64(dense_d), 'cfg_keys': np.array(list(cfg.keys())), 'cfg_vals': np.array([str(v) for v in cfg.values()])}
    if dense_d: out['dense_x'] = np.asarray(tv['skill'], dtype=np.float32)  # the inputs travel with the recording
    # the DataLoader prefetches nothing with num_workers=0, so `fetched` at bxe time is exactly this batch's
    # positions WITHIN the fold's train/valid subset; translate to team row ids.
    fold_ids = list(sp_['folds'].keys())
    si = 0
    for fi, k in enumerate(fold_ids):
        tr, va = np.asarray(sp_['folds'][k]['train']), np.asarray(sp_['folds'][k]['valid'])
        ck = torch.load(f'{mdl.output}/f{k}.pt', map_location='cpu', weights_only=False)
        nb_t, nb_v = -(-len(tr) // cfg.b), -(-len(va) // cfg.b)
        nsteps = (int(ck['e']) + 1) * (nb_t + nb_v)
        rec = steps[si:si + nsteps]; si += nsteps
        ph = np.array([0 if s['train'] else 1 for s in rec], dtype=np.int8)
        out[f'f{k}/phase'] = ph
        out[f'f{k}/loss'] = np.array([s['loss'] for s in rec], dtype=np.float64)
        rows = [(tr if s['train'] else va)[np.asarray(s['idx'], dtype=np.int64)] for s in rec]
        out[f'f{k}/rows_ptr'] = np.cumsum([0] + [len(r) for r in rows]).astype(np.int64)
        out[f'f{k}/rows'] = np.concatenate(rows).astype(np.int64)
        if rec and rec[0]['neg'] is not None: out[f'f{k}/neg'] = np.concatenate([s['neg'] for s in rec], axis=0).astype(np.int64)
        for n, a in inits[fi].items(): out[f'f{k}/init/{n}'] = a
        for n, t in ck['model_state_dict'].items(): out[f'f{k}/final/{n}'] = t.numpy()
        out[f'f{k}/e'] = np.int64(ck['e']); out[f'f{k}/t_loss'] = np.float64(ck['t_loss']); out[f'f{k}/v_loss'] = np.float64(ck['v_loss'])
        out[f'f{k}/ckpt_epochs'] = np.array(sorted(int(re.search(r'\.e(\d+)\.pt', f).group(1)) for f in os.listdir(mdl.output) if re.fullmatch(rf'f{k}\.e\d+\.pt', f)), dtype=np.int64)
    assert si == len(steps), (si, len(steps))
    tag = tag or nsd
    np.savez_compressed(f'{HERE}/traj_{key}_{tag}.npz', **out)
    print('traj', key, tag, {k: (int(out[f'f{k}/e']), float(out[f'f{k}/t_loss']), float(out[f'f{k}/v_loss'])) for k in fold_ids})
    return out


def record_tntf(key='gith', seed=0):
    """the reference's temporal wrapper (tntf.py) around its Fnn, unmodified: year-by-year fine-tuning with warm starts.  nsd is unset,
    so bxe draws nothing (fnn.py:39) and the run is a function of the seed alone -- the CUDA path is compared free-running."""
    Fnn = ref_shim.load_reference_fnn()
    from mdl.tntf import tNtf
    root, tv, sp_ = load_toy(key)
    n = tv['skill'].shape[0]
    year_idx = [(0, 2000), (n // 3, 2001), (2 * n // 3, 2002), (n - 4, 2003)]  # 3 training intervals, the last year is for test
    cfg = ref_shim.default_cfg(nsd=None, b=8, e=5, h=[16], lr=0.01, es=5, spe=0)
    tcfg = ref_shim.Cfg(tfolds=2, step_ahead=1)
    outdir = tempfile.mkdtemp()
    inner = Fnn(outdir, 'cpu', seed, cfg)
    w = tNtf(outdir, 'cpu', seed, tcfg, inner, year_idx)
    splits = {'test': np.arange(year_idx[-1][0], n), 'folds': {k: {'train': None, 'valid': None} for k in range(2)}}
    w.learn(tv, splits, None)
    out = {'year_idx': np.array(year_idx, dtype=np.int64), 'seed': np.int64(seed)}
    for _, year in year_idx[:-1]:
        sp_y = pickle.load(open(f'{w.output}/{year}/splits.pkl', 'rb'))
        for k in range(2):
            ck = torch.load(f'{w.output}/{year}/f{k}.pt', map_location='cpu', weights_only=False)
            for nm, t in ck['model_state_dict'].items(): out[f'{year}/f{k}/{nm}'] = t.numpy()
            out[f'{year}/f{k}/e'] = np.int64(ck['e']); out[f'{year}/f{k}/t_loss'] = np.float64(ck['t_loss']); out[f'{year}/f{k}/v_loss'] = np.float64(ck['v_loss'])
            out[f'{year}/f{k}/train'] = np.asarray(sp_y['folds'][k]['train'], dtype=np.int64)
            out[f'{year}/f{k}/valid'] = np.asarray(sp_y['folds'][k]['valid'], dtype=np.int64)
    np.savez_compressed(f'{HERE}/tntf_{key}.npz', **out)
    print('tntf', key, {y: (int(out[f'{y}/f0/e']), float(out[f'{y}/f0/t_loss'])) for _, y in year_idx[:-1]})


def record_staging(key):
    """the UNMODIFIED Team.gen_skill_coverage (team.py:302-341, test teams skipped as main.py:98 does) and calculate_skill_coverage
    (metric.py:44-73) on the committed toy teamsvecs; the predictions are seeded random scores (distinct per row: the ranking has no ties) and the
    committed f0.test.pred of the reference's own fnn run where it exists"""
    ref_shim._install_shims()
    if ref_shim.REF_SRC not in sys.path: sys.path.insert(0, ref_shim.REF_SRC)
    from cmn.team import Team
    from evl import metric as refmetric
    root, tv, sp_ = load_toy(key)
    out = {}
    co = Team.gen_skill_coverage(tv, tempfile.mkdtemp(), skipteams=sp_['test'])
    co = co.tocsr(); co.sort_indices()
    out['co/indptr'], out['co/indices'], out['co/data'] = co.indptr.astype(np.int64), co.indices.astype(np.int32), co.data.astype(np.int64)
    out['co/dtype'] = np.array(str(co.dtype))
    co_all = Team.gen_skill_coverage(tv, tempfile.mkdtemp(), skipteams=None).tocsr(); co_all.sort_indices()
    out['co_all/indptr'], out['co_all/indices'], out['co_all/data'] = co_all.indptr.astype(np.int64), co_all.indices.astype(np.int32), co_all.data.astype(np.int64)
    test = np.asarray(sp_['test'])
    X = tv['skill'][test]
    rng = np.random.default_rng(5)
    Y_ = rng.permutation(len(test) * tv['member'].shape[1]).reshape(len(test), -1).astype(np.float32) / (len(test) * tv['member'].shape[1])
    df, df_mean = refmetric.calculate_skill_coverage(X, Y_, co, per_instance=True, topks='2,5,10')
    out['random/Y_'] = Y_
    for c in df.columns: out[f'random/{c}'] = df[c].to_numpy(dtype=np.float64)
    predfile = f'{root}/{RUN}/f0.test.pred'
    if os.path.isfile(predfile):
        _stub_omegaconf()
        Yp = torch.load(predfile, map_location='cpu', weights_only=False)['y_pred']
        Yp = Yp.to_dense().numpy() if Yp.is_sparse else Yp.numpy()
        if all(len(np.unique(r)) == len(r) for r in Yp):  # (a dense, tie-free prediction: the ranking is unambiguous)
            df, _ = refmetric.calculate_skill_coverage(X, Yp, co, per_instance=True, topks='2,5,10')
            for c in df.columns: out[f'pred/{c}'] = df[c].to_numpy(dtype=np.float64)
    np.savez_compressed(f'{HERE}/staging_{key}.npz', **out)
    print('staging', key, 'co', co.shape, co.nnz, 'coverage means', {c: float(np.mean(out[f'random/{c}'])) for c in df.columns}, 'pred' if any(k.startswith('pred/') for k in out) else '')


if __name__ == '__main__':
    assert ref_shim.available(), 'run this where /root/reference is mounted'
    if sys.argv[1:] == ['staging']:
        for key in TOYS: record_staging(key)
        sys.exit(0)
    if sys.argv[1:] == ['tntf']:
        record_tntf()
        sys.exit(0)
    if sys.argv[1:] == ['dense']:
        record_trajectory('unigram_b', key='gith', tag='dense16', dense_d=16, b=32, h=[24], e=8, lr=0.01)
        sys.exit(0)
    for key in TOYS: dump_toy(key)
    t = record_trajectory('unigram_b')
    committed = np.load(f'{HERE}/toy_dblp.npz')
    for k in (0, 1, 2):  # the recorded run must BE the committed run
        assert int(t[f'f{k}/e']) == int(committed[f'ckpt/f{k}/e'])
        assert abs(float(t[f'f{k}/t_loss']) - float(committed[f'ckpt/f{k}/t_loss'])) < 1e-6, k
        assert abs(float(t[f'f{k}/v_loss']) - float(committed[f'ckpt/f{k}/v_loss'])) < 1e-6, k
        for n in ('layers.0.weight', 'layers.1.bias'):
            assert np.abs(t[f'f{k}/final/{n}'] - committed[f'ckpt/f{k}/{n}']).max() < 1e-6
    print('G2: recorded unigram_b trajectory reproduces the committed checkpoints')
    record_trajectory('uniform')
    record_trajectory('unigram')
    record_trajectory('unigram_b', key='imdb', tag='unigram_b_small', b=4, h=[16, 8], e=6)  # short last batches + 3 layers
    record_trajectory('unigram_b', key='gith', tag='dense16', dense_d=16, b=32, h=[24], e=8, lr=0.01)  # dense (embedded) skill input, several batches per epoch
    record_tntf()
    for key in TOYS: record_staging(key)
