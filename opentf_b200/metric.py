"""Ranking metrics for *.pred files: the consumer side of the hot path (reference: src/evl/metric.py).

`calculate_metrics` gives the same numbers as the reference's pytrec_eval call (metric.py:5-35) for the measure
families OpeNTF asks for -- P_k, recall_k, ndcg_cut_k, map_cut_k, success_k (src/__config__.yaml:91) -- computed
directly from the trec_eval definitions, so pytrec_eval is not needed.  Ranking follows trec_eval: score
descending, ties by document id string descending.  Pinned by the reference's committed eval CSVs
(tests/test_metric.py).
"""
import numpy as np
import scipy.sparse as sp

FAMILIES = ('P', 'recall', 'ndcg_cut', 'map_cut', 'success')


def parse_metrics(metrics):
    """['P_2,5,10', 'recall_2,5,10', ...] -> [(family, [k...])]"""
    out = []
    for m in metrics:
        fam, _, ks = m.rpartition('_')
        if fam not in FAMILIES: raise ValueError(f'unsupported trec measure {m!r} (supported: {FAMILIES})')
        out.append((fam, [int(k) for k in ks.split(',')]))
    return out


def _ranked_hits(cols, vals, rel, depth):
    """relevance (0/1) of the first `depth` documents in trec_eval order."""
    if len(vals) > depth:  # metric.py:19-21: first-stage cut by value
        cut = np.partition(vals, len(vals) - depth)[len(vals) - depth]
        keep = np.nonzero(vals >= cut)[0]
        cols, vals = cols[keep], vals[keep]
    order = sorted(range(len(vals)), key=lambda i: (float(vals[i]), 'd' + str(int(cols[i]))), reverse=True)[:depth]
    return np.fromiter((cols[i] in rel for i in order), dtype=np.float64, count=len(order))


def calculate_metrics(Y, Y_, topK=None, per_instance=False, metrics=('P_2,5,10', 'recall_2,5,10', 'ndcg_cut_2,5,10')):
    import pandas as pd
    assert Y.shape == Y_.shape, f'Shape mismatch between truth Y {Y.shape} vs preds Y_ {Y_.shape}!'
    fams = parse_metrics(metrics)
    kmax = max(k for _, ks in fams for k in ks)
    N, E = Y.shape
    Yc = sp.csr_matrix(Y)
    Yr = sp.csr_matrix(Y_) if sp.issparse(Y_) else None
    first_stage = min(topK, E) if topK else E
    names = [f'{f}_{k}' for f, ks in fams for k in ks]
    res = {n: np.zeros(N) for n in names}
    disc = 1.0 / np.log2(np.arange(2, kmax + 2))
    for i in range(N):
        rel = set(Yc.indices[Yc.indptr[i]:Yc.indptr[i + 1]].tolist())
        if Yr is not None: cols, vals = Yr.indices[Yr.indptr[i]:Yr.indptr[i + 1]], Yr.data[Yr.indptr[i]:Yr.indptr[i + 1]]
        else: cols, vals = np.arange(E), np.asarray(Y_[i])
        if len(vals) > first_stage:  # keep the first_stage best by value (metric.py:19-28), then rank
            keep = np.argsort(-vals, kind='stable')[:first_stage]
            cols, vals = cols[keep], vals[keep]
        hits = _ranked_hits(cols, vals, rel, kmax)
        R = len(rel)
        csum = np.cumsum(hits)
        for f, ks in fams:
            for k in ks:
                h = hits[:k]
                if f == 'P': v = h.sum() / k
                elif f == 'recall': v = h.sum() / R if R else 0.0
                elif f == 'success': v = float(h.sum() > 0)
                elif f == 'ndcg_cut': v = (h * disc[:len(h)]).sum() / disc[:min(R, k)].sum() if R else 0.0
                else: v = ((csum[:len(h)] / np.arange(1, len(h) + 1)) * h).sum() / R if R else 0.0  # map_cut
                res[f'{f}_{k}'][i] = v
    df = pd.DataFrame(res)
    df_mean = df.mean().to_frame('mean').rename_axis('metrics')
    return (df if per_instance else None), df_mean


EVAL_MAXK = 1024  # candidates per team ntf_eval_ranked ranks in shared memory


def calculate_metrics_device(Y, Y_, topK=None, per_instance=False, metrics=('P_2,5,10', 'recall_2,5,10', 'ndcg_cut_2,5,10'), device='cuda:0',
                             chunk=8192):
    """`calculate_metrics` with the per-team loop (metric.py:12-33) on the GPU: the candidates of each team (the stored top-K of a
    sparse .pred, or the first-stage top-K of a dense one, selected by ntf_topk_select) are ranked and scored by ntf_eval_ranked against
    the team's member row; only [n,5,nk] numbers come back.  Same frames as the host version.  No CPU fallback: raises without the library."""
    import pandas as pd
    import torch
    from . import ops
    assert Y.shape == Y_.shape, f'Shape mismatch between truth Y {Y.shape} vs preds Y_ {Y_.shape}!'
    fams = parse_metrics(metrics)
    ks = sorted({k for _, kk in fams for k in kk})
    N, E = Y.shape
    first_stage = min(topK, E) if topK else E
    if ks[-1] > EVAL_MAXK: raise ValueError(f'cut-off {ks[-1]} beyond the {EVAL_MAXK} candidates ntf_eval_ranked ranks per team')
    Yc = sp.csr_matrix(Y); Yc.sort_indices()
    dev = torch.device(device)
    out = np.zeros((N, 5, len(ks)))
    Yr = sp.csr_matrix(Y_) if sp.issparse(Y_) else None
    for r0 in range(0, N, chunk):
        r1 = min(N, r0 + chunk); n = r1 - r0
        m_ptr = torch.as_tensor((Yc.indptr[r0:r1 + 1] - Yc.indptr[r0]).astype(np.int32), device=dev)
        m_idx = torch.as_tensor(Yc.indices[Yc.indptr[r0]:Yc.indptr[r1]].astype(np.int32), device=dev)
        if m_idx.numel() == 0: m_idx = torch.zeros(1, dtype=torch.int32, device=dev)
        if Yr is not None:
            # sparse .pred: ONLY the stored entries are candidates (metric.py:15-23) -- a row with fewer stored scores than the cut-off
            # retrieves fewer documents, it never gains unstored experts.  Rows are padded with -1; a row holding more than `cap` entries
            # is cut to its cap best by value first (metric.py:19-21: argpartition on the stored values).
            cap = min(first_stage, EVAL_MAXK)
            lens = np.diff(Yr.indptr[r0:r1 + 1])
            K = max(int(min(lens.max(initial=0), cap)), 1)
            ci = np.full((n, K), -1, np.int32); cv = np.zeros((n, K), np.float32)
            short = lens <= cap
            ls = np.where(short, lens, 0)
            starts = Yr.indptr[r0:r1]
            src = np.repeat(starts, ls) + (np.arange(int(ls.sum())) - np.repeat(np.cumsum(ls) - ls, ls))
            pos = np.arange(int(ls.sum())) - np.repeat(np.cumsum(ls) - ls, ls)
            rows = np.repeat(np.arange(n), ls)
            ci[rows, pos] = Yr.indices[src]; cv[rows, pos] = Yr.data[src]
            for r in np.nonzero(~short)[0]:
                a, b = Yr.indptr[r0 + r], Yr.indptr[r0 + r + 1]
                v = Yr.data[a:b]
                keep = np.argpartition(-v, cap - 1)[:cap]
                ci[r, :cap] = Yr.indices[a:b][keep]; cv[r, :cap] = v[keep]
            idx, vals = torch.as_tensor(ci, device=dev), torch.as_tensor(cv, device=dev)
        else:  # first-stage retrieval on the device (metric.py:12,19-28)
            dense = np.ascontiguousarray(Y_[r0:r1], dtype=np.float32)
            K = min(first_stage, EVAL_MAXK)
            P = torch.as_tensor(dense, device=dev)
            vals = torch.empty(n, K, dtype=torch.float32, device=dev); idx = torch.empty(n, K, dtype=torch.int32, device=dev)
            ops.topk_select(P, n, E, K, 1.0, vals, idx)
        o = torch.empty(n, 5, len(ks), dtype=torch.float64, device=dev)
        ops.eval_ranked(idx, vals, m_ptr, m_idx, ks, o)
        out[r0:r1] = o.cpu().numpy()
    df = pd.DataFrame({f'{f}_{k}': out[:, FAMILIES.index(f), ks.index(k)] for f, kk in fams for k in kk})
    df_mean = df.mean().to_frame('mean').rename_axis('metrics')
    return (df if per_instance else None), df_mean


def calculate_auc_roc(Y, Y_, curve=False):
    """metric.py:37-42 (micro-averaged; densifies like the reference)."""
    from sklearn import metrics as skm
    assert Y.shape == Y_.shape
    yt = np.asarray(Y.todense()) if sp.issparse(Y) else np.asarray(Y)
    yp = np.asarray(Y_.todense()) if sp.issparse(Y_) else np.asarray(Y_)
    auc = skm.roc_auc_score(yt, yp, average='micro', multi_class='ovr')
    if curve:
        fpr, tpr, _ = skm.roc_curve(yt.ravel(), yp.ravel())
        return auc, (fpr, tpr)
    return auc, None


def calculate_skill_coverage(X, Y_, expertskillvecs, per_instance=False, topks='2,5,10', device='cuda:0'):
    """metric.py:44-73: fraction of a team's required skills held by the union of its top-k recommended experts.  The per-team loop runs on the
    device (staging.calculate_skill_coverage -> ntf_skill_coverage); no host implementation in the product."""
    from . import staging
    return staging.calculate_skill_coverage(X, Y_, expertskillvecs, per_instance, topks, device)
