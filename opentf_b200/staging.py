"""the sparse staging steps either side of the training path, run on the device (csrc/staging.cu) behind the reference's
own entry points:

    csr_from_lists / teamsvecs_from_lists   src/cmn/team.py:148-173, 240-276 (Team.bucketing over Team.get_one_hot: id lists -> multi-hot rows)
    gen_skill_coverage                      src/cmn/team.py:302-341 (member^T . skill, test teams left out, cached as skillcoverage.pkl)
    calculate_skill_coverage                src/evl/metric.py:44-73 (the per-team loop; called from Ntf.evaluate, src/mdl/ntf.py:72-78)

torch is the device-memory plumbing; every number is produced by libntf_b200.so.  No CPU fallback: without the library / a GPU the calls raise."""
import os
import pickle

import numpy as np
import scipy.sparse as sp


def _dev(device):
    import torch
    from . import util
    if not str(device).startswith('cuda'): raise RuntimeError(f'opentf_b200.staging runs on a CUDA device only (got {device}); there is no host implementation in the product')
    if not torch.cuda.is_available(): raise RuntimeError('opentf_b200.staging needs a CUDA device (none is visible); there is no host implementation in the product')
    return torch.device(util.first_device(device) if not isinstance(device, torch.device) else device)


def _i32(a, dev):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32), device=dev)


def csr_from_lists(indptr, ids, n_cols, device='cuda:0', dtype='u1'):
    """ragged id lists (indptr [n+1], ids in any order, duplicates allowed) -> scipy CSR [n, n_cols] of ones with sorted, duplicate-free rows:
    what Team.bucketing builds through a dense one-hot row per team (team.py:148-173)"""
    from . import ops
    dev = _dev(device)
    indptr = np.asarray(indptr, dtype=np.int64)
    if indptr[-1] >= 2 ** 31: raise ValueError('more than 2^31 ids: stage the teams in blocks')
    ws = ops.Workspace(dev)
    ptr, idx = ops.csr_from_lists(_i32(indptr, dev), _i32(ids, dev), int(n_cols), ws)
    ptr, idx = ptr.cpu().numpy(), idx.cpu().numpy()
    return sp.csr_matrix((np.ones(len(idx), dtype=dtype), idx, ptr), shape=(len(indptr) - 1, int(n_cols)))


def teamsvecs_from_lists(skill_lists, member_lists, n_skills, n_members, device='cuda:0'):
    """the 'skill' and 'member' matrices of teamsvecs (team.py:240-276 stacks Team.bucketing's blocks) from per-team id lists, lil like the
    reference's pickle"""
    out = {}
    for key, lists, n_cols in (('skill', skill_lists, n_skills), ('member', member_lists, n_members)):
        lens = np.fromiter((len(l) for l in lists), dtype=np.int64, count=len(lists))
        indptr = np.concatenate([[0], np.cumsum(lens)])
        ids = np.concatenate([np.asarray(l, dtype=np.int64) for l in lists]) if len(lists) and indptr[-1] else np.zeros(0, np.int64)
        out[key] = csr_from_lists(indptr, ids, n_cols, device).tolil()
    return out


def cooccurrence(member, skill, skipteams=None, device='cuda:0'):
    """member^T . skill as scipy CSR [E, S] with the dtype and the wrap-around of the reference's product (team.py:333,335: uint8 matrices multiply
    in uint8, a count of 256 stores an explicit 0); rows of `skipteams` left out (team.py:327-332)"""
    import torch
    from . import ops
    dev = _dev(device)
    M, S = sp.csr_matrix(member), sp.csr_matrix(skill)
    assert M.shape[0] == S.shape[0], f'member {M.shape} and skill {S.shape} disagree on the number of teams'
    M.sort_indices(); S.sort_indices()
    skip = None
    if skipteams is not None:
        flags = np.zeros(M.shape[0], dtype=np.uint8)
        flags[np.asarray(list(skipteams), dtype=np.int64)] = 1
        skip = torch.as_tensor(flags, device=dev)
    ws = ops.Workspace(dev)
    ptr, idx, val = ops.cooccur(_i32(M.indptr, dev), _i32(M.indices, dev), _i32(S.indptr, dev), _i32(S.indices, dev), M.shape[1], S.shape[1], skip, ws)
    dtype = np.result_type(M.dtype, S.dtype)
    co = sp.csr_matrix((val.cpu().numpy().astype(dtype), idx.cpu().numpy(), ptr.cpu().numpy()), shape=(M.shape[1], S.shape[1]))
    co.eliminate_zeros()  # (a wrapped-around count: scipy's product does not store it either)
    return co


def gen_skill_coverage(teamsvecs, output, skipteams=None, device='cuda:0'):
    """Team.gen_skill_coverage (team.py:302-341): load {output}/skillcoverage.pkl, or build the member-skill co-occurrence matrix (on the device)
    and cache it there"""
    if not os.path.isdir(output): os.makedirs(output)
    filepath = f'{output}/skillcoverage.pkl'
    try:
        with open(filepath, 'rb') as f: member_skill_co = pickle.load(f)
        assert member_skill_co.shape == (teamsvecs['member'].shape[1], teamsvecs['skill'].shape[1]), 'Incorrect matrix size!'
        return member_skill_co
    except FileNotFoundError:
        member_skill_co = cooccurrence(teamsvecs['member'], teamsvecs['skill'], skipteams, device)
        with open(filepath, 'wb') as f: pickle.dump(member_skill_co, f)
        return member_skill_co


def ranked_experts(Y_, K, device):
    """[n, K] int32 device tensor: every team's K best experts of Y_ in rank order (score descending, ties -> lower id), -1 where a sparse row stores
    fewer than K scores (the reference would go on into the unstored zeros in numpy's argsort order, metric.py:62: arbitrary, not reproduced)"""
    import torch
    from . import ops
    dev = _dev(device)
    n, E = Y_.shape
    K = min(int(K), E)
    vals = torch.empty(n, K, dtype=torch.float32, device=dev); idx = torch.empty(n, K, dtype=torch.int32, device=dev)
    if not sp.issparse(Y_):
        for r0 in range(0, n, 4096):
            r1 = min(n, r0 + 4096)
            P = torch.as_tensor(np.ascontiguousarray(Y_[r0:r1], dtype=np.float32), device=dev)
            ops.topk_select(P, r1 - r0, E, K, 1.0, vals[r0:r1], idx[r0:r1])
        return idx
    Yr = sp.csr_matrix(Y_); Yr.sort_indices()
    lens = np.diff(Yr.indptr)
    C = max(int(lens.max(initial=0)), K, 1)  # stored scores per row, padded: column position -> (score, expert)
    ci = np.full((n, C), -1, np.int32); cv = np.full((n, C), -np.inf, np.float32)
    rows = np.repeat(np.arange(n), lens)
    pos = np.arange(len(rows)) - np.repeat(Yr.indptr[:-1], lens)
    ci[rows, pos] = Yr.indices; cv[rows, pos] = Yr.data
    P, ids = torch.as_tensor(cv, device=dev), torch.as_tensor(ci, device=dev)
    where = torch.empty(n, K, dtype=torch.int32, device=dev)
    ops.topk_select(P, n, C, K, 1.0, vals, where)  # rank the stored scores (positions ascend with the expert id: the tie rule carries over)
    return torch.gather(ids, 1, where.long()).contiguous()


def calculate_skill_coverage(X, Y_, expertskillvecs, per_instance=False, topks='2,5,10', device='cuda:0'):
    """metric.py:44-73 with the per-team loop on the device (ntf_skill_coverage): same frames as the reference's function"""
    import pandas as pd
    import torch
    from . import ops
    assert X.shape[0] == Y_.shape[0], f'Shape[0] mismatch for number of teams in true skills in X {X.shape} vs experts (skill holders) in preds Y_ {Y_.shape}!'
    dev = _dev(device)
    ks = [int(k) for k in topks.split(',')]
    Xc = sp.csr_matrix(X); Xc.sort_indices()
    co = sp.csr_matrix(expertskillvecs).copy()
    co.eliminate_zeros()  # metric.py:66 keeps a skill only where the stored (possibly wrapped-around) count is > 0
    co.sort_indices()
    idx = ranked_experts(Y_, max(ks), dev)
    out = torch.empty(X.shape[0], len(ks), dtype=torch.float64, device=dev)
    ops.skill_coverage(idx, _i32(Xc.indptr, dev), _i32(Xc.indices, dev), _i32(co.indptr, dev), _i32(co.indices, dev), ks, out)
    o = out.cpu().numpy()
    df_skc = pd.DataFrame({f'skill_coverage_{k}': o[:, j] for j, k in enumerate(ks)})
    return df_skc, df_skc.mean().to_frame('mean').rename_axis('metrics')
