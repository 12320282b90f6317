"""CPU restatement of the reference's staging steps of SURVEY 8(f-4).  TEST INFRASTRUCTURE ONLY: imported by tests/ and by bench.py's
cpu_baseline leg as the checker / the CPU arm, never by the product (opentf_b200/).

Pinned (tests/test_oracle_golden.py) against tests/golden/staging_*.npz, which tests/golden/make_golden.py records from the UNMODIFIED reference
functions Team.gen_skill_coverage (src/cmn/team.py:302-341) and calculate_skill_coverage (src/evl/metric.py:44-73) on the reference's
committed toy teamsvecs, and against the committed teamsvecs themselves for the one-hot rows (the output of Team.bucketing, team.py:148-173)."""
import numpy as np
import scipy.sparse as sp


def rows_from_lists(indptr, ids, n_cols, dtype='u1'):
    """team.py:17-38 + 148-173: Team.get_one_hot writes x[0, id] = 1 for every id of the team (a duplicate writes the same 1 again), bucketing
    stores the dense row into a lil_matrix: the row's columns are the distinct ids, ascending."""
    n = len(indptr) - 1
    data = sp.lil_matrix((n, n_cols), dtype=dtype)
    for i in range(n):
        x = np.zeros((1, n_cols), dtype=dtype)
        for c in ids[indptr[i]:indptr[i + 1]]: x[0, c] = 1
        data[i] = x
    return data.tocsr()


def cooccurrence(member, skill, skipteams=None):
    """team.py:325-335: the member and skill rows of the skipped (test) teams are emptied, then scipy.sparse.csr_matrix(np.dot(member^T, skill)).
    The matrices are uint8 (team.py:154), so is the product: a count of 256 wraps to a stored 0."""
    member, skill = sp.lil_matrix(member).copy(), sp.lil_matrix(skill).copy()
    if skipteams is not None:
        for i in skipteams:
            member.rows[i] = []; member.data[i] = []
            skill.rows[i] = []; skill.data[i] = []
    co = sp.csr_matrix(np.dot(member.transpose(), skill))
    co.sort_indices()
    return co


def skill_coverage(X, Y_, expertskillvecs, topks='2,5,10'):
    """metric.py:53-69, line by line: experts ranked by np.argsort(row)[::-1]; the union of the first k experts' skill rows (max over rows > 0);
    covered required skills / required skills.  -> {k: array[teams]}"""
    teams = Y_.shape[0]
    Xc, Sk = sp.csr_matrix(X), sp.csr_matrix(expertskillvecs)
    cov = {int(k): np.zeros(teams) for k in topks.split(',')}
    for t in range(teams):
        row = np.asarray(Y_.getrow(t).toarray()).ravel() if sp.issparse(Y_) else np.asarray(Y_[t])
        ranked = np.argsort(row)[::-1]
        for k in cov:
            X_ = np.asarray((Sk[ranked[:k]].max(axis=0) > 0).todense()).ravel()
            need = np.asarray(Xc[t].todense()).ravel()
            cov[k][t] = (X_ * need).sum() / need.sum()
    return cov
