"""GPU parity tests of the fused top-K inference kernel (csrc/infer_topk.cu, fnn.py:204-218 + pkgmgr.py:125-134): the K best experts per team
selected from the tcgen05 product without the [B,E] score matrix.  The checker is the oracle's ranking (O.topk_rows: value descending, ties ->
lower expert id) applied to the dense scores the library's own unfused path writes for the same operands (ntf_infer_scores, tensor-core mode):
values must be bit-equal, index sets equal; where distinct logits round to the SAME probability the fused kernel ranks by logit, so an index may
differ from the oracle's only inside a group of exactly tied scores (the contract: "bit-exact modulo exact score ties")."""
import os

import numpy as np
import pytest
import torch

from oracle import fnn_oracle as O
from test_gpu_kernels import DEV, rand_csr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    from opentf_b200 import ops as _ops
    return _ops


@pytest.fixture(scope='module')
def ws(ops):
    return ops.Workspace(torch.device(DEV))


def run_fused(ops, ws, A, W, b, K, e_lo=0):
    B, h = A.shape; E = W.shape[0]
    W16 = torch.empty(E * h, dtype=torch.float16, device=DEV)
    ops.to_half(W, E * h, W16)
    vals, idx = torch.empty(B, K, device=DEV), torch.empty(B, K, dtype=torch.int32, device=DEV)
    ops.infer_topk(A, W16, b, B, h, E, K, vals, idx, ws, e_lo=e_lo)
    torch.cuda.synchronize()
    return vals.cpu().numpy(), idx.cpu().numpy()


def check_against_dense(P, vals, idx, K, e_lo=0):
    """P: the dense device scores [B,E] (numpy).  values bit-equal to the oracle's ranking of P; ids equal except inside exact score ties"""
    v_ref, i_ref = O.topk_rows(P, K)
    assert (vals == v_ref).all(), f'values differ at {np.argwhere(vals != v_ref)[:5]}'
    cols = idx - e_lo
    assert (cols >= 0).all() and (cols < P.shape[1]).all()
    assert (np.take_along_axis(P, cols, axis=1) == vals).all()  # every returned id really has the returned score
    for n in range(P.shape[0]): assert len(set(cols[n].tolist())) == K  # no expert twice
    diff = cols != i_ref
    if diff.any():  # only inside groups of exactly tied scores (rank order within a tie / which members of a tie at the K-th place)
        for n, k in np.argwhere(diff):
            tied = (v_ref[n] == v_ref[n, k]).sum() > 1 or (P[n] == v_ref[n, k]).sum() > 1
            assert tied, (n, k, cols[n, k], i_ref[n, k])
    return int(diff.sum())


@pytest.mark.parametrize('B,E,K', [(1, 4096, 10), (37, 3001, 2), (130, 5000, 10), (256, 4224, 100), (1000, 40000, 10), (1000, 40000, 128), (300, 44774, 20), (129, 70, 2), (200, 40000, 1000), (64, 32768, 1024), (130, 9000, 250)])
def test_fused_topk_matches_oracle_ranking_of_the_dense_scores(ops, ws, B, E, K):
    torch.manual_seed(B + E + K)
    h = 128
    A, W, b = torch.randn(B, h).abs().to(DEV), (torch.randn(E, h) * 0.2).to(DEV), (torch.randn(E) * 0.5).to(DEV)
    assert ops.infer_topk_supported(B, h, E, K)
    P = torch.empty(B, E, device=DEV)
    ops.infer_scores(1, A, W, b, B, h, E, P, ws)  # the unfused tensor-core path on the same operands
    vals, idx = run_fused(ops, ws, A, W, b, K)
    ndiff = check_against_dense(P.cpu().numpy(), vals, idx, K)
    assert ndiff <= max(2, B * K // (200 if K <= 128 else 100))  # rounding-collapsed ties are rare (less so deep down the ranking, where the scores crowd)


def test_fused_topk_exact_ties_follow_the_lower_id_rule(ops, ws):
    """duplicated expert rows (exactly equal logits) and an all-zero layer (every expert tied): the library's tie rule is the oracle's"""
    torch.manual_seed(3)
    B, h, E, K = 70, 128, 2048, 16
    A = torch.randn(B, h).abs().to(DEV)
    W = torch.randn(E, h) * 0.02   # small logits: no sigmoid saturation, so distinct logits (almost always) give distinct probabilities
    bias = torch.randn(E) * 0.2
    W[1024:] = W[:1024]; bias[1024:] = bias[:1024]  # expert j+1024 is a copy of expert j: every score appears twice
    W, bias = W.to(DEV), bias.to(DEV)
    P = torch.empty(B, E, device=DEV)
    ops.infer_scores(1, A, W, bias, B, h, E, P, ws)
    vals, idx = run_fused(ops, ws, A, W, bias, K)
    v_ref, i_ref = O.topk_rows(P.cpu().numpy(), K)
    check_against_dense(P.cpu().numpy(), vals, idx, K)
    for n in range(B):  # a duplicated pair (j, j+1024) has EQUAL logits: the copy with the lower id must come first, adjacent to its twin
        row = idx[n].tolist()
        for k, e in enumerate(row):
            if e < 1024 and (e + 1024) in row: assert row.index(e + 1024) == k + 1, (n, row)
            if e >= 1024 and (e - 1024) in row: assert row.index(e - 1024) == k - 1, (n, row)
    Z = torch.zeros(E, h, device=DEV); zb = torch.zeros(E, device=DEV)
    ops.infer_scores(1, A, Z, zb, B, h, E, P, ws)
    vals, idx = run_fused(ops, ws, A, Z, zb, K)
    assert (idx == np.arange(K)[None, :]).all() and (vals == P.cpu().numpy()[:, :K]).all() and abs(float(vals[0, 0]) - 0.5) < 1e-6


def test_fused_topk_many_equal_block_maxima_restart_the_select_on_block_numbers(ops, ws):
    """more than 128 blocks of a team with EQUAL maxima (an all-zero layer over 20000 experts, then a few experts raised above the plateau): the
    select's tie path ranks the equal blocks by block number, so the winners after the raised experts are the lowest expert ids"""
    B, h, E, K = 40, 128, 20000, 12
    A = torch.randn(B, h, generator=torch.Generator().manual_seed(9)).abs().to(DEV)
    Z, zb = torch.zeros(E, h, device=DEV), torch.zeros(E, device=DEV)
    vals, idx = run_fused(ops, ws, A, Z, zb, K)
    assert (idx == np.arange(K)[None, :]).all() and (np.abs(vals - 0.5) < 1e-6).all()
    raised = [19999, 7, 12345, 4100, 4101]  # two of them in one block
    zb[raised] = torch.tensor([3.0, 2.5, 2.0, 1.5, 1.0], device=DEV)
    P = torch.empty(B, E, device=DEV)
    ops.infer_scores(1, A, Z, zb, B, h, E, P, ws)
    vals, idx = run_fused(ops, ws, A, Z, zb, K)
    want = raised + [e for e in range(E) if e not in raised][:K - len(raised)]
    assert (idx == np.asarray(want)[None, :]).all(), idx[0]
    check_against_dense(P.cpu().numpy(), vals, idx, K)


def test_fused_topk_large_k_on_a_plateau_of_equal_scores_takes_the_refinement_path(ops, ws):
    """K = 1000 with far more than 4096 candidates at or above the bar (every expert scores the same but for a few): the final kernel's radix
    refinement over the unique (score, expert) composites must still return exactly the oracle's K, lowest expert ids first inside the plateau"""
    B, h, E, K = 9, 128, 40000, 1000
    A = torch.randn(B, h, generator=torch.Generator().manual_seed(4)).abs().to(DEV)
    Z, zb = torch.zeros(E, h, device=DEV), torch.zeros(E, device=DEV)
    raised = [39999, 3, 20000, 20001, 31]
    zb[raised] = torch.tensor([2.0, 1.5, 1.0, 0.75, 0.5], device=DEV)
    lowered = list(range(100, 164))  # a whole run of experts below the plateau: they must be skipped
    zb[lowered] = -1.0
    P = torch.empty(B, E, device=DEV)
    ops.infer_scores(1, A, Z, zb, B, h, E, P, ws)
    vals, idx = run_fused(ops, ws, A, Z, zb, K)
    want = raised + [e for e in range(E) if e not in raised and e not in lowered][:K - len(raised)]
    assert (idx == np.asarray(want)[None, :]).all(), (idx[0][:12], idx[0][-5:])
    check_against_dense(P.cpu().numpy(), vals, idx, K)


def test_fused_topk_of_an_expert_shard_reports_global_ids(ops, ws):
    torch.manual_seed(5)
    B, h, E, K, e_lo = 33, 128, 1000, 5, 7000
    A, W, b = torch.randn(B, h).abs().to(DEV), (torch.randn(E, h) * 0.2).to(DEV), (torch.randn(E) * 0.5).to(DEV)
    P = torch.empty(B, E, device=DEV)
    ops.infer_scores(1, A, W, b, B, h, E, P, ws)
    vals, idx = run_fused(ops, ws, A, W, b, K, e_lo=e_lo)
    check_against_dense(P.cpu().numpy(), vals, idx, K, e_lo=e_lo)


def test_engine_topk_fused_equals_unfused_path():
    """Engine.topk (what Fnn.test calls) with and without the fused kernel on a staged split, rows taken at an offset"""
    from opentf_b200.engine import Engine
    rng = np.random.default_rng(11)
    N, S, E, K, B = 700, 300, 6000, 10, 256
    skill, member = rand_csr(rng, N, S, 1, 8), rand_csr(rng, N, E, 1, 4)
    torch.manual_seed(1)
    layers = O.init_params(S, [128], E)
    eng = Engine(S, [128], E, DEV, precision='tf32', nsd='uniform', ns=5, max_batch=B)
    eng.stage(skill, member)
    eng.load_state_dict({f'layers.{i}.{n}': t for i, (W, b) in enumerate(layers) for n, t in (('weight', W), ('bias', b))})
    sp = eng.split(np.arange(N))
    assert eng.fused_topk_ok(B, K)
    scores = torch.empty(B, E, device=DEV)
    out = {}
    for mode in ('1', '0'):
        os.environ['NTF_FUSED_TOPK'] = mode
        try:
            vals, idx = torch.empty(B, K, device=DEV), torch.empty(B, K, dtype=torch.int32, device=DEV)
            eng.topk(sp, 300, B, K, scores, vals, idx)
            torch.cuda.synchronize()
            out[mode] = (vals.cpu().numpy(), idx.cpu().numpy())
        finally:
            os.environ.pop('NTF_FUSED_TOPK', None)
    check_against_dense(scores.cpu().numpy(), out['1'][0], out['1'][1], K)
    assert (out['1'][0] == out['0'][0]).all()
    # and against the fp32 oracle end to end (tensor-core tolerance on the scores: 2e-3)
    X = O.densify(skill, np.arange(300, 300 + B))
    logits, _, _ = O.forward(layers, X)
    p_ref = torch.sigmoid(logits).numpy()
    got = np.take_along_axis(p_ref, out['1'][1].astype(np.int64), axis=1)
    assert np.abs(got - out['1'][0]).max() < 2e-3
    kth = np.sort(p_ref, axis=1)[:, -K]
    assert (got >= kth[:, None] - 4e-3).all()  # every selected expert is within tolerance of the oracle's K-th score
