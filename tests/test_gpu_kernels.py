"""GPU parity tests, kernel by kernel: every call goes through the C ABI (opentf_b200.ops -> ctypes -> libntf_b200.so)
and is compared with the CPU oracle on the same seeded inputs.  Integer/index results must be bit-exact; fp32
results use the tolerance written next to each assert (fp32 mode: summation-order round-off only)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import fnn_oracle as O
from oracle import sampler_oracle as SO

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def ops():
    from opentf_b200 import ops as _ops
    return _ops


@pytest.fixture(scope='module')
def ws(ops):
    return ops.Workspace(torch.device(DEV))


def rand_csr(rng, n, ncols, lo, hi, empty_rows=()):
    rows, cols = [], []
    for i in range(n):
        k = 0 if i in empty_rows else int(rng.integers(lo, hi + 1))
        c = rng.choice(ncols, size=min(k, ncols), replace=False)
        rows += [i] * len(c); cols += list(c)
    m = sp.csr_matrix((np.ones(len(rows), dtype=np.uint8), (rows, cols)), shape=(n, ncols))
    m.sort_indices()
    return m


def dev_csr(m):
    return (torch.from_numpy(m.indptr.astype(np.int32)).to(DEV), torch.from_numpy(np.concatenate([m.indices, [0]]).astype(np.int32)).to(DEV))


def dense(m):
    return torch.from_numpy(np.asarray(m.todense(), dtype=np.float32))


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(1e-30, np.abs(b).max())


# ------------------------------------------------------------------------------------------------ CSR staging
@pytest.mark.parametrize('n,ncols', [(1, 5), (37, 11), (5000, 300)])
def test_csr_gather_matches_scipy_row_selection(ops, ws, n, ncols):
    rng = np.random.default_rng(n)
    m = rand_csr(rng, n, ncols, 0, min(ncols, 9), empty_rows=(0,))
    indptr, indices = dev_csr(m)
    order = rng.permutation(n)[: max(1, n * 3 // 4)].astype(np.int32)
    k = len(order)
    sub = m[order]; sub.sort_indices()
    d_ptr = torch.empty(k + 1, dtype=torch.int32, device=DEV)
    d_idx = torch.empty(max(1, sub.nnz), dtype=torch.int32, device=DEV)
    d_row = torch.empty(max(1, sub.nnz), dtype=torch.int32, device=DEV)
    ops.csr_gather(torch.from_numpy(order).to(DEV), k, indptr, indices, d_ptr, d_idx, d_row, ws)
    assert (d_ptr.cpu().numpy() == sub.indptr).all()
    assert (d_idx.cpu().numpy()[:sub.nnz] == sub.indices).all()
    assert (d_row.cpu().numpy()[:sub.nnz] == np.repeat(np.arange(k), np.diff(sub.indptr))).all()


# ------------------------------------------------------------------------------------------------ input layer
@pytest.mark.parametrize('B,S,h', [(1, 7, 128), (33, 50, 128), (257, 1000, 20), (64, 40, 6)])
def test_csr_bag_fwd_matches_dense_first_layer(ops, B, S, h):
    rng = np.random.default_rng(B + h)
    X = rand_csr(rng, B, S, 1, 9, empty_rows=(0,) if B > 1 else ())
    torch.manual_seed(B)
    W, b = torch.randn(h, S), torch.randn(h)
    ref = O.lrelu(dense(X) @ W.t() + b)  # fnn.py:25 on the densified rows of ntf.py:23
    indptr, indices = dev_csr(X)
    A = torch.empty(B, h, device=DEV)
    ops.csr_bag_fwd(B, indptr.data_ptr(), indices, W.t().contiguous().to(DEV), b.to(DEV), S, h, A)
    assert rel_err(A.cpu(), ref) < 2e-6


@pytest.mark.parametrize('B,S,h', [(1, 7, 128), (33, 50, 128), (300, 1000, 20), (513, 97, 128), (6000, 200, 128)])
def test_csr_bag_bwd_matches_dense_gradient_and_is_deterministic(ops, ws, B, S, h):
    rng = np.random.default_rng(B * 7 + h)
    X = rand_csr(rng, B, S, 1, 9).tolil()
    if B > 100:  # heavy-tailed popularity: skill 3 sits in ~80% of the teams, skill 5 in ~30% (the per-CTA "hot" path)
        X[np.nonzero(rng.random(B) < 0.8)[0], 3] = 1
        X[np.nonzero(rng.random(B) < 0.3)[0], 5] = 1
    X = X.tocsr(); X.sort_indices()
    torch.manual_seed(B)
    dZ = torch.randn(B, h)
    ref = (dZ.t() @ dense(X)).t()  # dW[h,S] = dZ^T X, stored transposed [S,h]
    indptr, indices = dev_csr(X)
    ent_row = torch.from_numpy(np.repeat(np.arange(B), np.diff(X.indptr)).astype(np.int32)).to(DEV)
    outs = []
    for _ in range(2):
        dW = torch.full((S, h), float('nan'), device=DEV)
        ops.csr_bag_bwd(B, indptr.data_ptr(), indices, ent_row, 0, dZ.to(DEV), S, h, dW, ws)
        outs.append(dW.cpu())
    assert torch.equal(outs[0], outs[1])  # run-to-run bit-stable (no atomics)
    assert rel_err(outs[0], ref) < 2e-6
    untouched = np.asarray(X.sum(axis=0)).ravel() == 0
    assert (outs[0].numpy()[untouched] == 0).all()  # every row is written, zeros where the batch has no such skill


def test_csr_bag_on_a_batch_slice_of_a_larger_split(ops, ws):
    """kernels address a batch by pointer offset into the split's indptr (absolute offsets)."""
    rng = np.random.default_rng(3)
    n, S, h, b0, B = 90, 40, 128, 32, 25
    X = rand_csr(rng, n, S, 1, 6)
    torch.manual_seed(0)
    W, b, dZ = torch.randn(h, S), torch.randn(h), torch.randn(B, h)
    indptr, indices = dev_csr(X)
    ent_row = torch.from_numpy(np.repeat(np.arange(n), np.diff(X.indptr)).astype(np.int32)).to(DEV)
    A = torch.empty(B, h, device=DEV)
    ops.csr_bag_fwd(B, indptr.data_ptr() + 4 * b0, indices, W.t().contiguous().to(DEV), b.to(DEV), S, h, A)
    assert rel_err(A.cpu(), O.lrelu(dense(X[b0:b0 + B]) @ W.t() + b)) < 2e-6
    dW = torch.empty(S, h, device=DEV)
    ops.csr_bag_bwd(B, indptr.data_ptr() + 4 * b0, indices, ent_row, b0, dZ.to(DEV), S, h, dW, ws)
    assert rel_err(dW.cpu(), (dZ.t() @ dense(X[b0:b0 + B])).t()) < 2e-6


# ------------------------------------------------------------------------------------------------ hidden layers
@pytest.mark.parametrize('B,i,o', [(5, 7, 3), (130, 128, 64), (1500, 64, 128)])
def test_dense_layer_forward_backward(ops, ws, B, i, o):
    torch.manual_seed(B)
    A, W, b, dY = torch.randn(B, i), torch.randn(o, i) * 0.2, torch.randn(o), torch.randn(B, o)
    Y = torch.empty(B, o, device=DEV)
    ops.dense_fwd(A.to(DEV), W.to(DEV), b.to(DEV), B, i, o, 1, Y)
    ref = O.lrelu(A @ W.t() + b)
    assert rel_err(Y.cpu(), ref) < 3e-6
    dZ, db = torch.empty(B, o, device=DEV), torch.empty(o, device=DEV)
    ops.act_bwd(dY.to(DEV), Y, B, o, 1, dZ, db, ws)
    dz_ref = dY * torch.where(ref > 0, 1.0, 0.01)
    assert rel_err(dZ.cpu(), dz_ref) < 1e-6 and rel_err(db.cpu(), dz_ref.sum(0)) < 3e-6
    dW, dA = torch.empty(o, i, device=DEV), torch.empty(B, i, device=DEV)
    ops.dense_bwd(A.to(DEV), W.to(DEV), dZ, B, i, o, dW, dA, ws)
    assert rel_err(dW.cpu(), dz_ref.t() @ A) < 3e-6 and rel_err(dA.cpu(), dz_ref @ W) < 3e-6


# ------------------------------------------------------------------------------------------------ sampler
def member_lists(m):
    return [m.indices[m.indptr[i]:m.indptr[i + 1]] for i in range(m.shape[0])]


@pytest.mark.parametrize('E', [13, 1000, 70001])
def test_expert_cdf_is_exact(ops, ws, E):
    rng = np.random.default_rng(E)
    Y = rand_csr(rng, 200, E, 1, min(E, 6))
    indptr, indices = dev_csr(Y)
    counts, cdf = torch.empty(E, dtype=torch.int32, device=DEV), torch.empty(E, dtype=torch.int32, device=DEV)
    ops.expert_cdf(200, indptr.data_ptr(), indices, E, counts, cdf, ws)
    c_ref, cdf_ref = SO.expert_cdf(member_lists(Y), E)
    assert (counts.cpu().numpy() == c_ref).all() and (cdf.cpu().numpy() == cdf_ref).all()


@pytest.mark.parametrize('nsd', ['uniform', 'unigram', 'unigram_b'])
@pytest.mark.parametrize('B,E,ns', [(17, 13, 5), (64, 500, 5), (300, 40000, 8)])
def test_neg_sample_bit_exact_against_cpu_restatement(ops, ws, nsd, B, E, ns):
    from opentf_b200._lib import NSD
    rng = np.random.default_rng(B + E)
    Y = rand_csr(rng, B, E, 1, min(E - 1, 6))
    indptr, indices = dev_csr(Y)
    counts, cdf = torch.empty(E, dtype=torch.int32, device=DEV), torch.empty(E, dtype=torch.int32, device=DEV)
    ops.expert_cdf(B, indptr.data_ptr(), indices, E, counts, cdf, ws)
    cdf_host = cdf.cpu().numpy().astype(np.int64)
    for step, row0 in ((0, 0), (7, 123), (2 ** 33 + 5, 10)):
        neg = torch.empty(B, ns, dtype=torch.int32, device=DEV)
        ops.neg_sample(NSD[nsd], 0xDEADBEEF12345, step, row0, B, indptr.data_ptr(), indices, E, ns, cdf if nsd == 'unigram' else None, neg)
        ref = SO.sample_negatives(nsd, 0xDEADBEEF12345, step, row0, member_lists(Y), E, ns, cdf_host)
        got = neg.cpu().numpy()
        assert (got == ref).all()
        for n, mem in enumerate(member_lists(Y)):  # the reference's contract: distinct non-members
            live = got[n][got[n] >= 0]
            assert len(set(live)) == len(live)
            if nsd == 'uniform' or cdf_host[-1] > sum(cdf_host[j] - (cdf_host[j - 1] if j else 0) for j in mem):
                assert not set(live) & set(mem)
            if nsd == 'unigram_b' and E > 10 * B * 6:  # candidates are experts of the batch's other teams (plenty of them here)
                assert set(live) <= set(Y.indices.tolist())


def test_neg_sample_unigram_b_pool_is_the_global_batch(ops):
    """a data-parallel rank samples for its slice of the rows from the member CSR of the whole global batch"""
    from opentf_b200._lib import NSD
    rng = np.random.default_rng(3)
    Y = rand_csr(rng, 96, 5000, 2, 5)
    indptr, indices = dev_csr(Y)
    lists = member_lists(Y)
    full = torch.empty(96, 5, dtype=torch.int32, device=DEV)
    ops.neg_sample(NSD['unigram_b'], 11, 4, 0, 96, indptr.data_ptr(), indices, 5000, 5, None, full)
    for lo, hi in ((0, 32), (32, 96)):
        part = torch.empty(hi - lo, 5, dtype=torch.int32, device=DEV)
        ops.neg_sample(NSD['unigram_b'], 11, 4, lo, hi - lo, indptr.data_ptr() + 4 * lo, indices, 5000, 5, None, part, indptr.data_ptr(), 96)
        assert torch.equal(part, full[lo:hi])  # the same rows get the same negatives however the batch is cut across ranks
        ref = SO.sample_negatives('unigram_b', 11, 4, lo, lists[lo:hi], 5000, 5, None, pool_rows=lists)
        assert (part.cpu().numpy() == ref).all()


def test_neg_sample_fallbacks(ops, ws):
    from opentf_b200._lib import NSD
    # rows 0 and 2: all mass on their own members -> uniform over ALL experts (fnn.py:67-69); row 1: two weighted candidates, rest topped up
    Y = sp.csr_matrix(np.array([[1, 1, 0, 0, 0, 0], [0, 0, 1, 0, 0, 0], [1, 1, 1, 1, 0, 0]], dtype=np.uint8))
    cdf_host = np.cumsum([1, 1, 0, 0, 0, 0])
    indptr, indices = dev_csr(Y)
    cdf = torch.from_numpy(cdf_host.astype(np.int32)).to(DEV)
    neg = torch.empty(3, 4, dtype=torch.int32, device=DEV)
    ops.neg_sample(NSD['unigram'], 5, 1, 0, 3, indptr.data_ptr(), indices, 6, 4, cdf, neg)
    ref = SO.sample_negatives('unigram', 5, 1, 0, member_lists(Y), 6, 4, cdf_host)
    assert (neg.cpu().numpy() == ref).all()
    assert {0, 1} <= set(ref[1].tolist()) and 2 not in ref[1] and len(set(ref[1].tolist())) == 4
    for r in (0, 2): assert len(set(ref[r].tolist())) == 4 and min(ref[r]) >= 0
    # unigram_b, the pool is the batch itself: row 2 owns every expert of the batch -> ALL experts; rows 0, 1: the others' experts, then the top-up
    ops.neg_sample(NSD['unigram_b'], 5, 1, 0, 3, indptr.data_ptr(), indices, 6, 4, None, neg)
    ref = SO.sample_negatives('unigram_b', 5, 1, 0, member_lists(Y), 6, 4)
    assert (neg.cpu().numpy() == ref).all()
    assert {2, 3} <= set(ref[0].tolist()) and not {0, 1} & set(ref[0].tolist())
    assert {0, 1, 3} <= set(ref[1].tolist()) and 2 not in ref[1]
    assert len(set(ref[2].tolist())) == 4 and min(ref[2]) >= 0
    # a batch of one team: its own members are the whole pool
    ops.neg_sample(NSD['unigram_b'], 5, 1, 0, 1, indptr.data_ptr(), indices, 6, 4, None, neg)
    assert (neg.cpu().numpy()[:1] == SO.sample_negatives('unigram_b', 5, 1, 0, member_lists(Y)[:1], 6, 4)).all()
    # fewer negatives than ns (uniform): the two non-members, then -1 padding
    ops.neg_sample(NSD['uniform'], 5, 1, 0, 3, indptr.data_ptr(), indices, 6, 4, None, neg)
    assert sorted(neg.cpu().numpy()[2].tolist()) == [-1, -1, 4, 5]
    assert (neg.cpu().numpy() == SO.sample_negatives('uniform', 5, 1, 0, member_lists(Y), 6, 4, None)).all()


def test_special_bits_set_and_clear(ops):
    rng = np.random.default_rng(1)
    B, E, ns = 50, 1000, 5
    Y = rand_csr(rng, B, E, 1, 6)
    indptr, indices = dev_csr(Y)
    negs = rng.integers(-1, E, (B, ns)).astype(np.int32)
    pitch = (E + 31) // 32
    plane = torch.zeros(B, pitch, dtype=torch.int32, device=DEV)
    ops.special_bits(1, B, indptr.data_ptr(), indices, torch.from_numpy(negs).to(DEV), ns, E, plane, pitch)
    ref = np.asarray(Y.todense()).astype(bool)
    for n in range(B):
        for j in negs[n]:
            if j >= 0: ref[n, j] = True
    words = plane.cpu().numpy().view(np.uint32)
    got = ((words[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(B, -1)[:, :E].astype(bool)
    assert (got == ref).all()
    ops.special_bits(0, B, indptr.data_ptr(), indices, torch.from_numpy(negs).to(DEV), ns, E, plane, pitch)
    assert int(plane.abs().sum()) == 0


@pytest.mark.parametrize('B,E', [(50, 1000), (300, 129), (1, 5)])
def test_special_tiles_layout_set_and_clear(ops, B, E):
    """the tile-transposed planes the tensor-core kernel reads: word ((n/128)*Epad + j)*4 + (n%128)/32, bit n%32"""
    rng = np.random.default_rng(B)
    ns = 5
    Y = rand_csr(rng, B, E, 1, min(E - 1, 6))
    indptr, indices = dev_csr(Y)
    negs = rng.integers(-1, E, (B, ns)).astype(np.int32)
    words = ops.special_tiles_bytes(B, E) // 4
    Epad = (E + 127) // 128 * 128
    assert words == (B + 127) // 128 * Epad * 4
    ps, pm = torch.zeros(words, dtype=torch.int32, device=DEV), torch.zeros(words, dtype=torch.int32, device=DEV)
    ops.special_tiles(1, B, indptr.data_ptr(), indices, torch.from_numpy(negs).to(DEV), ns, E, ps, pm)
    member = np.asarray(Y.todense()).astype(bool)
    special = member.copy()
    for n in range(B):
        for j in negs[n]:
            if j >= 0: special[n, j] = True
    def unpack(plane):
        w = plane.cpu().numpy().view(np.uint32).reshape(-1, Epad, 4)                      # [tile][expert][4 words]
        bits = ((w[:, :, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(w.shape[0], Epad, 128)  # [tile][expert][team in tile]
        return bits.transpose(0, 2, 1).reshape(-1, Epad)[:B, :E].astype(bool)
    assert (unpack(ps) == special).all() and (unpack(pm) == member).all()
    ops.special_tiles(0, B, indptr.data_ptr(), indices, torch.from_numpy(negs).to(DEV), ns, E, ps, pm)
    assert int(ps.abs().sum()) == 0 and int(pm.abs().sum()) == 0


# ------------------------------------------------------------------------------------------------ output layer
def run_out_train(ops, ws, precision, A, W, b, Y, negs, tpw, tnw, train=True):
    from opentf_b200._lib import OutTrainArgs
    B, h = A.shape; E = W.shape[0]
    indptr, indices = dev_csr(Y)
    pitch = (E + 31) // 32
    plane = torch.zeros(B, pitch, dtype=torch.int32, device=DEV)
    negd = None if negs is None else torch.from_numpy(np.ascontiguousarray(negs, dtype=np.int32)).to(DEV)
    ops.special_bits(1, B, indptr.data_ptr(), indices, negd, 0 if negs is None else negs.shape[1], E, plane, pitch)
    Ad, Wd, bd = A.to(DEV), W.to(DEV), b.to(DEV)
    dW, db, dA, loss = torch.empty(E, h, device=DEV), torch.empty(E, device=DEV), torch.empty(B, h, device=DEV), torch.zeros(1, device=DEV)
    a = OutTrainArgs()
    a.A, a.W, a.b, a.special, a.pitch_words = Ad.data_ptr(), Wd.data_ptr(), bd.data_ptr(), plane.data_ptr(), pitch
    a.m_indptr, a.m_indices, a.B, a.h, a.E = indptr.data_ptr(), indices.data_ptr(), B, h, E
    a.tpw, a.tnw, a.loss_scale, a.loss_out = tpw, tnw, 1.0 / B, loss.data_ptr()
    if train: a.dW, a.db, a.dA = dW.data_ptr(), db.data_ptr(), dA.data_ptr()
    ops.out_train(0, precision, a, ws)
    torch.cuda.synchronize()
    return loss.cpu().item(), dW.cpu(), db.cpu(), dA.cpu()


def oracle_out(A, W, b, Y, negs, tpw, tnw):
    y = dense(Y)
    z = A @ W.t() + b
    x = O.lrelu(z)
    w = O.loss_weights(y, None if negs is None else torch.as_tensor(negs, dtype=torch.int64), tpw, tnw)
    loss = O.bce_with_logits(x, y, w).sum(1).mean().item()
    dz = w * (torch.sigmoid(x) - y) / A.shape[0] * torch.where(z > 0, 1.0, 0.01)
    return loss, dz.t() @ A, dz.sum(0), dz @ W


@pytest.mark.parametrize('B,h,E,ns', [(1, 8, 5, 2), (37, 24, 301, 5), (300, 128, 5000, 5), (1000, 128, 4097, 5)])
def test_out_train_fp32_matches_oracle(ops, ws, B, h, E, ns):
    rng = np.random.default_rng(B + E)
    torch.manual_seed(E)
    A, W, b = torch.randn(B, h).abs() * 0.3, torch.randn(E, h) * 0.3, torch.randn(E) * 0.1
    Y = rand_csr(rng, B, E, 1, min(E - 1, 5))
    negs = rng.integers(-1, E, (B, ns))  # includes -1 (no sample), duplicates and members: all must be harmless
    for tpw, tnw in ((10.0, 1.0), (1.0, 0.0)):
        loss, dW, db, dA = run_out_train(ops, ws, 0, A, W, b, Y, negs, tpw, tnw)
        l_ref, dW_ref, db_ref, dA_ref = oracle_out(A, W, b, Y, negs, tpw, tnw)
        assert abs(loss - l_ref) <= 1e-5 * abs(l_ref)  # loss rel-err <= 1e-5 in fp32 mode
        assert rel_err(dW, dW_ref) < 1e-5 and rel_err(db, db_ref) < 1e-5 and rel_err(dA, dA_ref) < 1e-5
    loss_v, _, _, _ = run_out_train(ops, ws, 0, A, W, b, Y, None, 10.0, 1.0, train=False)
    assert abs(loss_v - oracle_out(A, W, b, Y, None, 10.0, 1.0)[0]) <= 1e-5 * abs(loss_v)


def test_adam_matches_torch_semantics(ops):
    torch.manual_seed(0)
    n = 100_003
    p0 = torch.randn(n)
    mine, m, v = p0.clone().to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    ref = [p0.clone()]
    adam = O.Adam(ref, 1e-3)
    for t in range(1, 8):
        g = torch.randn(n) * (0.0 if t == 3 else 1.0)  # a zero-gradient step still moves the weights (momentum)
        adam.step([g])
        ops.adam_step(mine, g.to(DEV), m, v, n, 1e-3, 0.9, 0.999, 1e-8, t)
    assert rel_err(mine.cpu(), ref[0]) < 1e-6
    assert (mine.cpu() - ref[0]).abs().max() < 3e-7  # one ulp at the largest |p| (in [2,4)) after 7 steps


# ------------------------------------------------------------------------------------------------ inference + ranking
@pytest.mark.parametrize('B,h,E,K', [(3, 8, 13, 13), (50, 128, 3001, 10), (20, 128, 40000, 1000), (7, 16, 5000, 1)])
def test_scores_and_topk_match_oracle(ops, ws, B, h, E, K):
    torch.manual_seed(K)
    A, W, b = torch.randn(B, h).abs(), (torch.randn(E, h) * 0.2), torch.randn(E) * 0.1
    if K == 10: W = (W * 4).round() / 4  # quantised -> many exactly tied scores
    P = torch.empty(B, E, device=DEV)
    ops.infer_scores(0, A.to(DEV), W.to(DEV), b.to(DEV), B, h, E, P, ws)
    ref = torch.sigmoid(O.lrelu(A @ W.t() + b))
    assert (P.cpu() - ref).abs().max() < 5e-7
    vals, idx = torch.empty(B, K, device=DEV), torch.empty(B, K, dtype=torch.int32, device=DEV)
    ops.topk_select(P, B, E, K, 1.0, vals, idx)
    v_ref, i_ref = O.topk_rows(P.cpu(), K)  # rank the DEVICE scores: selection must be exact incl. the tie rule
    assert (idx.cpu().numpy() == i_ref).all() and (vals.cpu().numpy() == v_ref).all()


def test_topk_merge_of_expert_shards(ops):
    torch.manual_seed(0)
    B, E, K, G = 9, 4000, 20, 4
    P = torch.rand(B, E).round(decimals=3)
    v_ref, i_ref = O.topk_rows(P, K)
    per = E // G
    cv, ci = [], []
    for g in range(G):
        v, i = O.topk_rows(P[:, g * per:(g + 1) * per], K)
        cv.append(v); ci.append(i + g * per)
    vals, idx = torch.empty(B, K, device=DEV), torch.empty(B, K, dtype=torch.int32, device=DEV)
    ops.topk_merge(torch.from_numpy(np.stack(cv)).to(DEV), torch.from_numpy(np.stack(ci).astype(np.int32)).to(DEV), G, B, K, vals, idx)
    assert (idx.cpu().numpy() == i_ref).all() and (vals.cpu().numpy() == v_ref).all()


# ------------------------------------------------------------------------------------------------ Flipout parameter pass
def test_flipout_prepare_and_grads(ops, ws):
    torch.manual_seed(0)
    n = 70_001
    mu, rho, eps, gd = torch.randn(n) * 0.1, torch.randn(n) * 0.1 - 3, torch.randn(n), torch.randn(n)
    delta, kl = torch.empty(n, device=DEV), torch.zeros(1, device=DEV)
    ops.flipout_prepare(mu.to(DEV), rho.to(DEV), eps.to(DEV), n, 1.0 / n, delta, kl, ws)
    sig = O.softplus(rho)
    assert rel_err(delta.cpu(), sig * eps) < 1e-6
    assert abs(kl.item() - O.kl_mean(mu, sig).item()) < 1e-5 * abs(kl.item())
    B = 17
    g_mu0 = torch.randn(n)
    g_mu, g_rho = g_mu0.clone().to(DEV), torch.empty(n, device=DEV)
    ops.flipout_grads(mu.to(DEV), rho.to(DEV), eps.to(DEV), gd.to(DEV), n, 1.0 / (n * B), g_mu, g_rho)
    assert rel_err(g_mu.cpu(), g_mu0 + mu / (n * B)) < 1e-6
    assert rel_err(g_rho.cpu(), (gd * eps + (sig - 1 / sig) / (n * B)) * torch.sigmoid(rho)) < 1e-5


def test_device_noise_generators(ops):
    n = 1 << 20
    z = torch.empty(n, device=DEV)
    ops.fill_normal(1, 2, 3, n, z)
    z2 = torch.empty(n, device=DEV)
    ops.fill_normal(1, 2, 3, n, z2)
    assert torch.equal(z, z2)  # counter based: reproducible
    assert abs(z.mean().item()) < 5e-3 and abs(z.std().item() - 1) < 5e-3 and abs((z ** 4).mean().item() - 3) < 0.05
    bits = torch.zeros(n // 32, dtype=torch.int32, device=DEV)
    ops.fill_sign_bits(1, 2, 3, n // 32, bits)
    ones = sum(bin(int(x) & 0xFFFFFFFF).count('1') for x in bits.cpu().numpy()[:4096])
    assert abs(ones / (4096 * 32) - 0.5) < 0.01
    A = torch.randn(64, 40, device=DEV)
    As = torch.empty_like(A)
    ops.apply_sign(A, bits, 2, 64, 40, As)
    w = bits.cpu().numpy().view(np.uint32)[:128].reshape(64, 2)
    sgn = 1 - 2 * ((w[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(64, 64)[:, :40].astype(np.float32)
    assert torch.equal(As.cpu(), A.cpu() * torch.from_numpy(sgn))
