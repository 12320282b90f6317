"""GPU parity tests of the tcgen05 (NTF_TF32) output-layer kernels against the CPU oracle, output by output so that a
failure points at one product: raw logits and loss (TMA + tf32 MMA + TMEM read-back + epilogue), db (epilogue),
dW (fp16 MMA with the activations as MN-major operand), dA (fp16 MMA with the weight tile as MN-major operand).

Tolerances (stated per assert): the forward product reads fp32 operands as TF32 (10-bit mantissa, truncated), so a
logit may be off by up to ~2^-10 * sum_k |a_k w_k|; the backward products see dz, A and W rounded to 10-bit mantissas."""
import os

import numpy as np
import pytest
import torch

from oracle import fnn_oracle as O
from test_gpu_kernels import DEV, dense, dev_csr, oracle_out, rand_csr, rel_err

pytestmark = pytest.mark.gpu
TF32_EPS = 2.0 ** -10


@pytest.fixture(scope='module')
def ops():
    from opentf_b200 import ops as _ops
    return _ops


@pytest.fixture(scope='module')
def ws(ops):
    return ops.Workspace(torch.device(DEV))


def run_tc(ops, ws, A, W, b, Y, negs, tpw, tnw, train=True, zdbg=False, planes=False):
    """planes=False: the persistent kernel + sparse correction pass (out_tc2.cu; the member CSR and `neg` name the pairs of weight tpw);
    planes=True: round 1's kernel (out_tc.cu MODE 0, still the Flipout layer's pipeline), fed the tile-transposed bit planes"""
    from opentf_b200._lib import OutTrainArgs
    B, h = A.shape; E = W.shape[0]
    indptr, indices = dev_csr(Y)
    words = ops.special_tiles_bytes(B, E) // 4
    plane_s, plane_m = torch.zeros(words, dtype=torch.int32, device=DEV), torch.zeros(words, dtype=torch.int32, device=DEV)
    negd = None if negs is None else torch.from_numpy(np.ascontiguousarray(negs, dtype=np.int32)).to(DEV)
    if planes: ops.special_tiles(1, B, indptr.data_ptr(), indices, negd, 0 if negs is None else negs.shape[1], E, plane_s, plane_m)
    Ad, Wd, bd = A.to(DEV), W.to(DEV), b.to(DEV)
    dW, db, dA, loss = torch.full((E, h), float('nan'), device=DEV), torch.full((E,), float('nan'), device=DEV), torch.empty(B, h, device=DEV), torch.zeros(1, device=DEV)
    Z = torch.full((B, E), float('nan'), device=DEV) if zdbg else None
    a = OutTrainArgs()
    a.A, a.W, a.b = Ad.data_ptr(), Wd.data_ptr(), bd.data_ptr()
    if planes: a.special_t, a.member_t = plane_s.data_ptr(), plane_m.data_ptr()
    elif negd is not None: a.neg, a.ns = negd.data_ptr(), negs.shape[1]
    a.m_indptr, a.m_indices, a.B, a.h, a.E = indptr.data_ptr(), indices.data_ptr(), B, h, E
    a.tpw, a.tnw, a.loss_scale, a.loss_out = tpw, tnw, 1.0 / B, loss.data_ptr()
    if train: a.dW, a.db, a.dA = dW.data_ptr(), db.data_ptr(), dA.data_ptr()
    if zdbg: os.environ['NTF_TC_ZDBG'] = str(Z.data_ptr())
    try:
        ops.out_train(0, 1, a, ws)
        torch.cuda.synchronize()
    finally:
        os.environ.pop('NTF_TC_ZDBG', None)
    assert int(plane_s.abs().sum()) == 0 and int(plane_m.abs().sum()) == 0  # the kernel clears the plane words it consumed
    return loss.cpu().item(), dW.cpu(), db.cpu(), dA.cpu(), (Z.cpu() if zdbg else None)


def make_case(B, E, seed, h=128, ns=5):
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    A = torch.randn(B, h).abs() * 0.3 * (torch.rand(B, h) < 0.7)  # post-lrelu-like activations with zeros
    W, b = torch.randn(E, h) * 0.2, torch.randn(E) * 0.1
    Y = rand_csr(rng, B, E, 1, min(E - 1, 5))
    negs = rng.integers(-1, E, (B, ns))
    return A, W, b, Y, negs


# (1000, 40000) and (1000, 44774) are the exact output-layer shapes of BASELINE configs[1] / configs[2] (C2: b=1000, E=40000; C3: E=44774);
# (256, 300): fewer expert tiles than SMs -> every expert tile cut along the batch; (2000, 20000): 157 expert tiles x 16 batch tiles in ranges
SHAPES = [(64, 128), (100, 200), (1, 13), (1000, 4097), (333, 40000), (256, 300), (2000, 20000), (1000, 40000), (1000, 44774)]


@pytest.mark.parametrize('B,E', SHAPES)
def test_tc_forward_logits_and_loss(ops, ws, B, E):
    A, W, b, Y, negs = make_case(B, E, B + E)
    loss, _, _, _, Z = run_tc(ops, ws, A, W, b, Y, negs, 10.0, 1.0, train=False, zdbg=True)
    z_ref = A @ W.t() + b
    bound = 2.1 * TF32_EPS * (A.abs() @ W.abs().t()) + 1e-6  # both operands truncated to 10 mantissa bits
    assert not torch.isnan(Z).any()
    assert ((Z - z_ref).abs() <= bound).all(), float(((Z - z_ref).abs() / bound).max())
    l_ref = oracle_out(A, W, b, Y, negs, 10.0, 1.0)[0]
    assert abs(loss - l_ref) <= 1e-3 * abs(l_ref), (loss, l_ref)  # stated TF32 tolerance on the loss: 1e-3 relative


def oracle_from_logits(Z, A, W, Y, negs, tpw, tnw):
    """loss and gradients of SURVEY 9.2 evaluated AT the given pre-activation logits Z (the device's TF32 logits): the lrelu
    slope is discontinuous at z = 0, so a logit within TF32 error of zero may legitimately take either branch; testing the
    backward products at the device's own logits separates that from the accuracy of the products themselves."""
    y = dense(Y)
    x = O.lrelu(Z)
    w = O.loss_weights(y, None if negs is None else torch.as_tensor(negs, dtype=torch.int64), tpw, tnw)
    loss = O.bce_with_logits(x, y, w).sum(1).mean().item()
    dz = w * (torch.sigmoid(x) - y) / A.shape[0] * torch.where(Z > 0, 1.0, 0.01)
    return loss, dz.t() @ A, dz.sum(0), dz @ W


@pytest.mark.parametrize('B,E', SHAPES)
@pytest.mark.parametrize('tpw,tnw', [(10.0, 1.0), (1.0, 0.0)])
def test_tc_train_step_gradients(ops, ws, B, E, tpw, tnw):
    A, W, b, Y, negs = make_case(B, E, 7 * B + E)
    loss, dW, db, dA, Z = run_tc(ops, ws, A, W, b, Y, negs, tpw, tnw, zdbg=True)
    l_true = oracle_out(A, W, b, Y, negs, tpw, tnw)[0]
    assert abs(loss - l_true) <= 1e-3 * abs(l_true) + 1e-6  # against the fp32 oracle: stated TF32 tolerance on the loss
    l_ref, dW_ref, db_ref, dA_ref = oracle_from_logits(Z, A, W, Y, negs, tpw, tnw)
    assert abs(loss - l_ref) <= 2e-5 * abs(l_ref) + 1e-6    # given the logits, the loss arithmetic is fp32 (MUFU ex2/lg2)
    assert not (torch.isnan(dW).any() or torch.isnan(db).any() or torch.isnan(dA).any())
    # given the logits: db is an fp32 sum (MUFU sigmoid); dW / dA carry dz, A, W rounded to 10-bit mantissas
    assert rel_err(db, db_ref) < 2e-5, ('db', rel_err(db, db_ref))
    assert rel_err(dW, dW_ref) < 2e-3, ('dW', rel_err(dW, dW_ref))
    assert rel_err(dA, dA_ref) < 2e-3, ('dA', rel_err(dA, dA_ref))


def test_tc_operand_range_edges(ops, ws):
    """the tensor-core path reads fp16 operands: same 10-bit mantissa as TF32 but a 5-bit exponent (max 65504, normals down to 6.1e-5,
    subnormals to 6e-8).  Activations up to ~1e4 and weights spread over 1e-7 .. 1 (a third of them in fp16's subnormal range, where the
    ABSOLUTE rounding error is <= 2^-25) still meet the elementwise logit bound; beyond 65504 is outside the kernel's contract."""
    rng = np.random.default_rng(99)
    torch.manual_seed(99)
    B, E, h = 200, 1000, 128
    A = torch.randn(B, h).abs() * torch.tensor(10.0) ** torch.randint(-3, 5, (B, h)).float()          # 1e-3 .. 1e4
    W = torch.randn(E, h) * torch.tensor(10.0) ** torch.randint(-7, 1, (E, h)).float() * 1e-2           # 1e-9 .. 1e-2 scale: many fp16 subnormals / zeros
    b = torch.randn(E) * 0.1
    Y = rand_csr(rng, B, E, 1, 5)
    negs = rng.integers(-1, E, (B, 5))
    assert float(A.max()) < 65504 and float(A.max()) > 1e4
    loss, _, _, _, Z = run_tc(ops, ws, A, W, b, Y, negs, 10.0, 1.0, train=False, zdbg=True)
    z_ref = A.double() @ W.double().t() + b.double()
    bound = 2.1 * TF32_EPS * (A.abs() @ W.abs().t()) + (A.abs().sum(1, keepdim=True) * 2.0 ** -25) + 1e-6  # relative part + subnormal absolute part
    assert not torch.isnan(Z).any() and not torch.isinf(Z).any()
    assert ((Z.double() - z_ref).abs() <= bound.double()).all(), float(((Z.double() - z_ref).abs() / bound.double()).max())


@pytest.mark.parametrize('B,E', [(100, 200), (1000, 4097), (256, 300)])
def test_tc_round1_kernel_fed_bit_planes_still_matches(ops, ws, B, E):
    """out_tc.cu MODE 0 (one CTA per expert tile, fix-up from the tile-transposed planes inside the epilogue): the pipeline the Flipout layer
    still runs on; kept selectable for the Fnn layer (planes given / NTF_TC_V1=1) for A/B runs"""
    A, W, b, Y, negs = make_case(B, E, 7 * B + E)
    loss, dW, db, dA, Z = run_tc(ops, ws, A, W, b, Y, negs, 10.0, 1.0, zdbg=True, planes=True)
    l_ref, dW_ref, db_ref, dA_ref = oracle_from_logits(Z, A, W, Y, negs, 10.0, 1.0)
    assert abs(loss - l_ref) <= 2e-5 * abs(l_ref) + 1e-6
    assert rel_err(db, db_ref) < 2e-5 and rel_err(dW, dW_ref) < 2e-3 and rel_err(dA, dA_ref) < 2e-3


def test_tc_matches_fp32_kernel_on_device(ops, ws):
    """same inputs through both precisions of the library: away from the lrelu kink the two implementations agree"""
    from test_gpu_kernels import run_out_train
    A, W, b, Y, negs = make_case(777, 5000, 5)
    l32, dW32, db32, dA32 = run_out_train(ops, ws, 0, A, W, b, Y, negs, 10.0, 1.0)
    l_tc, dW, db, dA, Z = run_tc(ops, ws, A, W, b, Y, negs, 10.0, 1.0, zdbg=True)
    assert abs(l_tc - l32) <= 1e-3 * abs(l32)
    z32 = A @ W.t() + b
    flipped = ((Z > 0) != (z32 > 0))                      # logits whose sign TF32 changed: each moves one dz entry by <= tpw*0.99/B
    allowed = 10.0 * 0.99 / 777 * flipped.sum(0).float()  # per expert, on db
    assert ((db - db32).abs() <= allowed + 2e-3 * db32.abs().max()).all()
    assert flipped.float().mean() < 2e-3                   # and such logits are rare


@pytest.mark.parametrize('B,E', [(64, 128), (1000, 4097), (50, 40000)])
def test_tc_inference_scores(ops, ws, B, E):
    A, W, b, _, _ = make_case(B, E, 3 * B + E)
    P = torch.full((B, E), float('nan'), device=DEV)
    ops.infer_scores(1, A.to(DEV), W.to(DEV), b.to(DEV), B, 128, E, P, ws)
    ref = torch.sigmoid(O.lrelu(A @ W.t() + b))
    assert not torch.isnan(P).any()
    assert (P.cpu() - ref).abs().max() < 1.5e-3  # |dp| <= |dz|/4, dz bounded as in the forward test
    k = min(10, E)
    top_ref = set(map(tuple, np.argwhere(ref.numpy() >= np.sort(ref.numpy(), axis=1)[:, -k][:, None] - 2e-3)))
    top = torch.topk(P.cpu(), k, dim=1).indices.numpy()
    assert all((n, int(j)) in top_ref for n in range(B) for j in top[n])  # every pick is within the stated tolerance of the true top-k


def test_tc_end_to_end_training_tracks_fp32(toy, tmp_path):
    """Fnn.learn in tf32 mode on a 128-wide model stays within 2e-3 of the fp32 run's epoch losses."""
    from opentf_b200.fnn import Fnn
    skill, member, splits, _ = toy('gith')
    tv = {'skill': skill.tolil(), 'member': member.tolil()}
    hist = {}
    for prec in ('fp32', 'tf32'):
        cfg = dict(b=8, e=5, ns=5, lr=0.001, es=10, h=[128], spe=0, l='bce', tpw=10, tnw=1, nsd='unigram_b', precision=prec)
        m = Fnn(str(tmp_path / prec), 'cuda:0', 0, cfg)
        m.learn(tv, {'test': splits['test'], 'folds': {0: splits['folds'][0]}}, None)
        from opentf_b200 import _lib
        assert m.engine.precision == (_lib.NTF_TF32 if prec == 'tf32' else _lib.NTF_FP32)
        hist[prec] = m.last_history[0]
    for (t32, v32), (ttc, vtc) in zip(hist['fp32'], hist['tf32']):
        assert abs(t32 - ttc) <= 2e-3 * t32 and abs(v32 - vtc) <= 2e-3 * v32
