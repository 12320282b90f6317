"""Dense (embedded) skill input, SURVEY.md 8f-2: main.py:148-153 replaces teamsvecs['skill'] by d2v/gnn vectors and ntf.py:24 hands the
row through as is, so layer 0 is a plain dense layer.  Parity against the trajectory recorded from the UNMODIFIED reference on seeded
dense vectors (tests/golden/traj_gith_dense16.npz) and against the CPU oracle; everything through the C ABI."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import fnn_oracle as O
from test_gpu_e2e import Replay, base_cfg, make

pytestmark = pytest.mark.gpu


def _tv(toy):
    _, member, splits, _ = toy('gith')
    t = np.load(os.path.join(GOLDEN, 'traj_gith_dense16.npz'))
    return {'skill': t['dense_x'], 'member': member.tolil()}, splits, t


def test_rows_gather_is_exact():
    from opentf_b200 import ops
    rng = np.random.default_rng(0)
    for n_src, n, d in ((100, 257, 16), (33, 5, 7), (1000, 1000, 128)):
        src = torch.from_numpy(rng.standard_normal((n_src, d)).astype(np.float32)).cuda()
        rows = torch.from_numpy(rng.integers(0, n_src, n).astype(np.int32)).cuda()
        dst = torch.empty(n, d, device='cuda')
        ops.rows_gather(rows, n, d, src, dst)
        assert torch.equal(dst, src[rows.long()])
        ops.rows_gather(None, min(n, n_src), d, src, dst)
        assert torch.equal(dst[:min(n, n_src)], src[:min(n, n_src)])


def test_recorded_reference_trajectory_on_dense_input_is_retraced(toy, tmp_path):
    """Fnn.learn fed the reference's initial weights, batch order and negatives reproduces the reference's run on dense input:
    stop epoch, epoch losses to 1e-5 relative, final weights to 2e-5 absolute (fp32 mode) -- same bars as the sparse G2 test."""
    tv, splits, t = _tv(toy)
    cfg = base_cfg(b=32, e=8, h=[24], lr=0.01)
    m = make(tmp_path, cfg)
    m.replay = Replay(t, splits, cfg['b'])
    m.learn(tv, splits, None)
    assert m.engine.dense_input
    for k in range(3):
        ck = torch.load(f'{m.output}/f{k}.pt', weights_only=False)
        assert ck['e'] == int(t[f'f{k}/e'])
        assert abs(ck['t_loss'] - float(t[f'f{k}/t_loss'])) <= 1e-5 * ck['t_loss']
        assert abs(ck['v_loss'] - float(t[f'f{k}/v_loss'])) <= 1e-5 * ck['v_loss']
        for n_, w in ck['model_state_dict'].items():
            assert tuple(w.shape) == t[f'f{k}/final/{n_}'].shape  # torch layout [out,in] for layer 0 too
            assert np.abs(w.numpy() - t[f'f{k}/final/{n_}']).max() < 2e-5, (k, n_)


@pytest.mark.parametrize('topK', [None, 7])
def test_predictions_on_dense_input_match_the_oracle(toy, tmp_path, topK):
    """test(): sigmoid(model(X)) for the reference's final weights -- dense .pred to 1e-6, sparse top-K .pred with identical indices"""
    tv, splits, t = _tv(toy)
    cfg = base_cfg(b=32, e=8, h=[24], lr=0.01)
    m = make(tmp_path, cfg)
    E = tv['member'].shape[1]
    for k in range(3):
        sd = {n: torch.from_numpy(t[f'f{k}/final/{n}']) for n in ('layers.0.weight', 'layers.0.bias', 'layers.1.weight', 'layers.1.bias')}
        torch.save({'model_state_dict': sd, 'cfg': cfg, 'f': k, 'e': 0, 't_loss': 0.0, 'v_loss': 0.0}, f'{m.output}/f{k}.pt')
    m.test(tv, splits, dict(on_train=False, per_epoch=False, topK=topK))
    for k in range(3):
        layers = [(torch.from_numpy(t[f'f{k}/final/layers.{i}.weight']), torch.from_numpy(t[f'f{k}/final/layers.{i}.bias'])) for i in range(2)]
        ref = O.predict(layers, tv['skill'], splits['test'], 32)
        y = torch.load(f'{m.output}/f{k}.test.pred', weights_only=False)['y_pred']
        if topK is None:
            assert not y.is_sparse and (y - ref).abs().max() <= 1e-6
        else:
            r = O.topk_sparse(ref, topK)
            assert y.is_sparse and tuple(y.shape) == (len(splits['test']), E)
            assert torch.equal(y.indices(), r.indices()) and (y.values() - r.values()).abs().max() <= 1e-6


def test_tensor_core_output_layer_on_dense_input(toy):
    """precision tf32: dense layer 0 (fp32 CUDA cores; 2*B*d*h flop, negligible) feeds the tcgen05 output layer: one train step
    against the oracle at the tensor-core tolerance (loss 2e-3 relative, gradients 5e-3 in norm)."""
    from opentf_b200 import _lib
    from opentf_b200.engine import Engine
    rng = np.random.default_rng(3)
    N, d, E, B, h = 512, 128, 1024, 256, 128
    if not _lib.lib().ntf_tc_supported(B, h, E, 0): pytest.skip('shape not on the tensor-core path')
    import scipy.sparse as sp
    X = rng.standard_normal((N, d)).astype(np.float32)
    member = sp.random(N, E, density=3.0 / E, random_state=4, format='csr', dtype=np.float32); member.data[:] = 1
    member = member.astype(np.uint8)
    torch.manual_seed(0)
    layers = O.init_params(d, [h], E)
    for precision, tol, gtol in (('fp32', 1e-5, 1e-5), ('tf32', 2e-3, 5e-3)):
        eng = Engine(d, [h], E, 'cuda:0', precision=precision, tpw=10, tnw=1, nsd='uniform', ns=5, max_batch=B, dense_input=True)
        eng.stage(X, member)
        eng.load_state_dict({f'layers.{i}.{n}': v for i, (W, b) in enumerate(layers) for n, v in (('weight', W), ('bias', b))})
        rows = rng.permutation(N)[:B]
        spd = eng.split(rows)
        neg = rng.integers(0, E, (B, 5))
        Xb, y = O.densify(X, rows), O.densify(member, rows)
        logits, acts, pre = O.forward(layers, Xb)
        w = O.loss_weights(y, torch.as_tensor(neg), 10, 1)
        loss_ref = O.bce_with_logits(logits, y, w).sum(1).mean().item()
        grads = O.backward(layers, acts, pre, y, w)
        eng.step(spd, 0, B, True, lr=1e-3, loss_slot=0, neg_host=neg)
        assert abs(eng.loss_buf[0].item() - loss_ref) <= tol * abs(loss_ref)
        for i in range(2):
            g = eng.view(f'layers.{i}.weight', eng.grads).cpu()
            assert tuple(g.shape) == tuple(grads[i][0].shape)
            # tf32: a logit that 10-bit operands move across the output lrelu's kink switches its slope (1 <-> 0.01); those rare entries
            # reach layer 0 through dA, so its gradient is compared in norm at 2e-2 (measured 7e-3), the output layer's own at 5e-3
            assert ((g - grads[i][0]).norm() / grads[i][0].norm()).item() <= (gtol if (i == 1 or precision == 'fp32') else 2e-2), (precision, i)
