"""GPU parity tests of the Bnn (Flipout) path: engine step, MC inference and the Bnn class end to end, against the CPU
oracle's restatement of bayesian-torch 0.5.0 LinearFlipout (oracle/fnn_oracle.py, SURVEY.md 9.5 -- "parity unpinned"
upstream of that restatement: the package is not vendored in the reference tree).  The Flipout draws (eps, +-1 signs)
are made on the host with the oracle's generator order and fed to both sides, so results compare to fp32 round-off."""
import numpy as np
import pytest
import torch

from oracle import fnn_oracle as O
from test_gpu_kernels import DEV, dense, rand_csr, rel_err

pytestmark = pytest.mark.gpu


def make_engine(S, hidden, E, B, skill, member, layers, **kw):
    from opentf_b200.engine import Engine
    eng = Engine(S, hidden, E, DEV, bayesian=True, precision=kw.get('precision', 'fp32'), tpw=10, tnw=1, nsd=kw.get('nsd', 'uniform'), ns=5, seed=3, max_batch=B)
    eng.stage(skill, member)
    sd = {}
    for i, L in enumerate(layers):
        sd[f'layers.{i}.mu_weight'], sd[f'layers.{i}.rho_weight'] = L['mu_w'], L['rho_w']
        sd[f'layers.{i}.mu_bias'], sd[f'layers.{i}.rho_bias'] = L['mu_b'], L['rho_b']
    eng.load_state_dict(sd)
    return eng


def grads_of(eng, i):
    g = lambda k, w: eng.view(f'layers.{i}.{k}_{w}', eng.grads).cpu()
    t = (lambda x: x.t()) if i == 0 else (lambda x: x)  # layer 0 is stored transposed
    return dict(mu_w=t(g('mu', 'weight')), rho_w=t(g('rho', 'weight')), mu_b=g('mu', 'bias'), rho_b=g('rho', 'bias'))


@pytest.mark.parametrize('B,S,hidden,E', [(37, 50, [24], 301), (64, 40, [16, 8], 100), (130, 27, [128], 1000), (1, 9, [8], 5)])
def test_flipout_step_matches_oracle(B, S, hidden, E):
    rng = np.random.default_rng(B + E)
    torch.manual_seed(B)
    skill, member = rand_csr(rng, B, S, 1, min(S, 6)), rand_csr(rng, B, E, 1, min(E - 1, 4))
    layers = O.init_flipout_params(S, hidden, E)
    noise = O.draw_flipout_noise(layers, B)
    neg = rng.integers(0, E, (B, 5))
    X, y = dense(skill), dense(member)
    logits, acts, pre = O.flipout_forward(layers, noise, X)
    w = O.loss_weights(y, torch.as_tensor(neg), 10, 1)
    kl = O.flipout_kl(layers)
    loss_ref = (O.bce_with_logits(logits, y, w).sum(1).mean() + kl / B).item()
    g_ref = O.flipout_backward(layers, noise, acts, pre, y, w)

    eng = make_engine(S, hidden, E, B, skill, member, layers)
    sp = eng.split(np.arange(B))
    p0 = eng.params.clone()
    eng.step(sp, 0, B, True, lr=1e-3, loss_slot=0, neg_host=neg, noise_host=noise)
    torch.cuda.synchronize()
    loss = eng.loss_buf[0].item()
    assert abs(loss - loss_ref) <= 2e-5 * abs(loss_ref), (loss, loss_ref)  # fp32 mode, summation order only
    for i in range(len(layers)):
        mine = grads_of(eng, i)
        for k in ('mu_w', 'mu_b', 'rho_w', 'rho_b'):
            assert rel_err(mine[k], g_ref[i][k]) < 3e-5, (i, k, rel_err(mine[k], g_ref[i][k]))
    # the optimiser moved every mu / rho by one Adam step of those gradients (|step| = lr at t = 1 wherever g != 0)
    moved = (eng.params - p0).abs().cpu()
    assert moved.max().item() <= 1.001e-3 and moved.max().item() > 0.9e-3

    # a validation step: same loss arithmetic (BCE + KL/B), fresh draws, no parameter change
    p1 = eng.params.clone()
    noise2 = O.draw_flipout_noise(layers, B)
    sd = eng.state_dict()
    layers2 = [dict(mu_w=sd[f'layers.{i}.mu_weight'], rho_w=sd[f'layers.{i}.rho_weight'], mu_b=sd[f'layers.{i}.mu_bias'], rho_b=sd[f'layers.{i}.rho_bias'])
               for i in range(len(layers))]
    logits2, _, _ = O.flipout_forward(layers2, noise2, X)
    ref2 = (O.bce_with_logits(logits2, y, w).sum(1).mean() + O.flipout_kl(layers2) / B).item()
    eng.step(sp, 0, B, False, loss_slot=1, neg_host=neg, noise_host=noise2)
    assert abs(eng.loss_buf[1].item() - ref2) <= 2e-5 * abs(ref2)
    assert torch.equal(eng.params, p1)


def test_flipout_step_on_a_batch_slice(B=48, S=30, E=200):
    """rows [16, 48) of a split: CSR offsets are absolute, entry signs are relative to the slice"""
    rng = np.random.default_rng(5)
    torch.manual_seed(5)
    skill, member = rand_csr(rng, B, S, 1, 5), rand_csr(rng, B, E, 1, 4)
    layers = O.init_flipout_params(S, [16], E)
    b0, Bs = 16, 32
    noise = O.draw_flipout_noise(layers, Bs)
    neg = rng.integers(0, E, (Bs, 5))
    X, y = dense(skill)[b0:], dense(member)[b0:]
    logits, acts, pre = O.flipout_forward(layers, noise, X)
    w = O.loss_weights(y, torch.as_tensor(neg), 10, 1)
    g_ref = O.flipout_backward(layers, noise, acts, pre, y, w)
    eng = make_engine(S, [16], E, B, skill, member, layers)
    sp = eng.split(np.arange(B))
    eng.step(sp, b0, Bs, True, lr=1e-3, neg_host=neg, noise_host=noise)
    for i in range(2):
        mine = grads_of(eng, i)
        for k in ('mu_w', 'mu_b', 'rho_w', 'rho_b'): assert rel_err(mine[k], g_ref[i][k]) < 3e-5, (i, k)


def test_mc_inference_matches_oracle(B=29, S=20, E=150, nmc=4):
    rng = np.random.default_rng(1)
    torch.manual_seed(1)
    skill, member = rand_csr(rng, B, S, 1, 4), rand_csr(rng, B, E, 1, 3)
    layers = O.init_flipout_params(S, [32], E)
    noises = [O.draw_flipout_noise(layers, B) for _ in range(nmc)]
    mc = np.stack([torch.sigmoid(O.flipout_forward(layers, nz, dense(skill))[0]).numpy() for nz in noises])
    eng = make_engine(S, [32], E, B, skill, member, layers)
    sp = eng.split(np.arange(B))
    out, scratch = torch.empty(B, E, device=DEV), torch.empty(B, E, device=DEV)
    ep, em = torch.empty(B, device=DEV), torch.empty(B, device=DEV)
    eng.scores_mc(sp, 0, B, nmc, out, scratch, ep, em, noise_host=noises)
    assert np.abs(out.cpu().numpy() - mc.mean(0)).max() < 1e-6
    assert np.abs(ep.cpu().numpy() - O.predictive_entropy(mc)).max() < 1e-4 * E  # sums of E terms of ~0.35
    assert np.abs(em.cpu().numpy() - O.mutual_information(mc)).max() < 2e-3      # a small difference of two such sums


def test_device_noise_is_reproducible_and_has_the_right_statistics(B=256, S=27, E=2000):
    rng = np.random.default_rng(2)
    torch.manual_seed(2)
    skill, member = rand_csr(rng, B, S, 1, 3), rand_csr(rng, B, E, 2, 4)
    layers = O.init_flipout_params(S, [128], E)
    losses = []
    for rep in range(2):
        eng = make_engine(S, [128], E, B, skill, member, layers)
        sp = eng.split(np.arange(B))
        for i in range(3): eng.step(sp, 0, B, True, lr=1e-3, loss_slot=i)
        losses.append(eng.loss_buf[:3].cpu().numpy().copy())
        if rep == 0:
            eps = eng.nview(eng.eps, '1.weight').cpu()
            assert abs(eps.mean().item()) < 0.01 and abs(eps.std().item() - 1) < 0.01
            bits = eng.sign_out[1].cpu().numpy().view(np.uint32)
            frac = np.unpackbits(bits.view(np.uint8)).mean()
            assert abs(frac - 0.5) < 0.01
    assert np.array_equal(losses[0], losses[1])  # counter RNG keyed by (seed, step): bit-reproducible
    assert losses[0][2] < losses[0][0]


def test_bnn_class_end_to_end(toy, tmp_path):
    """Bnn.learn / test on toy dblp: the reference's files, keys and layouts (G4: state_dict names/shapes of the committed
    bnn checkpoints), probabilities in (0,1), the last batch's uncertainty vectors, and a falling training loss."""
    from opentf_b200.bnn import Bnn
    skill, member, splits, z = toy('dblp')
    tv = {'skill': skill.tolil(), 'member': member.tolil()}
    cfg = dict(b=8, e=6, ns=5, lr=0.01, es=5, h=[128], spe=2, l='bce', tpw=10, tnw=1, nsd='unigram_b', nmc=3, precision='fp32')
    m = Bnn(str(tmp_path), 'cuda:0', 0, cfg)
    assert m.name().startswith('/bnn.b8.e6.ns5.lr0.01') and m.is_bayesian
    one = {'test': splits['test'], 'folds': {0: splits['folds'][0]}}
    m.learn(tv, one, None)
    ck = torch.load(f'{m.output}/f0.pt', weights_only=False)
    assert list(ck['model_state_dict']) == [str(n) for n in z['bnn/names']]
    for n in z['bnn/names']: assert tuple(ck['model_state_dict'][str(n)].shape) == tuple(z[f'bnn/f0/{n}'].shape)
    hist = m.last_history[0]
    assert hist[-1][0] < hist[0][0]
    m.test(tv, one, dict(on_train=False, per_epoch=True, topK=1000))
    p = torch.load(f'{m.output}/f0.test.pred', weights_only=False)
    y = p['y_pred']
    assert tuple(y.shape) == (len(splits['test']), member.shape[1]) and float(y.min()) > 0 and float(y.max()) < 1
    last = len(splits['test']) % 8 or 8
    assert set(p['uncertainty']) == {'pred', 'model'} and p['uncertainty']['pred'][0].shape == (last,)
    assert (p['uncertainty']['pred'][0] > 0).all() and np.abs(p['uncertainty']['model'][0]).max() < 1.0
    import os
    assert os.path.exists(f'{m.output}/f0.e0.pt') and os.path.exists(f'{m.output}/f0.test.e1.pred')
    # warm start from that checkpoint (tntf.py:38 chain): the loaded state is what test() saw
    m2 = Bnn(str(tmp_path / 'again'), 'cuda:0', 0, dict(cfg, e=1))
    m2.learn(tv, one, {0: f'{m.output}/f0.pt'})
    assert m2.last_history[0][0][0] < hist[0][0]


# ---- tensor-core Flipout output layer (out_tc_kernel modes 2 / 0 / 3): fp16 operands (10-bit mantissa) and fp16 perturbation-term /
# gradient tiles, fp32 accumulation.  Tolerances: loss 2e-3 relative; gradients 5e-3 in norm (a logit within rounding distance of the
# lrelu kink may take the other slope: rare entries, so the comparison is in norm, as for the Fnn kernel in test_gpu_tc.py) -- except the
# output layer's rho gradients, 2e-2: dW_delta = sum_n s_out*dz*a_s sums RANDOMLY SIGNED terms (norm ~ sqrt(B) terms) while a kink flip
# changes a term by its full size, so a fraction f of flipped logits costs sqrt(f) relative (measured f ~ 4e-5 -> 6e-3); the mu
# gradients sum coherently (norm ~ B terms) and see f^(1/2)/B^(1/2) of it.
# The (256, 40, [128], 300) case -- two FULL team tiles per CTA on a near-empty machine -- is the shape that exposed round 1's race in the
# hand-over of the perturbation-term tile (out_tc.cu: the single-stage Q slot was released before its reads had completed; fixed in round 2,
# DESIGN.md section 5); test_flipout_tensor_core_step_is_reproducible_on_fresh_engines below hammers it.
@pytest.mark.parametrize('B,S,hidden,E', [(130, 27, [128], 1000),
                                          (256, 40, [128], 300),
                                          (1000, 27, [128], 20000), (77, 30, [16, 128], 129)])
def test_flipout_step_tensor_core_matches_oracle(B, S, hidden, E):
    from opentf_b200 import _lib
    rng = np.random.default_rng(B + E)
    torch.manual_seed(B)
    skill, member = rand_csr(rng, B, S, 1, min(S, 6)), rand_csr(rng, B, E, 1, min(E - 1, 4))
    layers = O.init_flipout_params(S, hidden, E)
    noise = O.draw_flipout_noise(layers, B)
    neg = rng.integers(0, E, (B, 5))
    X, y = dense(skill), dense(member)
    logits, acts, pre = O.flipout_forward(layers, noise, X)
    w = O.loss_weights(y, torch.as_tensor(neg), 10, 1)
    loss_ref = (O.bce_with_logits(logits, y, w).sum(1).mean() + O.flipout_kl(layers) / B).item()
    g_ref = O.flipout_backward(layers, noise, acts, pre, y, w)
    eng = make_engine(S, hidden, E, B, skill, member, layers, precision='tf32')
    assert eng.precision == _lib.NTF_TF32  # the tcgen05 path is the one under test
    sp = eng.split(np.arange(B))
    eng.step(sp, 0, B, True, lr=1e-3, loss_slot=0, neg_host=neg, noise_host=noise)
    torch.cuda.synchronize()
    loss = eng.loss_buf[0].item()
    assert abs(loss - loss_ref) <= 2e-3 * abs(loss_ref), (loss, loss_ref)
    nrm = lambda a, b: ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
    errs, bad = {}, []
    for i in range(len(layers)):
        mine = grads_of(eng, i)
        for k in ('mu_w', 'mu_b', 'rho_w', 'rho_b'):
            tol = 2e-2 if (i == len(layers) - 1 and k.startswith('rho')) else 5e-3
            errs[(i, k)] = round(nrm(mine[k], g_ref[i][k]), 6)
            if not errs[(i, k)] < tol: bad.append((i, k))
    if bad:
        import warnings
        warnings.warn(f'flipout tc step B={B} E={E}: out of tolerance {bad}; all errors {errs}')  # (shown even when the case is xfail)
    assert not bad, (bad, errs)
    assert int(eng.special_t.abs().sum()) == 0 and int(eng.member_t.abs().sum()) == 0  # the planes were consumed
    # a second step on the same engine (planes clean, workspace reused) and a validation step
    noise2 = O.draw_flipout_noise(layers, B)
    sd = eng.state_dict()
    layers2 = [dict(mu_w=sd[f'layers.{i}.mu_weight'], rho_w=sd[f'layers.{i}.rho_weight'], mu_b=sd[f'layers.{i}.mu_bias'], rho_b=sd[f'layers.{i}.rho_bias'])
               for i in range(len(layers))]
    logits2, _, _ = O.flipout_forward(layers2, noise2, X)
    ref2 = (O.bce_with_logits(logits2, y, w).sum(1).mean() + O.flipout_kl(layers2) / B).item()
    p1 = eng.params.clone()
    eng.step(sp, 0, B, False, loss_slot=1, neg_host=neg, noise_host=noise2)
    assert abs(eng.loss_buf[1].item() - ref2) <= 2e-3 * abs(ref2)
    assert torch.equal(eng.params, p1)


def test_mc_inference_tensor_core_matches_oracle(B=200, S=20, E=1500, nmc=3):
    from opentf_b200 import _lib
    rng = np.random.default_rng(1)
    torch.manual_seed(1)
    skill, member = rand_csr(rng, B, S, 1, 4), rand_csr(rng, B, E, 1, 3)
    layers = O.init_flipout_params(S, [128], E)
    noises = [O.draw_flipout_noise(layers, B) for _ in range(nmc)]
    mc = np.stack([torch.sigmoid(O.flipout_forward(layers, nz, dense(skill))[0]).numpy() for nz in noises])
    eng = make_engine(S, [128], E, B, skill, member, layers, precision='tf32')
    assert eng.precision == _lib.NTF_TF32
    sp = eng.split(np.arange(B))
    out, scratch = torch.empty(B, E, device=DEV), torch.empty(B, E, device=DEV)
    ep, em = torch.empty(B, device=DEV), torch.empty(B, device=DEV)
    eng.scores_mc(sp, 0, B, nmc, out, scratch, ep, em, noise_host=noises)
    assert np.abs(out.cpu().numpy() - mc.mean(0)).max() < 2e-3  # probabilities in (0,1): absolute
    assert np.abs(ep.cpu().numpy() - O.predictive_entropy(mc)).max() < 2e-3 * E


def test_flipout_tensor_core_step_is_reproducible_on_fresh_engines(iters=60):
    """the stress loop that reproduced round 1's intermittent failure (scripts/flip_stress.py: 8 of 30 iterations off), as a test: the same
    Flipout tensor-core step on a FRESH engine every iteration (new allocations), after a step of another shape; every iteration must give
    bit-identical losses / hidden-layer gradients and parameter gradients within round-off of iteration 0 (dA is summed by L2 in arrival
    order: 1e-6), and layer 0's mu gradient must match the oracle."""
    def case(B, S, hidden, E):
        rng = np.random.default_rng(B + E)
        torch.manual_seed(B)
        skill, member = rand_csr(rng, B, S, 1, min(S, 6)), rand_csr(rng, B, E, 1, min(E - 1, 4))
        layers = O.init_flipout_params(S, hidden, E)
        return dict(B=B, S=S, hidden=hidden, E=E, skill=skill, member=member, layers=layers, noise=O.draw_flipout_noise(layers, B), neg=rng.integers(0, E, (B, 5)))

    def run(c):
        eng = make_engine(c['S'], c['hidden'], c['E'], c['B'], c['skill'], c['member'], c['layers'], precision='tf32')
        sp = eng.split(np.arange(c['B']))
        eng.step(sp, 0, c['B'], True, lr=1e-3, loss_slot=0, neg_host=c['neg'], noise_host=c['noise'])
        torch.cuda.synchronize()
        return eng, {'loss': eng.loss_buf[:1].clone(), 'dz0': eng.dz[0][:c['B']].clone(), 'grads': eng.grads.clone(), 'gdelta': eng.gdelta.clone()}

    prev, main = case(130, 27, [128], 1000), case(256, 40, [128], 300)
    X, y = dense(main['skill']), dense(main['member'])
    logits, acts, pre = O.flipout_forward(main['layers'], main['noise'], X)
    w = O.loss_weights(y, torch.as_tensor(main['neg']), 10, 1)
    g_ref = O.flipout_backward(main['layers'], main['noise'], acts, pre, y, w)
    nrm = lambda a, b: ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
    ref, keep = None, []
    for it in range(iters):
        keep.append(torch.empty((1 + (7 * it) % 5) * 300 * 1024, dtype=torch.uint8, device=DEV))  # shift the next engine's addresses
        if len(keep) > 4: keep.pop(0)
        run(prev)
        eng, out = run(main)
        if ref is None: ref = out
        for k in out:
            assert nrm(out[k].float(), ref[k].float()) < 2e-6, (it, k, nrm(out[k].float(), ref[k].float()))
        assert nrm(grads_of(eng, 0)['mu_w'], g_ref[0]['mu_w']) < 5e-3, it
        del eng


@pytest.mark.parametrize('B,d,hidden,E', [(40, 16, [24], 301), (64, 128, [16, 8], 100)])
def test_flipout_step_on_dense_skill_input_matches_oracle(B, d, hidden, E):
    """Bnn on embedded skills (main.py:148-153, ntf.py:24; the reference's CI runs mdl.bnn.Bnn with d2v / gnn vectors): layer 0 is a dense
    Flipout layer -- forward, loss (+ KL/B) and every gradient against the oracle, host-drawn noise on both sides (fp32 mode: round-off only)."""
    from opentf_b200.engine import Engine
    rng = np.random.default_rng(B + E)
    torch.manual_seed(B)
    X = torch.from_numpy(rng.standard_normal((B, d)).astype(np.float32))
    member = rand_csr(rng, B, E, 1, min(E - 1, 4))
    layers = O.init_flipout_params(d, hidden, E)
    noise = O.draw_flipout_noise(layers, B)
    neg = rng.integers(0, E, (B, 5))
    y = dense(member)
    logits, acts, pre = O.flipout_forward(layers, noise, X)
    w = O.loss_weights(y, torch.as_tensor(neg), 10, 1)
    loss_ref = (O.bce_with_logits(logits, y, w).sum(1).mean() + O.flipout_kl(layers) / B).item()
    g_ref = O.flipout_backward(layers, noise, acts, pre, y, w)
    eng = Engine(d, hidden, E, DEV, bayesian=True, precision='fp32', tpw=10, tnw=1, nsd='uniform', ns=5, seed=3, max_batch=B, dense_input=True)
    eng.stage(X.numpy(), member)
    sd = {}
    for i, L in enumerate(layers):
        sd[f'layers.{i}.mu_weight'], sd[f'layers.{i}.rho_weight'] = L['mu_w'], L['rho_w']
        sd[f'layers.{i}.mu_bias'], sd[f'layers.{i}.rho_bias'] = L['mu_b'], L['rho_b']
    eng.load_state_dict(sd)
    sp = eng.split(np.arange(B))
    eng.step(sp, 0, B, True, lr=1e-3, loss_slot=0, neg_host=neg, noise_host=noise)
    torch.cuda.synchronize()
    loss = eng.loss_buf[0].item()
    assert abs(loss - loss_ref) <= 2e-5 * abs(loss_ref), (loss, loss_ref)
    for i in range(len(layers)):
        g = lambda k, w_: eng.view(f'layers.{i}.{k}_{w_}', eng.grads).cpu()  # (dense input: layer 0 is stored in torch layout too)
        mine = dict(mu_w=g('mu', 'weight'), rho_w=g('rho', 'weight'), mu_b=g('mu', 'bias'), rho_b=g('rho', 'bias'))
        for k in ('mu_w', 'mu_b', 'rho_w', 'rho_b'):
            assert rel_err(mine[k], g_ref[i][k]) < 3e-5, (i, k, rel_err(mine[k], g_ref[i][k]))
    # a second, graph-less step with device-drawn noise runs (counter RNG: signs for the [B,d] input plane)
    eng.step(sp, 0, B, True, lr=1e-3, loss_slot=1)
    torch.cuda.synchronize()
    assert np.isfinite(eng.loss_buf[1].item())
