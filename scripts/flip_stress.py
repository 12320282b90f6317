"""stress / self-consistency loop for the intermittent failure of tests/test_gpu_bnn.py::test_flipout_step_tensor_core_matches_oracle
[256-40-hidden1-300] (known issue): the same Flipout tensor-core step on a FRESH engine every iteration (new
allocations, as in the test), after a step of the preceding test case's shape; every buffer of the step is compared with iteration 0's
(identical inputs: anything beyond round-off is the bug) and the parameter gradients with the oracle.   usage: python scripts/flip_stress.py [iters]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, torch
from oracle import fnn_oracle as O
from test_gpu_kernels import dense, rand_csr
from test_gpu_bnn import make_engine, grads_of

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 25


def case(B, S, hidden, E):
    rng = np.random.default_rng(B + E)
    torch.manual_seed(B)
    skill, member = rand_csr(rng, B, S, 1, min(S, 6)), rand_csr(rng, B, E, 1, min(E - 1, 4))
    layers = O.init_flipout_params(S, hidden, E)
    noise = O.draw_flipout_noise(layers, B)
    neg = rng.integers(0, E, (B, 5))
    return dict(B=B, S=S, hidden=hidden, E=E, skill=skill, member=member, layers=layers, noise=noise, neg=neg)


def run(c):
    eng = make_engine(c['S'], c['hidden'], c['E'], c['B'], c['skill'], c['member'], c['layers'], precision='tf32')
    sp = eng.split(np.arange(c['B']))
    eng.step(sp, 0, c['B'], True, lr=1e-3, loss_slot=0, neg_host=c['neg'], noise_host=c['noise'])
    torch.cuda.synchronize()
    out = {'loss': eng.loss_buf[:1].clone(), 'dact0 (dA + s*dA_s)': eng.dact[0][:c['B']].clone(), 'dact_s0 (dA_s)': eng.dact_s[0][:c['B']].clone(),
           'dz0': eng.dz[0][:c['B']].clone(), 'grads': eng.grads.clone(), 'gdelta': eng.gdelta.clone()}
    return eng, out


prev, main = case(130, 27, [128], 1000), case(256, 40, [128], 300)
X, y = dense(main['skill']), dense(main['member'])
logits, acts, pre = O.flipout_forward(main['layers'], main['noise'], X)
w = O.loss_weights(y, torch.as_tensor(main['neg']), 10, 1)
g_ref = O.flipout_backward(main['layers'], main['noise'], acts, pre, y, w)
nrm = lambda a, b: ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
def poison(word):
    """fill what the caching allocator will hand out next (large and small pool) with a bit pattern: -1 = NaN floats / all sign bits set"""
    big = torch.empty(256 * 1024 * 1024, dtype=torch.int32, device='cuda'); big.fill_(word)
    small = [torch.empty(64 * 1024, dtype=torch.int32, device='cuda').fill_(word) for _ in range(512)]
    torch.cuda.synchronize()
    del big, small


mode = sys.argv[2] if len(sys.argv) > 2 else 'none'  # none | zero | nan: what freshly allocated (torch.empty) buffers contain
ref, bad = None, 0
for it in range(iters):
    if mode != 'none': poison(0 if mode == 'zero' else -1)
    run(prev)
    eng, out = run(main)
    if ref is None: ref = out
    dev = {k: (float('nan') if torch.isnan(out[k]).any() else round(nrm(out[k].float(), ref[k].float()), 7)) for k in out}
    mine = grads_of(eng, 0)
    e0 = round(nrm(mine['mu_w'], g_ref[0]['mu_w']), 5)
    flag = any((v != v) or v > 1e-4 for v in dev.values()) or e0 > 5e-3
    bad += flag
    if flag or it == 0: print(f'iter {it}: layer0 mu_w vs oracle {e0}; vs iteration 0: {dev}', flush=True)
    del eng
print(f'mode {mode}: {bad} of {iters} iterations off')
