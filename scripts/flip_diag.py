"""diagnostic: per-tensor errors of the tensor-core Flipout step against the oracle (prints, does not assert).  usage: python scripts/flip_diag.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, torch
from oracle import fnn_oracle as O
from test_gpu_kernels import dense, rand_csr
from test_gpu_bnn import make_engine, grads_of

for B, S, hidden, E in [(130, 27, [128], 1000), (256, 40, [128], 300), (1000, 27, [128], 20000), (77, 30, [16, 128], 129)]:
    rng = np.random.default_rng(B + E); torch.manual_seed(B)
    skill, member = rand_csr(rng, B, S, 1, min(S, 6)), rand_csr(rng, B, E, 1, min(E - 1, 4))
    layers = O.init_flipout_params(S, hidden, E)
    noise = O.draw_flipout_noise(layers, B)
    neg = rng.integers(0, E, (B, 5))
    X, y = dense(skill), dense(member)
    logits, acts, pre = O.flipout_forward(layers, noise, X)
    w = O.loss_weights(y, torch.as_tensor(neg), 10, 1)
    loss_ref = (O.bce_with_logits(logits, y, w).sum(1).mean() + O.flipout_kl(layers) / B).item()
    g_ref = O.flipout_backward(layers, noise, acts, pre, y, w)
    nrm = lambda a, b: ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
    for prec in ('fp32', 'tf32'):
        eng = make_engine(S, hidden, E, B, skill, member, layers, precision=prec)
        sp = eng.split(np.arange(B))
        eng.step(sp, 0, B, True, lr=1e-3, loss_slot=0, neg_host=neg, noise_host=noise)
        torch.cuda.synchronize()
        loss = eng.loss_buf[0].item()
        errs = {f'{i}.{k}': round(nrm(grads_of(eng, i)[k], g_ref[i][k]), 6) for i in range(len(layers)) for k in ('mu_w', 'mu_b', 'rho_w', 'rho_b')}
        print(B, S, hidden, E, prec, 'loss rel', abs(loss - loss_ref) / abs(loss_ref), errs, flush=True)
    near = ((pre[-1].abs() < 2e-3).float().mean().item())
    print('   fraction of output pre-activations within 2e-3 of the kink:', near)
