"""Flipout tensor-core step: WHICH operand of MODE 0's logits is wrong in a bad iteration?  Builds on flip_stress2.py (same hand-run step on
a fresh engine per iteration) and adds the kernel's debug taps: the TMEM accumulator as read (NTF_TC_ZRAW) and the perturbation-term values as
read from shared memory (NTF_TC_QDBG); the latter is compared with the Q tiles in HBM (what the bulk copy should have delivered).
usage: NTF_TC_EXP=<bits> python scripts/flip_stress3.py [iters]     bits: 16 = proxy fence before the EMPTY arrivals, 32 = poison shared memory"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, torch
from oracle import fnn_oracle as O
from opentf_b200 import ops
from opentf_b200._lib import OutTrainArgs
from test_gpu_kernels import rand_csr
from test_gpu_bnn import make_engine

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 12
B, S, hidden, E = (int(os.environ.get('FS_B', 256)), 40, [128], int(os.environ.get('FS_E', 300)))
TE = TB = 128
rng = np.random.default_rng(B + E)
torch.manual_seed(B)
skill, member = rand_csr(rng, B, S, 1, min(S, 6)), rand_csr(rng, B, E, 1, min(E - 1, 4))
layers = O.init_flipout_params(S, hidden, E)
noise = O.draw_flipout_noise(layers, B)
neg_host = rng.integers(0, E, (B, 5))
up = lambda v, a: (v + a - 1) // a * a
Epad, nt, h = up(E, TE), up(B, TB) // TB, 128
base = up(B * h * 2, 256) + up((up(E, TE) // TE + 1024) * 4, 1024)
off_as16, off_sign = base, base + up(B * h * 2, 1024)
off_q = off_sign + up(nt * Epad * 4 * 4, 1024)
off_dzs = off_q + up(nt * Epad * TB * 2, 1024)
end = off_dzs + up(nt * Epad * TB * 2, 1024)
zraw = torch.zeros(B, E, device='cuda'); qdbg = torch.zeros(B, E, device='cuda')
os.environ['NTF_TC_ZRAW'] = str(zraw.data_ptr()); os.environ['NTF_TC_QDBG'] = str(qdbg.data_ptr())


def q_from_tiles(qt):
    """[tile][etile][16 units][128 experts][8] -> [B, E]"""
    q = qt.view(nt, Epad // TE, 16, TE, 8).permute(0, 2, 4, 1, 3).reshape(nt * TB, Epad)
    return q[:B, :E]


def run():
    eng = make_engine(S, hidden, E, B, skill, member, layers, precision='tf32')
    sp = eng.split(np.arange(B))
    Lo = eng.L - 1
    eng._draw_noise(sp, 0, B, noise)
    eng._prepare_bayes()
    eng._forward_hidden_bayes(sp, 0, B)
    neg = eng._sample(sp, 0, B, neg_host, None)
    mptr = sp.m_indptr.data_ptr()
    ops.special_tiles(1, B, mptr, sp.m_indices, neg, neg.shape[1], eng.E, eng.special_t, eng.member_t)
    zraw.fill_(float('nan')); qdbg.fill_(float('nan'))
    torch.cuda.synchronize()
    a = OutTrainArgs()
    a.A, a.W, a.b = eng.act[-1].data_ptr(), eng._pv(Lo, 'mu', 'weight').data_ptr(), eng._pv(Lo, 'mu', 'bias').data_ptr()
    a.pitch_words = eng.pitch
    a.special_t, a.member_t = eng.special_t.data_ptr(), eng.member_t.data_ptr()
    a.m_indptr, a.m_indices = mptr, sp.m_indices.data_ptr()
    a.B, a.h, a.E = B, h, eng.E
    a.tpw, a.tnw, a.loss_scale = eng.tpw, eng.tnw, 1.0 / B
    a.loss_out = eng.loss_buf.data_ptr()
    a.A_s, a.W_delta, a.b_delta = eng.act_s[-1].data_ptr(), eng.nview(eng.delta, f'{Lo}.weight').data_ptr(), eng.nview(eng.delta, f'{Lo}.bias').data_ptr()
    a.sign_out = eng.sign_out[Lo].data_ptr()
    a.dW, a.db = eng._pv(Lo, 'mu', 'weight', eng.grads).data_ptr(), eng._pv(Lo, 'mu', 'bias', eng.grads).data_ptr()
    a.dA = eng.dact[-1].data_ptr()
    a.dW_delta, a.db_delta = eng.nview(eng.gdelta, f'{Lo}.weight').data_ptr(), eng.nview(eng.gdelta, f'{Lo}.bias').data_ptr()
    a.dA_s = eng.dact_s[-1].data_ptr()
    ops.out_train(eng.dev_index, eng.precision, a, eng.ws)
    torch.cuda.synchronize()
    w = eng.ws.buf
    f16 = lambda lo, hi: w[lo:hi].clone().view(torch.float16).float()
    qh = q_from_tiles(f16(off_q, off_q + nt * Epad * TB * 2))
    return {'Q hbm': qh, 'Q read': qdbg.clone(), 'Z raw': zraw.clone(), 'loss': eng.loss_buf[:1].clone(), 'dW': eng._pv(Lo, 'mu', 'weight', eng.grads).clone(),
            'DZS': f16(off_dzs, end), 'A16': w[:B * h * 2].clone().view(torch.float16).float()}


def shuffle_addresses(i):
    junk = [torch.empty((1 + (7 * i + k) % 5) * 300 * 1024, dtype=torch.uint8, device='cuda') for k in range(6)]
    keep = junk[::2]
    del junk
    return keep


def where(mask):
    idx = mask.nonzero()
    n, e = idx[:, 0].cpu().numpy(), idx[:, 1].cpu().numpy()
    return f'{len(n)} entries; team tiles {sorted(set((n // TB).tolist()))}, expert tiles {sorted(set((e // TE).tolist()))}, 32-team blocks {sorted(set(((n % TB) // 32).tolist()))}, experts%128 range [{(e % TE).min()}, {(e % TE).max()}]'


ref, bad, keep = None, 0, []
for it in range(iters):
    keep.append(shuffle_addresses(it))
    out = run()
    if ref is None: ref = out
    msgs = []
    qbad = ~(out['Q read'] == out['Q hbm'])  # (NaN counts as bad)
    if qbad.any():
        msgs.append('Q read != Q hbm: ' + where(qbad))
        sub = out['Q read'][qbad]
        msgs.append(f'   of those NaN: {int(torch.isnan(sub).sum())}')
        if nt >= 2:  # does the stale value equal another tile's Q at the same (team % 128, expert)?
            qh = out['Q hbm']
            for sh in range(1, nt):
                other = torch.roll(qh, -sh * TB, 0) if B % TB == 0 else None
                if other is not None: msgs.append(f'   equal to tile+{sh}\'s Q: {int((out["Q read"][qbad] == other[qbad]).sum())}')
    zbad = ~(out['Z raw'] == ref['Z raw'])
    if zbad.any(): msgs.append('Z raw differs from iteration 0: ' + where(zbad) + f'; max |d| {float((out["Z raw"] - ref["Z raw"])[zbad].abs().max()):.4g}; NaN {int(torch.isnan(out["Z raw"][zbad]).sum())}')
    for k in ('loss', 'dW', 'DZS', 'A16', 'Q hbm'):
        if not torch.equal(out[k], ref[k]) : msgs.append(f'{k} differs from iteration 0' + (' (NaN)' if torch.isnan(out[k]).any() else ''))
    bad += bool(msgs)
    if msgs: print(f'iter {it}:\n  ' + '\n  '.join(msgs), flush=True)
print(f'exp={os.environ.get("NTF_TC_EXP", "0")} B={B} E={E}: {bad} of {iters} iterations bad', flush=True)
