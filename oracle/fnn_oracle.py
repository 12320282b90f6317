"""CPU oracle for the OpeNTF Fnn/Bnn hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A from-scratch restatement, in explicit fp32 tensor algebra on the CPU (torch CPU / numpy), of what
`/root/reference/src/mdl/fnn.py`, `bnn.py`, `ntf.py`, `earlystopping.py`, `pkgmgr.py:125-134` and
`evl/metric.py:5-35` compute on the skill->expert train / test / rank path.  Every function cites the
reference lines it follows.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package; `opentf_b200/` never does.

Pinning: `tests/test_oracle_golden.py` checks this file against
  G1  the reference's committed checkpoint -> prediction pairs (43 pairs, 4 toy datasets),
  G2  a seed-0 trajectory recorded from the UNMODIFIED reference (`tests/golden/make_golden.py`),
  G3  the committed per-instance / mean evaluation CSVs (metrics restated from trec_eval definitions),
and, in this container only, directly against the imported reference (`oracle/ref_shim.py`).
Bnn/Flipout follows bayesian-torch 0.5.0, which is NOT vendored in the reference tree: that part is
"parity unpinned" beyond layout + statistics (G4) and is labelled so in DESIGN.md.

Gradients are written out by hand (no autograd) so that the oracle states the arithmetic the CUDA
kernels must reproduce; `tests/test_oracle_golden.py` cross-checks them against autograd.
"""
import math
import numpy as np
import scipy.sparse as sp
import torch

LRELU_SLOPE = 0.01  # torch.nn.functional.leaky_relu default, fnn.py:25


# --------------------------------------------------------------------------------------------------
# model: fnn.py:15-26
# --------------------------------------------------------------------------------------------------
def lrelu(v):
    return torch.where(v > 0, v, v * LRELU_SLOPE)


def init_params(input_size, hidden, output_size):
    """fnn.py:19-23.  `nn.Linear` default init (kaiming-uniform(a=sqrt 5) weight = U(+-1/sqrt(fan_in)),
    bias U(+-1/sqrt(fan_in))) for every layer in order, THEN xavier_uniform_ over every weight in
    order -- both consume the global torch generator, so the draw order is part of the contract.
    Returns [(W[out,in], b[out]), ...] in torch layout."""
    sizes = [input_size] + list(hidden) + [output_size]
    layers = []
    for fan_in, fan_out in zip(sizes[:-1], sizes[1:]):
        bound = 1.0 / math.sqrt(fan_in)
        W = torch.empty(fan_out, fan_in).uniform_(-bound, bound)
        b = torch.empty(fan_out).uniform_(-bound, bound)
        layers.append([W, b])
    for W, _ in layers:
        a = math.sqrt(6.0 / (W.shape[0] + W.shape[1]))
        W.uniform_(-a, a)
    return [(W, b) for W, b in layers]


def forward(layers, X):
    """fnn.py:24-26: leaky_relu after EVERY layer, the output layer included.
    Returns (logits, [activations a_0..a_L], [pre-activations z_0..z_L])."""
    acts, pre = [X], []
    a = X
    for W, b in layers:
        z = a @ W.t() + b
        a = lrelu(z)
        pre.append(z); acts.append(a)
    return a, acts, pre


def densify(mat, rows=None):
    """ntf.py:23: row -> dense fp32.  `mat` is scipy lil/csr uint8 (team.py:154); ntf.py:24: a dense ndarray of skill
    embeddings (main.py:148-153) is handed through as is (`.float()`)."""
    m = mat if rows is None else mat[rows]
    if isinstance(m, np.ndarray): return torch.as_tensor(np.ascontiguousarray(m, dtype=np.float32))
    return torch.as_tensor(np.asarray(m.todense(), dtype=np.float32))


# --------------------------------------------------------------------------------------------------
# negative sampling: fnn.py:48-76
# --------------------------------------------------------------------------------------------------
def ns_uniform(y, ns):
    """fnn.py:48-56: top-ns of iid U[0,1) keys, positives forced to -1."""
    keys = torch.rand_like(y, dtype=torch.float)
    keys = keys.masked_fill(y != 0, -1.0)
    return torch.topk(keys, k=ns, dim=1).indices


def ns_unigram(y, unigram, ns):
    """fnn.py:58-72: multinomial without replacement over unigram[j]*[y==0]; rows whose mass is 0 fall
    back to uniform over ALL experts (positives included)."""
    w = torch.where(y == 0, unigram.expand(y.shape[0], -1), torch.zeros((), dtype=unigram.dtype))
    dead = w.sum(dim=1) == 0
    w[dead] = 1.0
    return torch.multinomial(w, ns, replacement=False)


def batch_unigram(y):
    """fnn.py:75: per-batch expert frequency, fp32, no smoothing."""
    return y.sum(dim=0) / y.shape[0]


def global_unigram(member):
    """fnn.py:82: frequency over ALL teams (fp64 [1,E], numpy matrix division)."""
    return torch.tensor(np.asarray(member.sum(axis=0) / member.shape[0]))


def sample_negatives(y, nsd, ns, unigram=None):
    if nsd == 'uniform': return ns_uniform(y, ns)
    if nsd == 'unigram': return ns_unigram(y, unigram, ns)
    if nsd == 'unigram_b': return ns_unigram(y, batch_unigram(y), ns)
    return None  # fnn.py:39: falsy nsd -> no mask, no RNG draw


# --------------------------------------------------------------------------------------------------
# loss: fnn.py:32-46,135  and its gradient
# --------------------------------------------------------------------------------------------------
def loss_weights(y, neg_idx, tpw, tnw):
    """fnn.py:33-45: weight = tpw on positives and on the sampled negatives, tnw elsewhere."""
    cond = y == 1
    if neg_idx is not None:
        sel = torch.zeros_like(y, dtype=torch.bool)
        valid = neg_idx >= 0  # (-1 = "no sample": extension used by the product's padded lists)
        rows = torch.arange(y.shape[0]).unsqueeze(1).expand_as(neg_idx)
        sel[rows[valid], neg_idx[valid]] = True
        cond = cond | sel
    return torch.where(cond, float(tpw), float(tnw)).to(torch.float32)


def bce_with_logits(x, y, w):
    """torch's stable form of w*((1-y)*x - log sigmoid(x)) (fnn.py:46)."""
    return w * ((1 - y) * x + torch.clamp(-x, min=0) + torch.log1p(torch.exp(-x.abs())))


def batch_loss(logits, y, neg_idx, tpw, tnw):
    """fnn.py:135: per-team sum over experts, mean over the (actual) batch rows."""
    return bce_with_logits(logits, y, loss_weights(y, neg_idx, tpw, tnw)).sum(dim=1).mean()


def backward(layers, acts, pre, y, w):
    """hand-written gradient of `batch_loss` wrt every (W, b); SURVEY.md 9.2.
    d loss / d logits = w*(sigmoid(x)-y)/B ; through each lrelu: *1 if z>0 else *0.01."""
    B = y.shape[0]
    g = w * (torch.sigmoid(acts[-1]) - y) / B
    grads = [None] * len(layers)
    for i in range(len(layers) - 1, -1, -1):
        dz = g * torch.where(pre[i] > 0, 1.0, LRELU_SLOPE)
        grads[i] = (dz.t() @ acts[i], dz.sum(dim=0))
        if i: g = dz @ layers[i][0]
    return grads


# --------------------------------------------------------------------------------------------------
# optimiser + epoch control: fnn.py:104-110,153-169 ; earlystopping.py:26-39
# --------------------------------------------------------------------------------------------------
class Adam:
    """torch.optim.Adam(lr, betas=(.9,.999), eps=1e-8, weight_decay=0, amsgrad=False), fnn.py:104,
    with the operation order of torch's single-tensor CPU path (lerp / addcmul / addcdiv)."""

    def __init__(self, tensors, lr, b1=0.9, b2=0.999, eps=1e-8):
        self.p, self.lr, self.b1, self.b2, self.eps, self.t = tensors, lr, b1, b2, eps, 0
        self.m = [torch.zeros_like(t) for t in tensors]
        self.v = [torch.zeros_like(t) for t in tensors]

    def step(self, grads):
        self.t += 1
        bc1 = 1 - self.b1 ** self.t
        bc2 = 1 - self.b2 ** self.t
        step_size = self.lr / bc1
        bc2_sqrt = math.sqrt(bc2)
        for p, g, m, v in zip(self.p, grads, self.m, self.v):
            m.lerp_(g, 1 - self.b1)
            v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            denom = (v.sqrt() / bc2_sqrt).add_(self.eps)
            p.addcdiv_(m, denom, value=-step_size)


class Plateau:
    """ReduceLROnPlateau(mode='min', factor=.1, patience=2, threshold=1e-4 'rel', cooldown=0), fnn.py:105,163."""

    def __init__(self, factor=0.1, patience=2, threshold=1e-4):
        self.factor, self.patience, self.threshold = factor, patience, threshold
        self.best, self.bad = math.inf, 0

    def step(self, metric, lr):
        if metric < self.best * (1 - self.threshold): self.best, self.bad = metric, 0
        else: self.bad += 1
        if self.bad > self.patience:
            self.bad = 0
            new = lr * self.factor
            if lr - new > 1e-8: return new  # torch's `eps` guard
        return lr


class EarlyStop:
    """earlystopping.py:26-39 with delta=lr (fnn.py:110): an epoch counts as an improvement only if
    -v_loss >= best + delta."""

    def __init__(self, patience, delta):
        self.patience, self.delta, self.best, self.count, self.stop = patience, delta, None, 0, False

    def __call__(self, v_loss):
        score = -v_loss
        if self.best is None: self.best = score
        elif score < self.best + self.delta:
            self.count += 1
            if self.count >= self.patience: self.stop = True
        else: self.best, self.count = score, 0
        return self.stop


def loader_order(n, shuffle):
    """index order of one pass of `DataLoader(dataset, batch_size, shuffle)` (fnn.py:95-96,118) under
    torch 2.x: creating the iterator draws a base seed from the global generator; a RandomSampler then
    draws its own seed and permutes with a private generator."""
    torch.empty((), dtype=torch.int64).random_()  # _BaseDataLoaderIter._base_seed
    if not shuffle: return list(range(n))
    seed = int(torch.empty((), dtype=torch.int64).random_().item())
    g = torch.Generator(); g.manual_seed(seed)
    return torch.randperm(n, generator=g).tolist()


def checkpoint_epochs_rule(spe, e):
    """fnn.py:158."""
    return bool(spe) and (e == 0 or ((e + 1) % spe) == 0)


def learn_fold(skill, member, train_rows, valid_rows, cfg, trace=None, feed=None, unigram=None):
    """One fold of fnn.py:86-168 (Fnn, non-Bayesian) with the reference's RNG consumption order when
    `feed` is None.  `feed` = iterator of recorded (phase, rows, neg_idx) tuples replaces every random
    draw (that is how the CUDA path is replayed against a recorded trajectory).
    `trace` (list) receives (epoch, phase, rows, neg_idx, loss) per step.
    Returns dict(layers, e, t_loss, v_loss, ckpt_epochs, history)."""
    S, E = skill.shape[1], member.shape[1]
    layers = cfg.get('init_layers') or init_params(S, cfg['h'], E)
    # fnn.py:99-101: init() has drawn fresh weights (the generator advanced), THEN prev_model's checkpoint overwrites them (tNtf's warm start)
    if cfg.get('prev_layers'): layers = cfg['prev_layers']
    layers = [(W.clone(), b.clone()) for W, b in layers]
    flat = [t for Wb in layers for t in Wb]
    opt, sched, es = Adam(flat, cfg['lr']), Plateau(), EarlyStop(cfg['es'], cfg['lr'])
    sets = {'train': np.asarray(train_rows), 'valid': np.asarray(valid_rows)}
    ckpts, history = [], []
    for e in range(cfg['e']):
        sums = {'train': 0.0, 'valid': 0.0}; nb = {'train': 0, 'valid': 0}
        for phase in ('train', 'valid'):
            rows_all = sets[phase]
            order = None if feed is not None else loader_order(len(rows_all), phase == 'train')
            nbatch = -(-len(rows_all) // cfg['b'])
            for bi in range(nbatch):
                if feed is not None:
                    ph, rows, neg = next(feed)
                    assert ph == phase
                    rows = np.asarray(rows); neg = None if neg is None else torch.as_tensor(neg)
                else:
                    rows = rows_all[order[bi * cfg['b']:(bi + 1) * cfg['b']]]
                X, y = densify(skill, rows), densify(member, rows)
                logits, acts, pre = forward(layers, X)
                if feed is None: neg = sample_negatives(y, cfg['nsd'], cfg['ns'], unigram)
                w = loss_weights(y, neg, cfg['tpw'], cfg['tnw'])
                loss = bce_with_logits(logits, y, w).sum(dim=1).mean()
                if phase == 'train':
                    grads = backward(layers, acts, pre, y, w)
                    opt.step([g for Wb in grads for g in Wb])
                sums[phase] += loss.item(); nb[phase] += 1
                if trace is not None: trace.append((e, phase, rows.copy(), None if neg is None else neg.clone(), loss.item()))
        t_loss, v_loss = sums['train'] / nb['train'], sums['valid'] / nb['valid']
        history.append((t_loss, v_loss))
        if checkpoint_epochs_rule(cfg.get('spe'), e): ckpts.append((e, [(W.clone(), b.clone()) for W, b in layers]))
        opt.lr = sched.step(v_loss, opt.lr)
        if es(v_loss): break
    return dict(layers=layers, e=e, t_loss=t_loss, v_loss=v_loss, ckpts=ckpts, history=history)


# --------------------------------------------------------------------------------------------------
# test-time output: fnn.py:200-218 ; pkgmgr.py:125-134
# --------------------------------------------------------------------------------------------------
def predict(layers, skill, rows, b):
    """fnn.py:200-213 (Fnn): sigmoid of the logits, batch by batch, stacked on the host."""
    out = []
    for i in range(0, len(rows), b):
        out.append(torch.sigmoid(forward(layers, densify(skill, rows[i:i + b]))[0]))
    return torch.vstack(out)


def topk_sparse(probs, k):
    """pkgmgr.py:125-134: keep the k largest per row -> coalesced COO (sorted by row then column)."""
    v, i = torch.topk(probs, k, dim=1)
    r = torch.arange(probs.shape[0]).unsqueeze(1).expand(-1, k)
    return torch.sparse_coo_tensor(torch.stack([r, i], 0).reshape(2, -1), v.reshape(-1), size=probs.shape).coalesce()


def topk_rows(probs, k):
    """(values, indices) per row in rank order, ties -> lower column first (the product's documented
    tie-break; torch.topk / argpartition leave it unspecified)."""
    p = probs.numpy() if isinstance(probs, torch.Tensor) else np.asarray(probs)
    idx = np.lexsort((np.broadcast_to(np.arange(p.shape[1]), p.shape), -p), axis=1)[:, :k]
    return np.take_along_axis(p, idx, axis=1), idx


# --------------------------------------------------------------------------------------------------
# Bnn / Flipout: bayesian-torch 0.5.0 (un-vendored; restated from the published algorithm)
# --------------------------------------------------------------------------------------------------
def softplus(r):
    return torch.log1p(torch.exp(r))


def init_flipout_params(input_size, hidden, output_size, mu_init=0.0, rho_init=-3.0):
    """bnn.py:19-25 -> dnn_to_bnn -> LinearFlipout.init_parameters: mu ~ N(mu_init, .1), rho ~ N(rho_init, .1)
    per tensor in the order mu_weight, rho_weight, mu_bias, rho_bias.  (The Fnn init is drawn first and
    discarded, fnn.py:29 via bnn.py:25, so it still consumes the generator.)"""
    init_params(input_size, hidden, output_size)
    sizes = [input_size] + list(hidden) + [output_size]
    out = []
    for fan_in, fan_out in zip(sizes[:-1], sizes[1:]):
        mw = torch.empty(fan_out, fan_in).normal_(mu_init, 0.1)
        rw = torch.empty(fan_out, fan_in).normal_(rho_init, 0.1)
        mb = torch.empty(fan_out).normal_(mu_init, 0.1)
        rb = torch.empty(fan_out).normal_(rho_init, 0.1)
        out.append(dict(mu_w=mw, rho_w=rw, mu_b=mb, rho_b=rb))
    return out


def draw_flipout_noise(layers, B):
    """LinearFlipout.forward draw order per layer: eps_weight.normal_, eps_bias.normal_,
    sign_input.uniform_(-1,1).sign(), sign_output.uniform_(-1,1).sign()."""
    noise = []
    for L in layers:
        out_f, in_f = L['mu_w'].shape
        ew = torch.empty(out_f, in_f).normal_()
        eb = torch.empty(out_f).normal_()
        si = torch.empty(B, in_f).uniform_(-1, 1).sign()
        so = torch.empty(B, out_f).uniform_(-1, 1).sign()
        noise.append(dict(eps_w=ew, eps_b=eb, s_in=si, s_out=so))
    return noise


def flipout_forward(layers, noise, X):
    """out = a mu_W^T + mu_b + ((a*s_in) dW^T + db) * s_out, then lrelu (every layer)."""
    acts, pre, a = [X], [], X
    for L, Nz in zip(layers, noise):
        dW = softplus(L['rho_w']) * Nz['eps_w']
        db = softplus(L['rho_b']) * Nz['eps_b']
        z = a @ L['mu_w'].t() + L['mu_b'] + ((a * Nz['s_in']) @ dW.t() + db) * Nz['s_out']
        a = lrelu(z); pre.append(z); acts.append(a)
    return a, acts, pre


def kl_mean(mu, sigma, prior_mu=0.0, prior_sigma=1.0):
    """base_variational_layer.kl_div: MEAN over the tensor of
    log(sp) - log(s) + (s^2 + (mu-mp)^2)/(2 sp^2) - 1/2."""
    return (math.log(prior_sigma) - torch.log(sigma) + (sigma ** 2 + (mu - prior_mu) ** 2) / (2 * prior_sigma ** 2) - 0.5).mean()


def flipout_kl(layers):
    """get_kl_loss: sum over layers of KL(weight) + KL(bias), each a mean."""
    return sum(kl_mean(L['mu_w'], softplus(L['rho_w'])) + kl_mean(L['mu_b'], softplus(L['rho_b'])) for L in layers)


def flipout_backward(layers, noise, acts, pre, y, w):
    """gradient of  batch_loss + KL/B  wrt mu/rho of every layer."""
    B = y.shape[0]
    g = w * (torch.sigmoid(acts[-1]) - y) / B
    grads = [None] * len(layers)
    for i in range(len(layers) - 1, -1, -1):
        L, Nz, a = layers[i], noise[i], acts[i]
        dz = g * torch.where(pre[i] > 0, 1.0, LRELU_SLOPE)
        dzs = dz * Nz['s_out']
        sw, sb = softplus(L['rho_w']), softplus(L['rho_b'])
        d_dW = dzs.t() @ (a * Nz['s_in'])
        d_db = dzs.sum(dim=0)

        def kl_g(mu, sig):  # d(KL mean)/dmu , d(KL mean)/dsigma ; divided by B as in fnn.py:136
            n = mu.numel() * B
            return mu / n, (-1.0 / sig + sig) / n
        kmw, ksw = kl_g(L['mu_w'], sw); kmb, ksb = kl_g(L['mu_b'], sb)
        grads[i] = dict(
            mu_w=dz.t() @ a + kmw, mu_b=dz.sum(dim=0) + kmb,
            rho_w=(d_dW * Nz['eps_w'] + ksw) * torch.sigmoid(L['rho_w']),
            rho_b=(d_db * Nz['eps_b'] + ksb) * torch.sigmoid(L['rho_b']))
        if i: g = dz @ L['mu_w'] + (dzs @ (sw * Nz['eps_w'])) * Nz['s_in']
    return grads


def predictive_entropy(mc):
    """bayesian_torch.utils.util.predictive_entropy: entropy of the MC-mean prediction, summed over the
    expert axis; `mc` is [nmc, B, E] (fnn.py:205-207)."""
    p = mc.mean(axis=0)
    return -(p * np.log(p + 1e-15)).sum(axis=-1)


def mutual_information(mc):
    """predictive entropy minus the mean per-sample entropy (fnn.py:208)."""
    return predictive_entropy(mc) - (-(mc * np.log(mc + 1e-15)).sum(axis=-1)).mean(axis=0)


# --------------------------------------------------------------------------------------------------
# ranking metrics: evl/metric.py:5-35 (pytrec_eval is not installable here: measures restated from the
# trec_eval definitions; pinned by the committed *.eval.instance.csv / *.eval.mean.csv, G3)
# --------------------------------------------------------------------------------------------------
def trec_metrics(Y, Y_, ks=(2, 5, 10), topK=None):
    """Y: scipy [N,E] 0/1 truth; Y_: dense [N,E] or scipy csr scores.  Returns {metric_k: ndarray[N]} for
    P, recall, ndcg_cut, map_cut, success.  Ranking as trec_eval does: score descending, ties broken
    by docno string DESCENDING (trec_eval sorts docno lexicographically as the secondary key)."""
    N, E = Y.shape
    Yc = sp.csr_matrix(Y)
    out = {f'{m}_{k}': np.zeros(N) for m in ('P', 'recall', 'ndcg_cut', 'map_cut', 'success') for k in ks}
    for i in range(N):
        rel = set(Yc.indices[Yc.indptr[i]:Yc.indptr[i + 1]].tolist())
        if sp.issparse(Y_):
            Yr = sp.csr_matrix(Y_)
            cols, vals = Yr.indices[Yr.indptr[i]:Yr.indptr[i + 1]], Yr.data[Yr.indptr[i]:Yr.indptr[i + 1]]
        else:
            cols, vals = np.arange(E), np.asarray(Y_[i])
        kk = min(topK, len(vals)) if topK else len(vals)
        keep = np.argsort(-vals, kind='stable')[:kk]  # metric.py:19-28 first-stage cut
        docs = sorted(((float(vals[j]), 'd' + str(int(cols[j]))) for j in keep), key=lambda t: (t[0], t[1]), reverse=True)
        hits = np.array([int(d[1:]) in rel for _, d in docs], dtype=np.float64)
        R = len(rel)
        for k in ks:
            h = hits[:k]
            out[f'P_{k}'][i] = h.sum() / k
            out[f'recall_{k}'][i] = h.sum() / R if R else 0.0
            out[f'success_{k}'][i] = float(h.sum() > 0)
            disc = 1.0 / np.log2(np.arange(2, len(h) + 2))
            idcg = (1.0 / np.log2(np.arange(2, min(R, k) + 2))).sum()
            out[f'ndcg_cut_{k}'][i] = (h * disc).sum() / idcg if R else 0.0
            prec = np.cumsum(h) / np.arange(1, len(h) + 1)
            out[f'map_cut_{k}'][i] = (prec * h).sum() / R if R else 0.0
    return out
