"""`Fnn`: OpeNTF's feed-forward skill->expert recommender (reference: src/mdl/fnn.py) on B200 kernels.

Drop-in for the reference class: same `init / learn / test` entry points and cfg knobs (`b e ns lr es h spe l tpw tnw
nsd`), same `f{fold}.pt`, `f{fold}.e{N}.pt`, `f{fold}.{set}.[e{N}.]pred` files.  What differs is where the work
happens: teamsvecs stay sparse on the GPU for the whole run, a step is ~10 kernel launches of libntf_b200.so,
and the host sees one loss vector per epoch instead of a `.item()` per step (fnn.py:140).

Host-side RNG is consumed in the reference's order (layer init, then per epoch the DataLoader base seed, the
shuffle seed and `randperm`), so with `nsd` unset a run retraces the reference batch for batch.  With
negative sampling on, the reference's draws (B x E random keys per step from torch's generator, fnn.py:51,71)
are replaced by the documented counter RNG of `ntf_neg_sample`; tests feed recorded indices instead
(`neg_provider`) to compare trajectories.
"""
import logging
import os
import re
import time

import numpy as np

from . import util
from .earlystopping import EarlyStopping
from .engine import Engine
from .ntf import Ntf

log = logging.getLogger(__name__)


def loader_order(torch, n, shuffle):
    """index order of one pass of DataLoader(dataset, batch_size, shuffle) (fnn.py:95-96,118) as torch 2.x draws it:
    the iterator takes a base seed from the global generator; a shuffling sampler then takes its own seed and
    permutes with a private generator."""
    torch.empty((), dtype=torch.int64).random_()
    if not shuffle: return None
    seed = int(torch.empty((), dtype=torch.int64).random_().item())
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g).numpy()


class _Plateau:
    """ReduceLROnPlateau(mode='min', factor=0.1, patience=2, threshold=1e-4 relative, cooldown=0, eps=1e-8), fnn.py:105,163."""

    def __init__(self, factor=0.1, patience=2, threshold=1e-4, eps=1e-8):
        self.factor, self.patience, self.threshold, self.eps = factor, patience, threshold, eps
        self.best, self.num_bad = float('inf'), 0

    def step(self, value, lr):
        if value < self.best * (1.0 - self.threshold): self.best, self.num_bad = value, 0
        else: self.num_bad += 1
        if self.num_bad > self.patience:
            self.num_bad = 0
            new_lr = lr * self.factor
            if lr - new_lr > self.eps:
                log.info(f'Reducing learning rate to {new_lr:.4e}')
                return new_lr
        return lr


class DeviceModel:
    """what `self.model` is on this path: parameters live in the engine's arena; the torch-facing surface is the
    state dict (keys/layout of the reference's nn.Module: layers.{i}.weight [out,in], layers.{i}.bias)."""

    def __init__(self, engine): self.engine = engine
    def state_dict(self): return self.engine.state_dict()
    def load_state_dict(self, sd): self.engine.load_state_dict(sd); return self
    def to(self, device): return self
    def train(self, mode=True): return self
    def eval(self): return self


class Fnn(Ntf):
    precision_default = 'tf32'

    def __init__(self, output, device, seed, cfg):
        super().__init__(output, device, seed, cfg)
        self.engine = None
        # parity tests replay a recorded reference run through this class: an object with
        #   init(fold) -> state dict | None, order(fold, epoch, phase, n) -> positions | None, neg(fold, epoch, phase, batch, team_rows) -> [B,ns] | None
        # replaces the host RNG draws (initial weights, shuffles) and supplies the negatives (SURVEY.md 9.3 "host-supplied indices").
        self.replay = None
        self.last_history = {}

    # ---- helpers --------------------------------------------------------------------------------------------
    def _c(self, key, default=None): return util.cfg_get(self.cfg, key, default)

    def _device(self):
        d = str(self.device)
        if d == 'cpu': raise RuntimeError("opentf_b200 has no CPU path: set acceleration to 'cuda' / 'cuda:N'")
        if ',' in d:  # the reference's multi-GPU spelling 'cuda:0,1,...' (nmt.py:74): one process per GPU picks its own
            ids = d.split(':')[1].split(',')
            return f'cuda:{ids[int(os.environ.get("LOCAL_RANK", 0)) % len(ids)]}'
        return d if ':' in d else 'cuda:0'

    def _host_init(self, input_size, output_size):
        """fnn.py:17-23: build the reference's torch modules on the host so that initial weights (and the generator
        state afterwards) are exactly the reference's."""
        torch = Ntf.torch
        h = list(self._c('h'))
        layers = [torch.nn.Linear(input_size, h[0])]
        for i in range(1, len(h)): layers.append(torch.nn.Linear(h[i - 1], h[i]))
        layers.append(torch.nn.Linear(h[-1], output_size))
        for m in layers: torch.nn.init.xavier_uniform_(m.weight)
        sd = {}
        for i, m in enumerate(layers): sd[f'layers.{i}.weight'], sd[f'layers.{i}.bias'] = m.weight.detach(), m.bias.detach()
        return sd

    def init(self, input_size, output_size):
        dense = bool(getattr(self, '_dense_input', False))  # set by learn/test from teamsvecs['skill'] (ndarray = embedded skills, ntf.py:24)
        if self.engine is None or (self.engine.S, self.engine.E_total, self.engine.dense_input) != (input_size, output_size, dense):
            torch = Ntf.torch
            world, rank = 1, 0
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                world, rank = torch.distributed.get_world_size(), torch.distributed.get_rank()
            # multi-GPU mode (extra knob, default 'dp'): 'dp' = data-parallel teams, 'shard' = output layer column-sharded by expert,
            # 'none' = this process trains alone even inside a process group (what the N-GPU modes are checked against)
            if self._c('parallel', os.environ.get('NTF_PARALLEL', 'dp')) == 'none': world, rank = 1, 0
            shard = (rank, world) if (world > 1 and self._c('parallel', os.environ.get('NTF_PARALLEL', 'dp')) == 'shard') else None
            self.engine = Engine(input_size, list(self._c('h')), output_size, self._device(), bayesian=self.is_bayesian_cls(),
                                 precision=self._c('precision', os.environ.get('NTF_PRECISION', self.precision_default)),
                                 tpw=self._c('tpw', 1), tnw=self._c('tnw', 1), nsd=self._c('nsd'), ns=self._c('ns', 5),
                                 seed=self.seed if self.seed is not None else 0, max_batch=self._c('b'), shard=shard, dense_input=dense)
            self.engine.world, self.engine.rank = world, rank
            # data-parallel ranks: how the gradients are exchanged (extra knob `exchange` / NTF_DP_EXCHANGE):
            #   'peer'  (default) reduce-scatter + Adam + all-gather as one pass over peer memory inside the step (csrc/peer.cu)
            #   'nccl'  ncclAllReduce on the library's own communicator inside the step (Fnn only)
            #   'torch' torch.distributed.all_reduce between the two halves of a step
            if shard is not None and self._c('exchange', os.environ.get('NTF_DP_EXCHANGE', 'peer')) == 'peer': self.engine.attach_shard_peers()
            if world > 1 and shard is None:
                mode = self._c('exchange', os.environ.get('NTF_DP_EXCHANGE', 'peer'))
                if mode == 'peer': self.engine.attach_peers()
                elif mode == 'nccl' and not self.is_bayesian_cls(): self.engine.attach_comm()
        self.engine.load_state_dict(self._host_init(input_size, output_size))
        self.engine.reset_optimizer()
        self.model = DeviceModel(self.engine)
        return self.model

    @classmethod
    def is_bayesian_cls(cls): return False

    def _stage(self, teamsvecs):
        if self.engine.skill is None or getattr(self, '_staged', None) is not teamsvecs:
            self.engine.stage(teamsvecs['skill'], teamsvecs['member'])
            self._staged = teamsvecs

    def _load_ckpt(self, path):
        return Ntf.torch.load(path, map_location='cpu', weights_only=False)['model_state_dict']

    def _rank_slice(self, B):
        """this rank's share [lo, hi) of a global batch of B teams (data parallel: teams are independent)."""
        G, r = self.engine.world, self.engine.rank
        if self.engine.shard[1] > 1: return 0, B  # expert-sharded: every rank runs the whole batch on its expert range
        per = -(-B // G)
        return min(B, r * per), min(B, (r + 1) * per)

    # ---- fnn.py:78-170 ----------------------------------------------------------------------------------------
    def learn(self, teamsvecs, splits, prev_model):
        torch = Ntf.torch
        self._dense_input = scipy_dense(teamsvecs['skill'])  # main.py:148-153: skill embeddings instead of the multi-hot rows
        input_size, output_size = teamsvecs['skill'].shape[1], teamsvecs['member'].shape[1]
        b, nsd = int(self._c('b')), self._c('nsd')
        w = self.writer(log_dir=f'{self.output}/logs4tboard/run_{int(time.time())}')
        first = True
        for foldidx in splits['folds'].keys():
            self.init(input_size, output_size)
            eng = self.engine
            if first:
                self._stage(teamsvecs)
                if nsd == 'unigram': eng.set_global_unigram()  # fnn.py:82
                first = False
            if prev_model: self.model.load_state_dict(self._load_ckpt(prev_model[foldidx]))  # fnn.py:101
            if self.replay is not None and self.replay.init(foldidx) is not None: self.model.load_state_dict(self.replay.init(foldidx))
            train_sp = eng.split(splits['folds'][foldidx]['train'])
            valid_sp = eng.split(splits['folds'][foldidx]['valid'])
            nb_t, nb_v = -(-train_sp.n // b), -(-valid_sp.n // b)
            if nb_t + nb_v > eng.loss_buf.numel(): eng.loss_buf = torch.zeros(nb_t + nb_v, dtype=torch.float32, device=eng.device)
            lr = float(self._c('lr'))
            sched = _Plateau()
            es = EarlyStopping(patience=self._c('es'), delta=self._c('lr'), verbose=True, trace_func=log.info)
            history = []
            for e in range(int(self._c('e'))):
                eng.loss_buf.zero_()
                for phase, sp, nb, slot0 in (('train', train_sp, nb_t, 0), ('valid', valid_sp, nb_v, nb_t)):
                    order = loader_order(torch, sp.n, phase == 'train')
                    if self.replay is not None: order = self.replay.order(foldidx, e, phase, sp.n)
                    if order is not None: sp.regather(order)
                    for bi in range(nb):
                        b0, B = bi * b, min(b, sp.n - bi * b)
                        lo, hi = self._rank_slice(B)
                        if hi <= lo:  # (data parallel, a last batch shorter than the number of ranks)
                            eng.idle_step(phase == 'train', lr, loss_scale=1.0 / B, loss_slot=slot0 + bi)
                            continue
                        neg = None
                        if self.replay is not None:
                            neg = self.replay.neg(foldidx, e, phase, bi, sp.rows_now[b0:b0 + B])
                            if neg is not None: neg = np.asarray(neg)[lo:hi]
                        noise = self.replay.noise(foldidx, e, phase, bi, B) if hasattr(self.replay, 'noise') else None  # Bnn: recorded Flipout draws
                        eng.step(sp, b0 + lo, hi - lo, phase == 'train', lr=lr, loss_slot=slot0 + bi, neg_host=neg, loss_scale=1.0 / B, gbatch=(b0, B),
                                 noise_host=noise)
                losses = eng.loss_buf[:nb_t + nb_v]
                if eng.world > 1: torch.distributed.all_reduce(losses)
                losses = losses.cpu().tolist()  # the one host sync of the epoch
                self._check_peers()
                t_loss = sum(losses[:nb_t]) / nb_t
                v_loss = sum(losses[nb_t:]) / nb_v
                history.append((t_loss, v_loss))
                w.add_scalar(tag=f'{foldidx}_t_loss', scalar_value=t_loss, global_step=e)
                w.add_scalar(tag=f'{foldidx}_v_loss', scalar_value=v_loss, global_step=e)
                log.info(f'Fold {foldidx}/{len(splits["folds"]) - 1}, Epoch {e}, Train Loss: {t_loss:.4f}')
                log.info(f'Fold {foldidx}/{len(splits["folds"]) - 1}, Epoch {e}, Valid Loss: {v_loss:.4f}')
                spe = self._c('spe')
                if spe and (e == 0 or ((e + 1) % spe) == 0): self._save(foldidx, e, t_loss, v_loss, f'{self.output}/f{foldidx}.e{e}.pt')
                lr = sched.step(v_loss, lr)
                if es(v_loss, self.model).early_stop:
                    log.info(f'Early stopping triggered at epoch: {e}')
                    break
            self._save(foldidx, e, t_loss, v_loss, f'{self.output}/f{foldidx}.pt')
            self.last_history[foldidx] = history
        w.close()

    def _check_peers(self):
        """data-parallel ranks exchanging gradients over peer memory: a flag barrier that gave up (a rank stalled past the timeout or died)
        leaves the replicas summed over stale gradients -- stop the run instead of training on / checkpointing corrupt parameters."""
        err = self.engine.peer_error()
        if err: raise RuntimeError(f'peer-memory gradient exchange timed out (flag {err}) on rank {self.engine.rank}: replicas are out of step; '
                                   f'raise NTF_PEER_TIMEOUT_S or use exchange=nccl')

    def _save(self, foldidx, e, t_loss, v_loss, path):
        self._check_peers()
        sd = self.model.state_dict() if (self.engine.shard[1] > 1 or self.engine.rank == 0) else None  # (sharded: a collective gather)
        if self.engine.rank == 0:
            Ntf.torch.save({'model_state_dict': sd, 'cfg': self.cfg, 'f': foldidx, 'e': e, 't_loss': t_loss, 'v_loss': v_loss}, path)
            log.info(f'{self.name()} model with {util.cfg2str(self.cfg)} saved at {path}')
        if self.engine.world > 1: Ntf.torch.distributed.barrier()  # the other ranks read this file next (test(), tNtf's next year)

    # ---- fnn.py:172-219 ---------------------------------------------------------------------------------------
    def test(self, teamsvecs, splits, testcfg):
        torch = Ntf.torch
        assert os.path.isdir(self.output), f'No folder for {self.output} exist!'
        self._dense_input = scipy_dense(teamsvecs['skill'])
        input_size, output_size = teamsvecs['skill'].shape[1], teamsvecs['member'].shape[1]
        b = int(self._c('b'))
        topK = util.cfg_get(testcfg, 'topK')
        test_sp = None
        for foldidx in splits['folds'].keys():
            modelfiles = [f'{self.output}/f{foldidx}.pt']
            if util.cfg_get(testcfg, 'per_epoch'):
                modelfiles += [f'{self.output}/{_}' for _ in os.listdir(self.output) if re.match(rf'f{foldidx}\.e\d+\.pt', _)]
            for modelfile in sorted(sorted(modelfiles), key=len):
                self.init(input_size, output_size)
                self._stage(teamsvecs)
                eng = self.engine
                self.model.load_state_dict(self._load_ckpt(modelfile))
                if test_sp is None or test_sp.eng is not eng: test_sp = eng.split(splits['test'])
                for pred_set in (['test', 'train', 'valid'] if util.cfg_get(testcfg, 'on_train') else ['test']):
                    sp = test_sp if pred_set == 'test' else eng.split(splits['folds'][foldidx][pred_set])
                    sparse = bool(topK) and topK < output_size
                    y_pred, unc = self._predict_split(sp, b, topK if sparse else None)
                    match = re.search(r'(e\d+)\.pt$', os.path.basename(modelfile))
                    epoch = (match.group(1) + '.') if match else ''
                    if eng.rank == 0:
                        torch.save({'y_pred': y_pred, 'uncertainty': unc}, f'{self.output}/f{foldidx}.{pred_set}.{epoch}pred', pickle_protocol=4)
                        log.info(f'{self.name()} model predictions for fold{foldidx}.{pred_set}.{epoch} has saved at {self.output}/f{foldidx}.{pred_set}.{epoch}pred')
        # rank 0 wrote the files; evaluate() (main.py:189) reads them next, on whichever rank runs it
        if self.engine is not None and self.engine.world > 1: torch.distributed.barrier()

    def _gather_columns(self, t): return _gather_columns_impl(Ntf.torch, self.engine, t)

    def _predict_split(self, sp, b, K):
        """fnn.py:198-218 for one prediction set: dense [N,E] probabilities, or -- when topK < E -- the K best per team
        selected on the GPU (only [N,K] leaves it) and stored as the same coalesced sparse COO tensor.
        Bnn: every batch is the mean of `nmc` stochastic passes; `uncertainty` keeps the reference's shape, including its
        quirk that the lists are re-created per batch so only the LAST batch's entropies are saved (fnn.py:203)."""
        torch, eng = Ntf.torch, self.engine
        bb = min(b, max(1, sp.n))
        bayes = eng.bayesian
        fused = K is not None and eng.fused_topk_ok(bb, K)  # output layer + sigmoid + top-K in one call: no [B,E] score buffer at all
        scores = None if fused else torch.empty(bb, eng.E, dtype=torch.float32, device=eng.device)
        if bayes:
            nmc = int(self._c('nmc', 2))
            scratch = torch.empty(bb, eng.E, dtype=torch.float32, device=eng.device)
            ent_pred, ent_model = torch.empty(bb, dtype=torch.float32, device=eng.device), torch.empty(bb, dtype=torch.float32, device=eng.device)
        unc = None
        # data-parallel ranks (replicas): the teams of the set are independent, so batch i is predicted by rank i % world and the
        # pieces are put together afterwards (the expert-sharded layer instead runs every batch on every rank)
        G, r = (eng.world, eng.rank) if eng.shard[1] == 1 else (1, 0)
        nb = -(-sp.n // b)
        out = torch.empty(sp.n, eng.E_total, dtype=torch.float32) if (K is None and eng.rank == 0) else None
        if K is not None:
            vals = torch.zeros(sp.n, K, dtype=torch.float32, device=eng.device)
            idx = torch.zeros(sp.n, K, dtype=torch.int32, device=eng.device)
        for bi in range(nb):
            b0 = bi * b
            B = min(b, sp.n - b0)
            mine = bi % G == r
            if mine:
                if bayes:
                    eng.scores_mc(sp, b0, B, nmc, scores, scratch, ent_pred, ent_model)
                    if bi == nb - 1: unc = {'pred': [ent_pred[:B].cpu().numpy()], 'model': [ent_model[:B].cpu().numpy()]}
                elif fused:
                    eng._forward_hidden(sp, b0, B)
                    eng.select_topk_fused(B, K, vals[b0:b0 + B], idx[b0:b0 + B])
                else:
                    eng.scores(sp, b0, B, scores)
                if K is not None and not fused: eng.select_topk(scores, B, K, vals[b0:b0 + B], idx[b0:b0 + B])
            if K is None:  # batch by batch, as fnn.py:211-212 does (expert-sharded: the column blocks of the ranks side by side)
                if eng.shard[1] > 1:
                    g = self._gather_columns(scores[:B])
                    if out is not None: out[b0:b0 + B] = g
                elif G == 1: out[b0:b0 + B] = scores[:B].cpu()
                elif bi % G == G - 1 or bi == nb - 1:  # a round of G batches is done: every rank hands its batch to rank 0
                    parts = [torch.empty_like(scores) for _ in range(G)] if r == 0 else None
                    torch.distributed.gather(scores, parts, dst=0)
                    if r == 0:
                        for q in range(bi % G + 1):
                            bq = (bi - bi % G + q) * b
                            out[bq:bq + min(b, sp.n - bq)] = parts[q][:min(b, sp.n - bq)].cpu()
        if G > 1 and bayes:  # the reference keeps the LAST batch's entropies (fnn.py:203): they live on the rank that predicted it
            last_B = sp.n - (nb - 1) * b
            buf = torch.zeros(2, last_B, dtype=torch.float32, device=eng.device)
            if unc is not None: buf[0], buf[1] = torch.as_tensor(unc['pred'][0], device=eng.device), torch.as_tensor(unc['model'][0], device=eng.device)
            torch.distributed.broadcast(buf, src=(nb - 1) % G)
            unc = {'pred': [buf[0].cpu().numpy()], 'model': [buf[1].cpu().numpy()]}
        if K is None: return out, unc
        if G > 1:  # every row was written by exactly one rank, zeros elsewhere
            torch.distributed.all_reduce(vals); torch.distributed.all_reduce(idx)
        return util.topk_to_sparse(torch, vals.cpu(), idx.cpu(), eng.E_total), unc


def _gather_columns_impl(torch, eng, t):
    n = eng.shard[1]
    widths = [eng.E_total * (i + 1) // n - eng.E_total * i // n for i in range(n)]
    pad = torch.zeros(t.shape[0], max(widths), dtype=t.dtype, device=t.device)
    pad[:, :t.shape[1]] = t
    parts = [torch.empty_like(pad) for _ in range(n)]
    torch.distributed.all_gather(parts, pad)
    return torch.cat([parts[i][:, :widths[i]] for i in range(n)], dim=1).cpu()


def scipy_dense(x):
    import scipy.sparse
    return not scipy.sparse.issparse(x)
