"""Device-resident training / inference engine behind `opentf_b200.fnn.Fnn` and `opentf_b200.bnn.Bnn`.

Everything numeric is a kernel of libntf_b200.so called through `ops`; this module owns the memory
(parameter arena, Adam state, activations, CSR staging) and the order of launches of one step.

Data layout in HBM:
  * one flat fp32 arena holds every parameter, 128-byte aligned segments; `grads`, `adam_m`, `adam_v` mirror it,
    so Adam (and the data-parallel gradient all-reduce) is one launch / one collective over a flat buffer.
    Layer 0's weight is stored TRANSPOSED ([S,h0]: a skill's vector is one contiguous 4*h0-byte row) -- the
    state_dict exporter/importer transposes to/from torch's [out,in].
  * the teamsvecs CSR (int32 indptr / indices; data is all ones, team.py:20-27) is staged once; a split is a
    gathered copy in batch order (`SplitData`), re-gathered per epoch with that epoch's permutation, so batch b
    is the row range [b*B, (b+1)*B) of it and kernels address it by pointer offset.
"""
import ctypes as C
import math
import os

import numpy as np
import torch

from . import _lib, ops
from ._lib import NSD, PRECISION, OutTrainArgs

ALIGN = 32  # floats (128 bytes)


def _round_up(n, a):
    return (n + a - 1) // a * a


def _pack_bits(neg):
    """bool [R,C] (True = -1) -> int32 [R, ceil(C/32)] bit planes, bit c%32 of word c/32"""
    neg = np.asarray(neg, dtype=bool)
    R, Cn = neg.shape
    pad = _round_up(max(Cn, 1), 32) - Cn
    if pad: neg = np.concatenate([neg, np.zeros((R, pad), dtype=bool)], axis=1)
    return np.ascontiguousarray(np.packbits(neg, axis=1, bitorder='little')).view(np.int32).reshape(R, -1)


def to_csr(mat):
    """accept scipy lil/csr/coo (the reference hands lil uint8, team.py:154) -> (indptr int32, indices int32, shape)."""
    c = mat.tocsr()
    c.sort_indices()
    if c.nnz >= 2 ** 31: raise ValueError('CSR with >= 2^31 entries is not supported')
    return np.ascontiguousarray(c.indptr, dtype=np.int32), np.ascontiguousarray(c.indices, dtype=np.int32), c.shape


class DeviceCSR:
    """all teams of one teamsvecs matrix on the device"""

    def __init__(self, mat, device):
        indptr, indices, self.shape = to_csr(mat)
        self.host_indptr = indptr
        self.indptr = torch.from_numpy(indptr).to(device)
        self.indices = torch.from_numpy(indices).to(device) if len(indices) else torch.zeros(1, dtype=torch.int32, device=device)
        self.nnz = int(len(indices))

    def nnz_of(self, rows):
        rows = np.asarray(rows, dtype=np.int64)
        return int((self.host_indptr[rows + 1] - self.host_indptr[rows]).sum())


class DeviceRows:
    """dense (embedded) skill vectors of all teams on the device: teamsvecs['skill'] as main.py:148-153 leaves it (ndarray [N,d])"""

    def __init__(self, mat, device):
        m = np.ascontiguousarray(np.asarray(mat), dtype=np.float32)  # ntf.py:24 `.float()`
        if m.ndim != 2: raise ValueError(f'dense skill input must be [N,d], got {m.shape}')
        self.shape = m.shape
        self.x = torch.from_numpy(m).to(device)


class SplitData:
    """teams `rows` (in that order) of the skill and member CSR, gathered on the device.  `regather(order)`
    rebuilds the copy for a new order of the same teams (the per-epoch shuffle of fnn.py:95)."""

    def __init__(self, eng, rows):
        self.eng = eng
        self.rows_host = np.ascontiguousarray(np.asarray(rows), dtype=np.int32)
        self.n = len(self.rows_host)
        dev = eng.device
        self.dense = eng.dense_input
        self.nnz_s, self.nnz_m = (0 if self.dense else eng.skill.nnz_of(self.rows_host)), eng.member.nnz_of(self.rows_host)
        if self.dense: self.x = torch.empty(max(1, self.n), eng.S, dtype=torch.float32, device=dev)  # [n,d] rows in batch order
        else:
            self.s_indptr = torch.empty(self.n + 1, dtype=torch.int32, device=dev)
            self.s_indices = torch.empty(max(1, self.nnz_s), dtype=torch.int32, device=dev)
            self.s_ent_row = torch.empty(max(1, self.nnz_s), dtype=torch.int32, device=dev)
        self.m_indptr = torch.empty(self.n + 1, dtype=torch.int32, device=dev)
        self.m_indices = torch.empty(max(1, self.nnz_m), dtype=torch.int32, device=dev)
        self.rows_dev = torch.empty(max(1, self.n), dtype=torch.int32, device=dev)
        self.regather(None)

    def regather(self, order):
        """order: positions into the split's rows (None = as given)."""
        rows = self.rows_host if order is None else self.rows_host[np.asarray(order)]
        self.rows_now = rows
        if self.n == 0: return
        self.rows_dev.copy_(torch.from_numpy(np.ascontiguousarray(rows)), non_blocking=False)
        e = self.eng
        if self.dense: ops.rows_gather(self.rows_dev, self.n, e.S, e.skill.x, self.x)
        else: ops.csr_gather(self.rows_dev, self.n, e.skill.indptr, e.skill.indices, self.s_indptr, self.s_indices, self.s_ent_row, e.ws)
        ops.csr_gather(self.rows_dev, self.n, e.member.indptr, e.member.indices, self.m_indptr, self.m_indices, None, e.ws)


class _HostStage:
    """device landing buffer of one host batch: ONE int32 block [s_indptr | m_indptr | s_indices | s_ent_row | m_indices] (one H2D copy
    per step), with the same attributes `Engine.step` reads from a SplitData as views into it"""

    def __init__(self, device, words):
        self.buf = torch.zeros(words, dtype=torch.int32, device=device)
        self.n = 0

    def view(self, n, cap_s, cap_m):
        """segment offsets depend on (n, cap_s, cap_m) only: batches packed with the same capacities land at the same addresses"""
        o, b = 0, self.buf
        self.n = n
        self.s_indptr = b[o:o + n + 1]; o += n + 1
        self.m_indptr = b[o:o + n + 1]; o += n + 1
        self.s_indices = b[o:o + cap_s]; o += cap_s
        self.s_ent_row = b[o:o + cap_s]; o += cap_s
        self.m_indices = b[o:o + cap_m]; o += cap_m
        return o


def pack_host_batch(s_ptr, s_idx, m_ptr, m_idx, cap_s=None, cap_m=None):
    """compact CSR of one batch (numpy int arrays, offsets starting at 0) -> the pinned int32 block `Engine.step_host` copies.
    cap_s / cap_m: pad the index segments to fixed capacities (>= nnz) so that every batch of a run has the same layout -- the
    device addresses then never change and the step replays one captured graph."""
    n = len(s_ptr) - 1
    cap_s = len(s_idx) if cap_s is None else int(cap_s)
    cap_m = len(m_idx) if cap_m is None else int(cap_m)
    assert cap_s >= len(s_idx) and cap_m >= len(m_idx)
    s_row = np.repeat(np.arange(n, dtype=np.int32), np.diff(s_ptr))
    pad = lambda a, c: np.concatenate([np.asarray(a, dtype=np.int32), np.zeros(c - len(a), dtype=np.int32)])
    blk = np.concatenate([s_ptr, m_ptr, pad(s_idx, cap_s), pad(s_row, cap_s), pad(m_idx, cap_m)]).astype(np.int32)
    return torch.from_numpy(blk).pin_memory(), n, cap_s, cap_m


class HostPacker:
    """packs global batches of the HOST teamsvecs CSR into pinned blocks for `Engine.step_host` (ntf_pack_host_batch: a memcpy per row), three
    blocks in rotation so that batch i+1 can be packed (on a loader thread: the call releases the GIL) while batch i is being copied and batch
    i-1 is on the GPU -- the loader side of the streaming entry point"""

    def __init__(self, skill_csr, member_csr, n, cap_s, cap_m):
        (self.sp, self.si, _), (self.mp, self.mi, _) = skill_csr, member_csr  # to_csr() triples: int32 indptr, indices
        self.n, self.cap_s, self.cap_m = int(n), int(cap_s), int(cap_m)
        self.words = 2 * (self.n + 1) + 2 * self.cap_s + self.cap_m
        self.blocks = [torch.zeros(self.words, dtype=torch.int32).pin_memory() for _ in range(3)]
        self.k = 0

    def pack(self, rows, lo=0, hi=None):
        """[lo, hi): the rows of the global batch this rank trains on (skill CSR packed for them only; members for the whole batch)"""
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        assert len(rows) == self.n
        blk = self.blocks[self.k]; self.k = (self.k + 1) % 3
        _lib.check(_lib.lib().ntf_pack_host_batch(rows.ctypes.data, self.n, self.sp.ctypes.data, self.si.ctypes.data, self.mp.ctypes.data, self.mi.ctypes.data,
                                                  self.cap_s, self.cap_m, int(lo), self.n if hi is None else int(hi), blk.data_ptr(), self.words), 'ntf_pack_host_batch')
        return blk


class Engine:
    def __init__(self, S, hidden, E, device, bayesian=False, precision='tf32', tpw=10.0, tnw=1.0, nsd='uniform', ns=5,
                 seed=0, max_batch=1000, shard=None, dense_input=False):
        """shard=(i, n): expert-sharded output layer (BASELINE config 4): this engine owns columns
        [E*i//n, E*(i+1)//n) of the last layer (weights, gradients, Adam state, special planes); everything else is replicated and
        the ranks exchange only dA [B,h] per step (`allreduce`)."""
        if not torch.cuda.is_available():
            raise _lib.NtfError('opentf_b200 needs a CUDA device: the kernels are sm_100a only and there is no CPU fallback')
        self.device = torch.device(device)
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.ctx(self.dev_index)  # fails loudly if this is not a B200-class device
        self.h = _lib.new_ctx(self.dev_index)  # this engine's own handle: side streams / events of ntf_fnn_step
        self._cap_stream = torch.cuda.Stream(device=self.device)  # graph captures run here
        self.S, self.hidden, self.E_total = int(S), [int(x) for x in hidden], int(E)
        self.shard = (0, 1) if shard is None else (int(shard[0]), int(shard[1]))
        self.e_lo, self.e_hi = self.E_total * self.shard[0] // self.shard[1], self.E_total * (self.shard[0] + 1) // self.shard[1]
        self.E = self.e_hi - self.e_lo  # the output columns this engine owns
        if self.shard[1] > 1 and bayesian: raise NotImplementedError('expert-sharded Bnn')
        # dense_input: teamsvecs['skill'] is a dense [N,d] matrix of skill embeddings (main.py:148-153, ntf.py:24): layer 0 is then
        # a dense layer in torch layout [h0,d] instead of the CSR bag over a transposed weight
        self.dense_input = bool(dense_input)
        self.bayesian, self.precision = bool(bayesian), PRECISION[precision]
        # 'tf32' selects the tcgen05 kernels where they exist for the shape; other shapes (toy sizes, odd widths) run the
        # CUDA-core fp32 kernels of the same library -- both are sm_100a code, neither is a fallback to another backend.
        if self.precision == _lib.NTF_TF32 and not _lib.lib().ntf_tc_supported(int(max_batch), int(hidden[-1]), int(E), int(bool(bayesian))):
            self.precision = _lib.NTF_FP32
        self.tpw, self.tnw, self.nsd, self.ns, self.seed = float(tpw), float(tnw), NSD[nsd], int(ns), int(seed)
        self.Bmax = int(max_batch)
        self.sizes = [self.S] + self.hidden + [self.E]
        self.L = len(self.sizes) - 1  # number of linear layers
        self.ws = ops.Workspace(self.device)
        self.skill = self.member = None
        self._layout()
        self._buffers()
        self.adam_t = 0
        # CUDA graphs (include/ntf_b200.h): batch i of a split is the same launch sequence in every epoch, so it is captured the first
        # time it runs and replayed afterwards; the scalars that do change (RNG counter, lr, Adam's bias corrections) sit in `dyn`.
        self.use_graphs = os.environ.get('NTF_GRAPHS', '1') != '0'
        self.dyn = torch.zeros(8, dtype=torch.int32, device=self.device)  # ntf_dyn (32 bytes)
        self._graphs, self.max_graphs, self._ws_gen = {}, 8192, -1
        self.graph_event_factory = None  # (bench.py) called at capture time -> a cudaEvent_t pair recorded around the output layer
        self.global_step = 0  # counts train AND valid steps: the sampler's Philox counter (fnn.py:148 samples in valid too)
        self.world, self.rank = 1, 0
        # data-parallel ranks: exchange the gradient arena in two overlapped segments (NTF_DP_OVERLAP=0: one all-reduce after the step)
        self.dp_overlap = os.environ.get('NTF_DP_OVERLAP', '1') != '0'
        self._dp_stream = self._dp_ev = None
        self.comm = None  # nccl.Comm: the library then runs the exchange inside ntf_fnn_step (attach_comm)
        self.peers = None  # _lib.Peers: exchange + Adam fused over peer memory inside ntf_fnn_step (attach_peers); preferred over comm

    # ------------------------------------------------------------------ memory
    def _layout(self):
        """name -> (offset, stored shape).  Fnn: layers.i.weight/bias.  Bnn: layers.i.{mu,rho}_{weight,bias}."""
        self.views, off = {}, 0
        kinds = ['mu_', 'rho_'] if self.bayesian else ['']
        for i in range(self.L):
            fin, fout = self.sizes[i], self.sizes[i + 1]
            wshape = (fin, fout) if (i == 0 and not self.dense_input) else (fout, fin)  # layer 0 stored transposed (CSR bag)
            for k in kinds:
                self.views[f'layers.{i}.{k}weight'] = (off, wshape); off += _round_up(fin * fout, ALIGN)
            for k in kinds:
                self.views[f'layers.{i}.{k}bias'] = (off, (fout,)); off += _round_up(fout, ALIGN)
        self.n_params = off
        if self.bayesian:  # per-step noise / perturbation buffers share one layout (one slot per (weight|bias) tensor)
            self.nviews, off = {}, 0
            for i in range(self.L):
                fin, fout = self.sizes[i], self.sizes[i + 1]
                self.nviews[f'{i}.weight'] = (off, (fin, fout) if (i == 0 and not self.dense_input) else (fout, fin)); off += _round_up(fin * fout, ALIGN)
                self.nviews[f'{i}.bias'] = (off, (fout,)); off += _round_up(fout, ALIGN)
            self.n_noise = off

    def _buffers(self):
        dev, f32 = self.device, torch.float32
        self.params = torch.zeros(self.n_params, dtype=f32, device=dev)
        self.grads = torch.zeros(self.n_params, dtype=f32, device=dev)
        self.adam_m = torch.zeros(self.n_params, dtype=f32, device=dev)
        self.adam_v = torch.zeros(self.n_params, dtype=f32, device=dev)
        B = self.Bmax
        self.act = [torch.empty(B, h, dtype=f32, device=dev) for h in self.hidden]
        self.dact = [torch.empty(B, h, dtype=f32, device=dev) for h in self.hidden]
        self.dz = [torch.empty(B, h, dtype=f32, device=dev) for h in self.hidden]
        self.pitch = _round_up(self.E, 32) // 32
        self.special = torch.zeros(B, self.pitch, dtype=torch.int32, device=dev)
        if self.precision == _lib.NTF_TF32:  # the tensor-core kernel reads the same sets as tile-transposed planes
            words = ops.special_tiles_bytes(B, self.E) // 4
            self.special_t = torch.zeros(words, dtype=torch.int32, device=dev)
            self.member_t = torch.zeros(words, dtype=torch.int32, device=dev)
        self.neg = torch.full((B, max(1, self.ns)), -1, dtype=torch.int32, device=dev)
        self.counts = torch.zeros(self.E_total, dtype=torch.int32, device=dev)  # the sampler works on the global expert axis
        self.cdf = torch.zeros(self.E_total, dtype=torch.int32, device=dev)
        self.loss_buf = torch.zeros(4096, dtype=f32, device=dev)
        if self.bayesian:
            self.eps = torch.zeros(self.n_noise, dtype=f32, device=dev)
            self.delta = torch.zeros(self.n_noise, dtype=f32, device=dev)
            self.gdelta = torch.zeros(self.n_noise, dtype=f32, device=dev)
            self.kl = torch.zeros(1, dtype=f32, device=dev)
            self.ent_sum = torch.zeros(B, dtype=f32, device=dev)
            self.act_s = [torch.empty(B, h, dtype=f32, device=dev) for h in self.hidden]
            self.dact_s = [torch.empty(B, h, dtype=f32, device=dev) for h in self.hidden]
            self.dzs = [torch.empty(B, h, dtype=f32, device=dev) for h in self.hidden]
            # (CSR input: the input signs of layer 0 exist only at the nnz positions, `ent_sign`; dense input: a [B, S] plane like the other layers')
            self.sign_in = [torch.zeros(B, _round_up(self.S, 32) // 32, dtype=torch.int32, device=dev) if self.dense_input else None]
            self.sign_in += [torch.zeros(B, _round_up(h, 32) // 32, dtype=torch.int32, device=dev) for h in self.hidden]
            if self.dense_input: self.x_s = torch.empty(B, self.S, dtype=f32, device=dev)  # x * s_in of layer 0
            self.sign_out = [torch.zeros(B, _round_up(o, 32) // 32, dtype=torch.int32, device=dev) for o in self.sizes[1:]]

    def view(self, name, buf=None):
        off, shape = self.views[name]
        return (self.params if buf is None else buf)[off:off + int(np.prod(shape))].view(*shape)

    def nview(self, buf, key):
        off, shape = self.nviews[key]
        return buf[off:off + int(np.prod(shape))].view(*shape)

    # ------------------------------------------------------------------ parameters
    def load_state_dict(self, sd):
        """torch-layout state dict (Fnn: layers.i.weight/bias; Bnn: layers.i.{mu,rho}_{weight,bias}) -> arena."""
        missing = [k for k in self.views if k not in sd]
        if missing: raise KeyError(f'state_dict is missing {missing}')
        for name in self.views:
            t = torch.as_tensor(sd[name]).detach().to(torch.float32)
            if name.startswith('layers.0.') and name.endswith('weight') and not self.dense_input: t = t.t()
            if name.startswith(f'layers.{self.L - 1}.') and self.shard[1] > 1: t = t[self.e_lo:self.e_hi]  # this shard's experts
            v = self.view(name)
            if tuple(t.shape) != tuple(v.shape): raise ValueError(f'{name}: shape {tuple(t.shape)} does not fit {tuple(v.shape)}')
            v.copy_(t.contiguous())
        self._w16_dirty = True

    def state_dict(self, gather=True):
        """torch-layout state dict.  Expert-sharded: the last layer's shards are all-gathered (a collective: every rank calls it);
        gather=False returns this shard only."""
        out = {}
        for name in self.views:
            t = self.view(name).detach()
            if name.startswith(f'layers.{self.L - 1}.') and self.shard[1] > 1 and gather:
                n = self.shard[1]
                rows = max(self.E_total * (i + 1) // n - self.E_total * i // n for i in range(n))
                pad = torch.zeros((rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
                pad[:t.shape[0]] = t
                parts = [torch.empty_like(pad) for _ in range(n)]
                torch.distributed.all_gather(parts, pad)
                t = torch.cat([parts[i][:self.E_total * (i + 1) // n - self.E_total * i // n] for i in range(n)])
            t = t.cpu().clone()
            if name.startswith('layers.0.') and name.endswith('weight') and not self.dense_input: t = t.t().contiguous()
            out[name] = t
        return out

    def allreduce(self, t):
        """SUM over the ranks that share a step (tests of the sharded layer replace this with an in-process sum)"""
        torch.distributed.all_reduce(t)

    def reset_optimizer(self):
        self.adam_m.zero_(); self.adam_v.zero_(); self.adam_t = 0

    # ------------------------------------------------------------------ data
    def stage(self, skill_mat, member_mat):
        if self.dense_input: self.skill = DeviceRows(skill_mat, self.device)
        else: self.skill = DeviceCSR(skill_mat, self.device)
        self.member = DeviceCSR(member_mat, self.device)
        assert self.skill.shape[1] == self.S and self.member.shape[1] == self.E_total and self.skill.shape[0] == self.member.shape[0]
        if self.bayesian and not self.dense_input:  # Flipout input signs exist only at the nnz positions: one bit per CSR entry of a batch
            maxlen = int(np.diff(self.skill.host_indptr).max()) if self.skill.shape[0] else 1
            self._ent_sign_words(self.Bmax * max(1, maxlen))
        return self

    def _ent_sign_words(self, entries):
        words = _round_up(max(1, entries), 128) // 32
        if getattr(self, 'ent_sign', None) is None or self.ent_sign.numel() < words:
            self.ent_sign = torch.zeros(words, dtype=torch.int32, device=self.device)
        return words

    def split(self, rows):
        return SplitData(self, rows)

    def set_global_unigram(self):
        """fnn.py:82: expert frequency over ALL teams -> integer counts + CDF (nsd == 'unigram')."""
        N = self.member.shape[0]
        ops.expert_cdf(N, self.member.indptr.data_ptr(), self.member.indices, self.E_total, self.counts, self.cdf, self.ws)

    # ------------------------------------------------------------------ one step
    def _forward_hidden(self, sp, b0, B):
        if self.dense_input:
            ops.dense_fwd(sp.x[b0:b0 + B], self.view('layers.0.weight'), self.view('layers.0.bias'), B, self.S, self.hidden[0], 1, self.act[0])
        else: ops.csr_bag_fwd(B, sp.s_indptr.data_ptr() + 4 * b0, sp.s_indices, self.view('layers.0.weight'), self.view('layers.0.bias'),
                        self.S, self.hidden[0], self.act[0])
        for i in range(1, self.L - 1):
            ops.dense_fwd(self.act[i - 1], self.view(f'layers.{i}.weight'), self.view(f'layers.{i}.bias'), B, self.hidden[i - 1],
                          self.hidden[i], 1, self.act[i])

    def _sample(self, sp, b0, B, neg_host, gbatch=None):
        """fills self.neg for this batch (or takes host-supplied indices: the parity-test contract).  `gbatch` =
        (first row, size) of the GLOBAL batch this slice belongs to: unigram_b counts experts over all of it
        (every rank holds the whole CSR, so no collective is needed for the histogram)."""
        if self.nsd == 0 and neg_host is None: return None
        if neg_host is not None:
            t = torch.as_tensor(np.ascontiguousarray(neg_host), dtype=torch.int32)
            assert t.shape[0] == B
            neg = self.neg[:B, :t.shape[1]] if t.shape[1] == self.neg.shape[1] else torch.empty(B, t.shape[1], dtype=torch.int32, device=self.device)
            neg.copy_(t)
            return neg.contiguous()
        mptr = sp.m_indptr.data_ptr() + 4 * b0
        g0, gB = (b0, B) if gbatch is None else gbatch  # unigram_b draws from the member CSR of the global batch
        ops.neg_sample(self.nsd, self.seed, self.global_step, b0, B, mptr, sp.m_indices, self.E_total, self.ns,
                       self.cdf if self.nsd == NSD['unigram'] else None, self.neg, sp.m_indptr.data_ptr() + 4 * g0, gB)
        return self.neg

    def _fnn_args(self):
        """the ntf_fnn_step_args of this engine with everything that does not change from step to step filled in"""
        a = getattr(self, '_fa', None)
        if a is not None: return a
        a = _lib.FnnStepArgs()
        a.n_layers, a.S, a.E, a.e_lo, a.E_total = self.L, self.S, self.E, self.e_lo, self.E_total
        for i, hh in enumerate(self.hidden):
            a.hidden[i] = hh
            a.act[i], a.dact[i], a.dz[i] = self.act[i].data_ptr(), self.dact[i].data_ptr(), self.dz[i].data_ptr()
        for i in range(self.L):
            a.W[i], a.b[i] = self.view(f'layers.{i}.weight').data_ptr(), self.view(f'layers.{i}.bias').data_ptr()
            a.gW[i], a.gb[i] = self.view(f'layers.{i}.weight', self.grads).data_ptr(), self.view(f'layers.{i}.bias', self.grads).data_ptr()
        a.nsd, a.seed, a.ns = self.nsd, self.seed, self.ns
        a.neg, a.counts, a.cdf = self.neg.data_ptr(), self.counts.data_ptr(), self.cdf.data_ptr()
        a.precision, a.tpw, a.tnw = self.precision, self.tpw, self.tnw
        if self.precision == _lib.NTF_TF32: a.special_t, a.member_t = self.special_t.data_ptr(), self.member_t.data_ptr()
        else: a.special, a.pitch_words = self.special.data_ptr(), self.pitch
        a.params, a.grads, a.adam_m, a.adam_v, a.n_params = self.params.data_ptr(), self.grads.data_ptr(), self.adam_m.data_ptr(), self.adam_v.data_ptr(), self.n_params
        a.beta1, a.beta2, a.eps = 0.9, 0.999, 1e-8
        self._fa = a
        return a

    def step(self, sp, b0, B, train, lr=None, loss_slot=0, neg_host=None, loss_scale=None, gbatch=None, noise_host=None):
        """one batch = rows [b0, b0+B) of split `sp`: forward + loss (+ backward + Adam when train), enqueued by one call of the
        library (ntf_fnn_step).  The loss lands in self.loss_buf[loss_slot] (device); nothing is synchronised here."""
        assert 0 < B <= self.Bmax and b0 + B <= sp.n
        if self.bayesian: return self._step_bayes(sp, b0, B, train, lr, loss_slot, neg_host, loss_scale, gbatch, noise_host)
        a = self._fnn_args()
        a.B, a.row_base, a.row0 = B, b0, b0
        if self.dense_input: a.x_dense, a.s_indptr, a.s_indices, a.s_ent_row = sp.x.data_ptr() + 4 * self.S * b0, None, None, None
        else: a.x_dense, a.s_indptr, a.s_indices, a.s_ent_row = None, sp.s_indptr.data_ptr() + 4 * b0, sp.s_indices.data_ptr(), sp.s_ent_row.data_ptr()
        a.m_indptr, a.m_indices = sp.m_indptr.data_ptr() + 4 * b0, sp.m_indices.data_ptr()
        if gbatch is None: a.gB, a.g_m_indptr = 0, None
        else: a.gB, a.g_m_indptr = gbatch[1], sp.m_indptr.data_ptr() + 4 * gbatch[0]
        a.step = self.global_step
        a.neg_given, a.ns, a.neg = 0, self.ns, self.neg.data_ptr()
        if neg_host is not None:  # host-supplied indices: the parity-test contract
            t = torch.as_tensor(np.ascontiguousarray(neg_host), dtype=torch.int32)
            assert t.shape[0] == B
            if t.shape[1] == self.neg.shape[1]: neg = self.neg[:B]
            else: neg = self._neg_alt = torch.empty(B, t.shape[1], dtype=torch.int32, device=self.device)
            neg.copy_(t)
            a.neg_given, a.ns, a.neg = 1, t.shape[1], neg.data_ptr()
        a.loss_scale = 1.0 / B if loss_scale is None else loss_scale
        a.loss_out = self.loss_buf.data_ptr() + 4 * loss_slot
        sharded = self.shard[1] > 1
        dp = self.world > 1 and not sharded  # data-parallel ranks all-reduce the gradient arena before Adam
        in_step = dp and train and (self.comm is not None or self.peers is not None)  # the library exchanges the gradients itself
        a.train, a.run_adam = int(bool(train)), int(bool(train) and (not dp or in_step))
        a.comm, a.allreduce = (self.comm.ptr, self.comm.allreduce_addr) if (in_step and self.peers is None) else (None, None)
        a.peers = C.addressof(self.peers) if (in_step and self.peers is not None) else None
        sp_tab = getattr(self, 'shard_peers', None)
        if sharded and train and sp_tab is not None: a.peers = C.addressof(sp_tab)  # the dA exchange runs inside the step (ntf_peer_allreduce)
        if train:
            a.lr, a.adam_t = float(lr), self.adam_t + 1
        pe = getattr(self, 'prof_events', None)  # (bench.py) a recorded-once torch event pair around the output-layer call
        a.prof_ev[0], a.prof_ev[1] = (pe[0].cuda_event, pe[1].cuda_event) if pe else (None, None)
        # tensor-core output layer: the fp16 weight image (made here if stale; the step's own Adam launch keeps it current afterwards)
        a.W16 = self.w16().data_ptr() if self.precision == _lib.NTF_TF32 else None
        a.dyn = None
        graphed = self.use_graphs and neg_host is None and pe is None
        if graphed:
            # everything that is baked into the captured launches; the rest (step counter, lr, adam_t) goes through `dyn`
            key = (a.comm, a.peers, a.x_dense, a.s_indptr, a.s_indices, a.s_ent_row, a.m_indptr, a.m_indices, b0, B, a.train, a.run_adam, a.loss_out, a.loss_scale, a.gB, a.g_m_indptr)
            graphed = len(self._graphs) < self.max_graphs or any(key + (ph,) in self._graphs for ph in (1, 3))
        if graphed:
            a.dyn = self.dyn.data_ptr()
            ops.dyn_update(self.dev_index, self.dyn, self.global_step, float(lr) if train else 0.0, 0.9, 0.999, 1e-8, self.adam_t + 1)
        if sharded and train and sp_tab is not None:
            self._run(a, 3, key if graphed else None)  # one call: the shards' dA are summed over peer memory between the two halves
        elif sharded and train:
            # every rank runs the same batch on its expert range; the only exchange is dA = sum over shards of dz W  [B,h]
            self._run(a, 1, key if graphed else None)
            self.allreduce(self.dact[-1][:B])
            self._run(a, 2, key if graphed else None)
        elif in_step:
            self._run(a, 3, key if graphed else None)
            self.global_step += 1; self.adam_t += 1
            return
        elif dp and train and self.dp_overlap and self._arena_split() is not None:
            return self._dp_overlapped(a, key if graphed else None, lr)
        else:
            self._run(a, 3, key if graphed else None)
        self.global_step += 1
        if not train: return
        if dp: self.optimizer_step(lr)
        else: self.adam_t += 1

    def clear_graphs(self):
        """forget the captured step graphs (they are captured again on the next use)"""
        self._graphs.clear()

    def attach_comm(self, comm=None):
        """data-parallel ranks: hand the library an NCCL communicator (nccl.Comm) so that ntf_fnn_step exchanges the gradients itself --
        two all-reduces on its own stream inside the (captured) step instead of torch.distributed calls between two halves of it."""
        if comm is None:
            from . import nccl
            comm = nccl.Comm(self.rank, self.world, self.device)
        self.comm = comm
        self._graphs.clear()
        return comm

    def attach_peers(self, local=None):
        """data-parallel ranks: move the parameter and gradient arenas into IPC-shareable blocks, exchange the handles over the existing
        process group and hand the library the table of every rank's arenas (ntf_peers): the step then runs reduce-scatter + Adam +
        all-gather as one pass over peer memory (csrc/peer.cu) -- no NCCL on the data path.  `local`: engines of all ranks living in THIS
        process on this GPU (tests: two ranks on one device, on different streams); their arenas are addressed directly."""
        d = self.dev_index
        if getattr(self, '_peer_blocks', None) is None:
            new_p, pp = ops.peer_alloc(d, self.n_params, torch.float32)
            new_g, pg = ops.peer_alloc(d, self.n_params, torch.float32)
            fl, pf = ops.peer_alloc(d, 64, torch.int32)  # NTF_PEER_FLAG_BYTES
            new_p.copy_(self.params); new_g.copy_(self.grads)
            self.params, self.grads, self._peer_flags = new_p, new_g, fl
            self._peer_blocks = (pp, pg, pf)
            self._fa = None
            self._graphs.clear()
        torch.cuda.synchronize(self.device)
        if local is not None:
            if any(getattr(e, '_peer_blocks', None) is None for e in local): return None  # (the last engine of the list builds every table)
            for e in local:
                t = _lib.Peers()
                t.rank, t.world = e.rank, len(local)
                for r, o in enumerate(local): t.params[r], t.grads[r], t.flags[r] = o._peer_blocks
                e.peers = t
            return self.peers
        handles = tuple(ops.peer_export(d, p) for p in self._peer_blocks)
        got = [None] * self.world
        torch.distributed.all_gather_object(got, handles)
        t = _lib.Peers()
        t.rank, t.world = self.rank, self.world
        self._imported = []
        for r, hs in enumerate(got):
            if r == self.rank: ptrs = self._peer_blocks
            else:
                ptrs = tuple(ops.peer_import(d, h) for h in hs)
                self._imported += list(ptrs)
            t.params[r], t.grads[r], t.flags[r] = ptrs
        torch.distributed.barrier()  # every rank has opened every block before anyone signals through it
        self.peers = t
        return t

    def attach_shard_peers(self, local=None):
        """expert-sharded output layer: put this rank's dA partial [Bmax, h_last] into an IPC-shareable block, exchange the handles and hand
        the library the table (ntf_peers with grads[r] = rank r's block): ntf_fnn_step then sums the shards' dA itself (ntf_peer_allreduce,
        one pass over peer memory between two flag barriers) and a sharded step becomes ONE call / one CUDA graph instead of two halves
        around a torch.distributed.all_reduce.  `local`: a list with this engine alone = a table of one rank (single-GPU plumbing tests)."""
        d = self.dev_index
        if getattr(self, '_x_blocks', None) is None:
            xb, px = ops.peer_alloc(d, self.Bmax * self.hidden[-1], torch.float32)
            fl, pf = ops.peer_alloc(d, 64, torch.int32)
            self._x_buf, self._peer_flags, self._x_blocks = xb, fl, (px, pf)
        torch.cuda.synchronize(self.device)
        t = _lib.Peers()
        if local is not None:
            assert local == [self]
            t.rank, t.world = 0, 1
            t.grads[0], t.params[0], t.flags[0] = self._x_blocks[0], self._x_blocks[0], self._x_blocks[1]
        else:
            handles = tuple(ops.peer_export(d, p) for p in self._x_blocks)
            got = [None] * self.shard[1]
            torch.distributed.all_gather_object(got, handles)
            t.rank, t.world = self.shard
            self._imported = []
            for r, hs in enumerate(got):
                if r == self.shard[0]: ptrs = self._x_blocks
                else:
                    ptrs = tuple(ops.peer_import(d, h) for h in hs)
                    self._imported += list(ptrs)
                t.grads[r], t.params[r], t.flags[r] = ptrs[0], ptrs[0], ptrs[1]
            torch.distributed.barrier()
        self.shard_peers = t
        self._graphs.clear()
        return t

    def peer_error(self):
        """non-zero if a barrier of the peer exchange gave up waiting (a rank died or fell out of step)"""
        return 0 if (self.peers is None and getattr(self, 'shard_peers', None) is None) else int(self._peer_flags[36].item())

    def _peer_exchange(self, lr):
        """the eager counterpart of what ntf_fnn_step does with `peers`: output layer's segment on channel 1, the rest on channel 0"""
        self._w16_dirty = True
        split = self._arena_split()
        t = self.adam_t + 1
        if split is not None:
            ops.peer_exchange_adam(self.dev_index, self.peers, self.adam_m, self.adam_v, split, self.n_params - split, lr, 0.9, 0.999, 1e-8, t, 1)
            ops.peer_exchange_adam(self.dev_index, self.peers, self.adam_m, self.adam_v, 0, split, lr, 0.9, 0.999, 1e-8, t, 0)
        else:
            ops.peer_exchange_adam(self.dev_index, self.peers, self.adam_m, self.adam_v, 0, self.n_params, lr, 0.9, 0.999, 1e-8, t, 0)
        self.adam_t = t

    def _last_kinds(self):
        return ('mu_weight', 'rho_weight', 'mu_bias', 'rho_bias') if self.bayesian else ('weight', 'bias')

    def _arena_split(self):
        """offset of the output layer's segment when a step exchanges the arena in two overlapped segments, else None -- the SAME decision
        ntf_fnn_step takes (step.cu: NTF_ADAM_SPLIT unset, the last layer's tensors at the end of the arena, offset a multiple of 32 floats),
        so that a rank that sits a batch out (idle_step) issues the same number of collectives as the ranks that step"""
        if os.environ.get('NTF_ADAM_SPLIT') is not None: return None
        split = min(self.views[f'layers.{self.L - 1}.{k}'][0] for k in self._last_kinds())
        others = [self.views[f'layers.{i}.{k}'][0] for i in range(self.L - 1) for k in self._last_kinds()]
        return split if (split % 32 == 0 and 0 < split < self.n_params and all(o < split for o in others)) else None

    def idle_step(self, train, lr=None, loss_scale=None, loss_slot=None):
        """a data-parallel rank whose slice of a (short, last) global batch is empty: it contributes zero data gradients but must take part in
        the exchange and step its replica with the same sums.  The exchange mirrors step()'s (two segments when overlapped).  Bnn: every rank
        adds its 1/world share of KL/B to the loss and of the KL gradient (engine._bayes_body), so an idle rank still owes that share."""
        step_now = self.global_step
        self.global_step += 1
        if self.world <= 1 or self.shard[1] > 1: return
        if self.bayesian and loss_scale is not None:
            ops.fill_normal(self.seed, step_now, 0, self.n_noise, self.eps)  # eps is the same on every rank (same counter)
            self._prepare_bayes()
            if loss_slot is not None: ops.axpy(1, loss_scale / self.world, self.kl, self.loss_buf[loss_slot:loss_slot + 1])
        if not train: return
        self.grads.zero_()
        if self.bayesian and loss_scale is not None:
            self.gdelta.zero_()
            for i in range(self.L):
                for what in ('weight', 'bias'):
                    mu, rho = self._pv(i, 'mu', what), self._pv(i, 'rho', what)
                    n = mu.numel()
                    ops.flipout_grads(mu, rho, self.nview(self.eps, f'{i}.{what}'), self.nview(self.gdelta, f'{i}.{what}'), n, loss_scale / (n * self.world),
                                      self._pv(i, 'mu', what, self.grads), self._pv(i, 'rho', what, self.grads))
        if self.peers is not None: return self._peer_exchange(lr)
        if self.bayesian or not (self.dp_overlap or self.comm): return self.optimizer_step(lr)
        split = self._arena_split()
        ar = self.comm.allreduce if self.comm else self.allreduce
        if split is None: ar(self.grads)
        else:
            ar(self.grads[split:])
            ar(self.grads[:split])
        self.adam_t += 1
        self._w16_dirty = True
        ops.adam_step(self.params, self.grads, self.adam_m, self.adam_v, self.n_params, lr, 0.9, 0.999, 1e-8, self.adam_t)

    def _dp_overlapped(self, a, key, lr):
        """data-parallel train step with the gradient exchange cut in two along the arena: the output layer's
        segment (its gradients are final after phase 1) is summed over the ranks and stepped on a side stream WHILE the backward
        pass through the hidden layers runs; the rest follows on the main stream while the side stream runs its Adam segment.
        Adam is elementwise, so the result equals optimizer_step()'s bit for bit given the same sums."""
        split = self._arena_split()
        n, t = self.n_params, self.adam_t + 1
        main = torch.cuda.current_stream(self.device)
        if self._dp_stream is None:
            self._dp_stream = torch.cuda.Stream(device=self.device)
            self._dp_ev = (torch.cuda.Event(), torch.cuda.Event())
        self._run(a, 1, key)
        self._dp_ev[0].record(main)
        with torch.cuda.stream(self._dp_stream):
            self._dp_stream.wait_event(self._dp_ev[0])
            self.allreduce(self.grads[split:])
            ops.adam_step(self.params[split:], self.grads[split:], self.adam_m[split:], self.adam_v[split:], n - split, lr, 0.9, 0.999, 1e-8, t)
            self._dp_ev[1].record(self._dp_stream)
        self._run(a, 2, key)
        self.allreduce(self.grads[:split])
        ops.adam_step(self.params, self.grads, self.adam_m, self.adam_v, split, lr, 0.9, 0.999, 1e-8, t)
        main.wait_event(self._dp_ev[1])
        self._w16_dirty = True
        self.adam_t = t
        self.global_step += 1

    def _run(self, a, phase, key):
        """enqueue one phase of ntf_fnn_step: directly (key None), or as the replay of its captured graph"""
        a.phase = phase
        if key is None:
            ops.fnn_step(self.dev_index, a, self.ws, self.h)
            return
        ops.fnn_step_workspace(self.dev_index, a, self.ws)  # the workspace must exist before a capture starts ...
        if self.ws.gen != self._ws_gen:  # ... and a workspace that had to grow moved: graphs captured over the old one are void
            self._graphs.clear(); self._ws_gen = self.ws.gen
        g = self._graphs.get(key + (phase,))
        if g is None:
            ev = self.graph_event_factory() if (self.graph_event_factory and phase & 1) else None
            if ev: a.prof_ev[0], a.prof_ev[1] = ev
            g = ops.Graph(self.dev_index, self.h, self._cap_stream)
            with g:
                ops.fnn_step(self.dev_index, a, self.ws, self.h)
            a.prof_ev[0], a.prof_ev[1] = None, None
            self._graphs[key + (phase,)] = g
        g.launch()

    def __del__(self):
        try:
            self._graphs.clear()
            if getattr(self, 'h', None): _lib.lib().ntf_destroy(self.h)
        except Exception: pass
        self.h = None

    def drop_graphs(self):
        """forget every captured step (call after buffers the graphs point into were re-allocated)"""
        self._graphs.clear()

    def optimizer_step(self, lr):
        if self.world > 1 and self.peers is not None: return self._peer_exchange(lr)  # exchange + Adam in one pass over peer memory
        if self.world > 1:
            self.allreduce(self.grads)  # sum over ranks; every rank scaled its loss by 1/B_global
        self.adam_t += 1
        self._w16_dirty = True
        ops.adam_step(self.params, self.grads, self.adam_m, self.adam_v, self.n_params, lr, 0.9, 0.999, 1e-8, self.adam_t)

    # ------------------------------------------------------------------ Bnn (Flipout): bnn.py:19-25, fnn.py:126-151 with is_bayesian
    def _pv(self, i, kind, what, buf=None):
        return self.view(f'layers.{i}.{kind}_{what}', buf)

    def _draw_noise(self, sp, b0, B, noise_host):
        """one LinearFlipout draw per layer (eps for weight and bias shared by the batch, +-1 input / output signs per team):
        counter RNG on the device, or host-supplied tensors (the parity-test contract, oracle.draw_flipout_noise layout)."""
        step = self.global_step
        if noise_host is None:
            sid = 64 * self.rank  # signs are per team: ranks draw different ones; eps must be the same on every rank
            ops.fill_normal(self.seed, step, 0, self.n_noise, self.eps)
            if not self.dense_input: ops.fill_sign_bits(self.seed, step, 1 + sid, self.ent_sign.numel(), self.ent_sign)
            for i in range(self.L):
                if i or self.dense_input: ops.fill_sign_bits(self.seed, step, 1 + 2 * i + sid, B * self.sign_in[i].shape[1], self.sign_in[i])
                ops.fill_sign_bits(self.seed, step, 2 + 2 * i + sid, B * self.sign_out[i].shape[1], self.sign_out[i])
            return
        if not self.dense_input:
            indptr = sp.s_indptr[b0:b0 + B + 1].cpu().numpy(); idx = sp.s_indices[int(indptr[0]):int(indptr[-1])].cpu().numpy()
        for i, nz in enumerate(noise_host):
            ew = torch.as_tensor(nz['eps_w'], dtype=torch.float32)
            self.nview(self.eps, f'{i}.weight').copy_((ew.t() if (i == 0 and not self.dense_input) else ew).contiguous())
            self.nview(self.eps, f'{i}.bias').copy_(torch.as_tensor(nz['eps_b'], dtype=torch.float32))
            s_in, s_out = np.asarray(nz['s_in']), np.asarray(nz['s_out'])
            if i == 0 and not self.dense_input:
                rows = np.repeat(np.arange(B), np.diff(indptr))
                self._ent_sign_words(len(idx))
                packed = _pack_bits(s_in[rows, idx][None, :] < 0)[0]  # the input sign matters only where x = 1: one bit per CSR entry
                self.ent_sign[:len(packed)].copy_(torch.from_numpy(packed))
            else:
                self.sign_in[i][:B].copy_(torch.from_numpy(_pack_bits(s_in < 0)))
            self.sign_out[i][:B].copy_(torch.from_numpy(_pack_bits(s_out < 0)))

    def _prepare_bayes(self):
        """delta = softplus(rho)*eps for every tensor, KL (sum over tensors of the per-tensor MEAN) -> self.kl"""
        self.kl.zero_()
        for i in range(self.L):
            for what in ('weight', 'bias'):
                mu, rho = self._pv(i, 'mu', what), self._pv(i, 'rho', what)
                n = mu.numel()
                ops.flipout_prepare(mu, rho, self.nview(self.eps, f'{i}.{what}'), n, 1.0 / n, self.nview(self.delta, f'{i}.{what}'), self.kl, self.ws)

    def _forward_hidden_bayes(self, sp, b0, B):
        h = self.hidden
        if self.dense_input:  # ntf.py:24: embedded skills -- layer 0 is a dense Flipout layer like the hidden ones
            x = sp.x[b0:b0 + B]
            ops.apply_sign(x, self.sign_in[0], self.sign_in[0].shape[1], B, self.S, self.x_s)
            ops.dense_flipout_fwd(x, self._pv(0, 'mu', 'weight'), self._pv(0, 'mu', 'bias'), self.x_s, self.nview(self.delta, '0.weight'),
                                  self.nview(self.delta, '0.bias'), self.sign_out[0], self.sign_out[0].shape[1], B, self.S, h[0], 1, self.act[0], self.ws)
        else:
            ops.csr_bag_flipout_fwd(B, sp.s_indptr.data_ptr() + 4 * b0, sp.s_indices, self.ent_sign, self._pv(0, 'mu', 'weight'), self._pv(0, 'mu', 'bias'),
                                    self.nview(self.delta, '0.weight'), self.nview(self.delta, '0.bias'), self.sign_out[0], self.sign_out[0].shape[1],
                                    self.S, h[0], self.act[0])
        for i in range(1, self.L):  # A*s_in for the next layer (the output layer included)
            ops.apply_sign(self.act[i - 1], self.sign_in[i], self.sign_in[i].shape[1], B, h[i - 1], self.act_s[i - 1])
            if i == self.L - 1: break
            ops.dense_flipout_fwd(self.act[i - 1], self._pv(i, 'mu', 'weight'), self._pv(i, 'mu', 'bias'), self.act_s[i - 1],
                                  self.nview(self.delta, f'{i}.weight'), self.nview(self.delta, f'{i}.bias'), self.sign_out[i],
                                  self.sign_out[i].shape[1], B, h[i - 1], h[i], 1, self.act[i], self.ws)

    def _step_bayes(self, sp, b0, B, train, lr, loss_slot, neg_host, loss_scale, gbatch, noise_host=None):
        """the Bnn step is ~45 calls of the library driven from here; like the Fnn step it is captured once per batch as a CUDA graph
        (the calls land on the capture stream; RNG counter / lr / Adam constants come from the ntf_dyn block: ntf_set_dyn) and replayed.
        Eager when the caller supplies negatives or noise (parity tests), for the very first step (workspaces grow to their final size
        outside a capture) and when the gradients travel through torch.distributed."""
        graphed = (self.use_graphs and neg_host is None and noise_host is None and getattr(self, '_bayes_warm', False)
                   and (self.world == 1 or not train or self.peers is not None))
        key = None
        if graphed:
            skey = (sp.x.data_ptr(),) if self.dense_input else (sp.s_indptr.data_ptr(), sp.s_indices.data_ptr(), sp.s_ent_row.data_ptr())
            key = ('bayes',) + skey + (sp.m_indptr.data_ptr(), sp.m_indices.data_ptr(),
                   b0, B, bool(train), loss_slot, loss_scale, gbatch, self.loss_buf.data_ptr(), id(self.peers))
            graphed = key in self._graphs or len(self._graphs) < self.max_graphs
        if not graphed:
            self._bayes_body(sp, b0, B, train, lr, loss_slot, neg_host, loss_scale, gbatch, noise_host)
            self._bayes_warm = True
            return
        ops.dyn_update(self.dev_index, self.dyn, self.global_step, float(lr) if train else 0.0, 0.9, 0.999, 1e-8, self.adam_t + 1)
        if self.ws.gen != self._ws_gen: self._graphs.clear(); self._ws_gen = self.ws.gen
        g = self._graphs.get(key)
        if g is None:
            counters = (self.global_step, self.adam_t)
            ops.set_dyn(self.dev_index, self.dyn)
            try:
                g = ops.Graph(self.dev_index, self.h, self._cap_stream)
                with g: self._bayes_body(sp, b0, B, train, lr if train else 0.0, loss_slot, None, loss_scale, gbatch, None)
            finally:
                ops.set_dyn(self.dev_index, None)
            self.global_step, self.adam_t = counters  # (the captured calls did not run; the host counters advance per replay below)
            if self.ws.gen != self._ws_gen:  # the workspace grew inside the capture: that graph points at freed memory
                self._ws_gen = self.ws.gen
                return self._step_bayes_eager_after_growth(sp, b0, B, train, lr, loss_slot, loss_scale, gbatch)
            self._graphs[key] = g
        g.launch()
        self.global_step += 1
        if train: self.adam_t += 1

    def _step_bayes_eager_after_growth(self, sp, b0, B, train, lr, loss_slot, loss_scale, gbatch):
        self._graphs.clear()
        self._bayes_body(sp, b0, B, train, lr, loss_slot, None, loss_scale, gbatch, None)

    def _bayes_body(self, sp, b0, B, train, lr, loss_slot, neg_host, loss_scale, gbatch, noise_host=None):
        h, Lo = self.hidden, self.L - 1
        scale = 1.0 / B if loss_scale is None else loss_scale
        self._draw_noise(sp, b0, B, noise_host)
        self._prepare_bayes()
        self._forward_hidden_bayes(sp, b0, B)
        neg = self._sample(sp, b0, B, neg_host, gbatch)
        mptr = sp.m_indptr.data_ptr() + 4 * b0
        ns = 0 if neg is None else neg.shape[1]
        tc = self.precision == _lib.NTF_TF32  # tcgen05 kernels: the sets travel as tile-transposed planes, consumed (cleared) by the kernel
        if tc: ops.special_tiles(1, B, mptr, sp.m_indices, neg, ns, self.E, self.special_t, self.member_t)
        else: ops.special_bits(1, B, mptr, sp.m_indices, neg, ns, self.E, self.special, self.pitch)
        a = OutTrainArgs()
        a.A, a.W, a.b = self.act[-1].data_ptr(), self._pv(Lo, 'mu', 'weight').data_ptr(), self._pv(Lo, 'mu', 'bias').data_ptr()
        a.pitch_words = self.pitch  # (of sign_out; the condition plane has the same pitch)
        if tc: a.special_t, a.member_t = self.special_t.data_ptr(), self.member_t.data_ptr()
        else: a.special = self.special.data_ptr()
        a.m_indptr, a.m_indices = mptr, sp.m_indices.data_ptr()
        a.B, a.h, a.E = B, h[-1], self.E
        a.tpw, a.tnw, a.loss_scale = self.tpw, self.tnw, scale
        a.loss_out = self.loss_buf.data_ptr() + 4 * loss_slot
        a.A_s, a.W_delta, a.b_delta = self.act_s[-1].data_ptr(), self.nview(self.delta, f'{Lo}.weight').data_ptr(), self.nview(self.delta, f'{Lo}.bias').data_ptr()
        a.sign_out = self.sign_out[Lo].data_ptr()
        if train:
            a.dW, a.db = self._pv(Lo, 'mu', 'weight', self.grads).data_ptr(), self._pv(Lo, 'mu', 'bias', self.grads).data_ptr()
            a.dA = self.dact[-1].data_ptr()
            a.dW_delta, a.db_delta = self.nview(self.gdelta, f'{Lo}.weight').data_ptr(), self.nview(self.gdelta, f'{Lo}.bias').data_ptr()
            a.dA_s = self.dact_s[-1].data_ptr()
        ops.out_train(self.dev_index, self.precision, a, self.ws)
        if not tc: ops.special_bits(0, B, mptr, sp.m_indices, neg, ns, self.E, self.special, self.pitch)
        ops.axpy(1, scale / self.world, self.kl, self.loss_buf[loss_slot:loss_slot + 1])  # + KL/B (fnn.py:136,149)
        self.global_step += 1
        if not train: return
        for i in range(self.L - 1, 0, -1):
            # dA_{i-1} = dz mu_W + ((dz*s_out) W_delta) * s_in
            ops.add_signed(self.dact_s[i - 1], self.sign_in[i], self.sign_in[i].shape[1], B, h[i - 1], self.dact[i - 1])
            j = i - 1  # backward through layer j's activation
            ops.act_bwd(self.dact[j], self.act[j], B, h[j], 1, self.dz[j], self._pv(j, 'mu', 'bias', self.grads), self.ws)
            ops.apply_sign(self.dz[j], self.sign_out[j], self.sign_out[j].shape[1], B, h[j], self.dzs[j])
            ops.act_bwd(self.dzs[j], None, B, h[j], 0, None, self.nview(self.gdelta, f'{j}.bias'), self.ws)
            if j == 0: break
            ops.dense_bwd(self.act[j - 1], self._pv(j, 'mu', 'weight'), self.dz[j], B, h[j - 1], h[j], self._pv(j, 'mu', 'weight', self.grads),
                          self.dact[j - 1], self.ws)
            ops.dense_bwd(self.act_s[j - 1], self.nview(self.delta, f'{j}.weight'), self.dzs[j], B, h[j - 1], h[j], self.nview(self.gdelta, f'{j}.weight'),
                          self.dact_s[j - 1], self.ws)
        if self.dense_input:  # dW0 = dz0^T x, dW0_delta = (dz0*s_out)^T (x*s_in); the input needs no gradient
            ops.dense_bwd(sp.x[b0:b0 + B], self._pv(0, 'mu', 'weight'), self.dz[0], B, self.S, h[0], self._pv(0, 'mu', 'weight', self.grads), None, self.ws)
            ops.dense_bwd(self.x_s, self.nview(self.delta, '0.weight'), self.dzs[0], B, self.S, h[0], self.nview(self.gdelta, '0.weight'), None, self.ws)
        else:
            sptr = sp.s_indptr.data_ptr() + 4 * b0
            ops.csr_bag_bwd(B, sptr, sp.s_indices, sp.s_ent_row, b0, self.dz[0], self.S, h[0], self._pv(0, 'mu', 'weight', self.grads), self.ws)
            ops.csr_bag_bwd_signed(B, sptr, sp.s_indices, sp.s_ent_row, b0, self.ent_sign, self.dzs[0], self.S, h[0], self.nview(self.gdelta, '0.weight'), self.ws)
        for i in range(self.L):
            for what in ('weight', 'bias'):
                mu, rho = self._pv(i, 'mu', what), self._pv(i, 'rho', what)
                n = mu.numel()
                ops.flipout_grads(mu, rho, self.nview(self.eps, f'{i}.{what}'), self.nview(self.gdelta, f'{i}.{what}'), n, scale / (n * self.world),
                                  self._pv(i, 'mu', what, self.grads), self._pv(i, 'rho', what, self.grads))
        self.optimizer_step(lr)

    def scores_mc(self, sp, b0, B, nmc, out, scratch, ent_pred, ent_model, noise_host=None):
        """fnn.py:202-209 (Bnn): out[:B] = mean over nmc stochastic passes of sigmoid(model(X)); ent_pred / ent_model [B] =
        predictive entropy of the mean / mutual information (bayesian_torch.utils.util)."""
        Lo = self.L - 1
        out[:B].zero_()
        for m in range(nmc):
            self._draw_noise(sp, b0, B, None if noise_host is None else noise_host[m])
            self._prepare_bayes()
            self._forward_hidden_bayes(sp, b0, B)
            ops.infer_scores(self.precision, self.act[-1], self._pv(Lo, 'mu', 'weight'), self._pv(Lo, 'mu', 'bias'), B, self.hidden[-1], self.E, scratch, self.ws,
                             accumulate=0, A_s=self.act_s[-1], W_delta=self.nview(self.delta, f'{Lo}.weight'), b_delta=self.nview(self.delta, f'{Lo}.bias'),
                             sign_out=self.sign_out[Lo], pitch=self.pitch)
            ops.row_entropy(scratch, B, self.E, 1.0, int(m > 0), self.ent_sum)   # sum over samples of the per-sample entropy
            ops.axpy(B * self.E, 1.0 / nmc, scratch, out)
            self.global_step += 1
        ops.row_entropy(out, B, self.E, 1.0, 0, ent_pred)
        ops.row_entropy(out, B, self.E, 1.0, 0, ent_model)
        ops.axpy(B, -1.0 / nmc, self.ent_sum, ent_model)   # mutual information = H(mean) - mean_m H_m
        return out

    # ------------------------------------------------------------------ streaming entry point (host batches)
    def step_host(self, packed, n, cap_s, cap_m, rank=0, G=1, lr=1e-3, train=True, sync=True, slot=0):
        """one step on a batch the HOST holds (`pack_host_batch`: compact CSR of the global batch in one pinned block): one H2D copy,
        this rank trains on its slice, and the loss comes back to the host -- the per-step shape of the reference's loop (fnn.py:118-140: H2D
        of the batch, .item()).
        sync=False is the streaming form: `slot` (0/1, alternating) names a device staging block, a loss slot and a pinned host slot; the copies
        run on a copy stream -- batch i+1 goes up while step i computes, loss i comes down while step i+1 computes -- and the caller reads the
        loss with step_host_loss(slot) AFTER it has enqueued the next step, so the compute stream runs the steps back to back."""
        b = -(-n // G)
        lo, hi = min(n, rank * b), min(n, (rank + 1) * b)
        if sync:
            st = getattr(self, '_hstage', None)
            if st is None or st.buf.numel() < packed.numel():
                st = self._hstage = _HostStage(self.device, 2 * packed.numel() + 1024)
            st.view(n, cap_s, cap_m)
            st.buf[:packed.numel()].copy_(packed, non_blocking=True)
            self.step(st, lo, hi - lo, train, lr=lr, loss_slot=slot, loss_scale=1.0 / n, gbatch=(0, n))
            return float(self.loss_buf[slot].item())
        hs = getattr(self, '_hstream', None)
        if hs is None or hs['stage'][0].buf.numel() < packed.numel():
            hs = self._hstream = {'stage': [_HostStage(self.device, 2 * packed.numel() + 1024) for _ in range(2)], 'shape': [None, None],
                                  'loss': torch.zeros(2, dtype=torch.float32).pin_memory()}
            hs['loss_np'] = hs['loss'].numpy()
        st = hs['stage'][slot]
        if hs['shape'][slot] != (n, cap_s, cap_m):
            st.view(n, cap_s, cap_m); hs['shape'][slot] = (n, cap_s, cap_m)
        L, stream = _lib.lib(), torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(L.ntf_host_batch_upload(self.h, stream, slot, st.buf.data_ptr(), packed.data_ptr(), packed.numel() * 4), 'ntf_host_batch_upload')
        self.step(st, lo, hi - lo, train, lr=lr, loss_slot=slot, loss_scale=1.0 / n, gbatch=(0, n))
        _lib.check(L.ntf_host_loss_download(self.h, stream, slot, self.loss_buf.data_ptr() + 4 * slot, hs['loss'].data_ptr() + 4 * slot), 'ntf_host_loss_download')
        return None

    def step_host_loss(self, slot=0):
        """the loss of the step_host(sync=False, slot=slot) enqueued last on that slot: waits for ITS copy only (fnn.py:140 `loss.item()`)"""
        _lib.check(_lib.lib().ntf_host_loss_wait(self.h, slot), 'ntf_host_loss_wait')
        return float(self._hstream['loss_np'][slot])

    # ------------------------------------------------------------------ inference
    def scores(self, sp, b0, B, out):
        """out[:B] = sigmoid(model(X)) for rows [b0,b0+B) (fnn.py:211), dense [B,E] on the device."""
        self._forward_hidden(sp, b0, B)
        Lo = self.L - 1
        ops.infer_scores(self.precision, self.act[-1], self.view(f'layers.{Lo}.weight'), self.view(f'layers.{Lo}.bias'), B, self.hidden[-1],
                         self.E, out, self.ws)
        return out

    def w16(self):
        """fp16 image [E,h_last] of the output layer's weight: what the tensor-core kernels TMA-load (training: out_tc2.cu, test time:
        infer_topk.cu).  ntf_fnn_step keeps it current while it steps the parameters itself; every other writer of the arena
        (load_state_dict, eager optimiser steps) marks it dirty and it is remade here on the next use."""
        if getattr(self, '_w16', None) is None:
            W = self.view(f'layers.{self.L - 1}.weight')
            self._w16, self._w16_dirty = torch.empty(W.numel(), dtype=torch.float16, device=self.device), True
        if self._w16_dirty:
            W = self.view(f'layers.{self.L - 1}.weight')
            ops.to_half(W, W.numel(), self._w16)
            self._w16_dirty = False
        return self._w16

    def topk(self, sp, b0, B, K, scores_buf, vals, idx):
        """the K best experts per team in rank order; only [B,K] leaves the GPU (fnn.py:213-218 keeps [N,E] on the host).
        Tensor-core mode with K <= 1024 and 32*K <= E: hidden layers + output layer + sigmoid + selection in ONE library call
        (ntf_fnn_infer_topk), the [B,E] scores never reach HBM; otherwise scores -> ntf_topk_select."""
        if self.fused_topk_ok(B, K):
            if self.shard[1] == 1: return self._topk_one_call(sp, b0, B, K, vals, idx)
            self._forward_hidden(sp, b0, B)
            return self.select_topk_fused(B, K, vals, idx)
        self.scores(sp, b0, B, scores_buf)
        return self.select_topk(scores_buf, B, K, vals, idx)

    def fused_topk_ok(self, B, K):
        key = (B, K)
        c = getattr(self, '_fused_ok', None)
        if c is None: c = self._fused_ok = {}
        if key not in c:
            c[key] = (not self.bayesian and self.precision == _lib.NTF_TF32 and ops.infer_topk_supported(B, self.hidden[-1], self.E, min(K, self.E)))
        return c[key] and os.environ.get('NTF_FUSED_TOPK', '1') != '0'

    def _topk_one_call(self, sp, b0, B, K, vals, idx, e_lo=0):
        a = getattr(self, '_ita', None)
        if a is None:
            a = self._ita = _lib.FnnInferTopkArgs()
            a.n_layers, a.S, a.E = self.L, self.S, self.E
            for i, hh in enumerate(self.hidden): a.hidden[i], a.act[i] = hh, self.act[i].data_ptr()
            self._ita_ws = {}
        if getattr(self, '_ita_params', None) != self.params.data_ptr():  # (the arena moves when peers are attached)
            for i in range(self.L): a.W[i], a.b[i] = self.view(f'layers.{i}.weight').data_ptr(), self.view(f'layers.{i}.bias').data_ptr()
            self._ita_params = self.params.data_ptr()
        a.W16 = self.w16().data_ptr()
        a.B, a.K, a.e_lo = B, K, e_lo
        if self.dense_input: a.x_dense, a.s_indptr, a.s_indices = sp.x.data_ptr() + 4 * self.S * b0, None, None
        else: a.x_dense, a.s_indptr, a.s_indices = None, sp.s_indptr.data_ptr() + 4 * b0, sp.s_indices.data_ptr()
        a.vals, a.idx = vals.data_ptr(), idx.data_ptr()
        self._ita_ws[(B, K)] = ops.fnn_infer_topk(self.dev_index, a, self.ws, self._ita_ws.get((B, K)))
        return vals, idx

    def select_topk_fused(self, B, K, vals, idx):
        """top-K of the batch whose last hidden activations are in self.act[-1].  Expert-sharded: local fused top-K (global ids), then the
        all-gather + ntf_topk_merge of select_topk"""
        Lo, n = self.L - 1, self.shard[1]
        if n == 1:
            ops.infer_topk(self.act[-1], self.w16(), self.view(f'layers.{Lo}.bias'), B, self.hidden[-1], self.E, K, vals, idx, self.ws)
            return vals, idx
        Kl = min(K, self.E)
        lv = torch.zeros(B, K, dtype=torch.float32, device=self.device)
        li = torch.full((B, K), -1, dtype=torch.int32, device=self.device)
        v_, i_ = torch.empty(B, Kl, dtype=torch.float32, device=self.device), torch.empty(B, Kl, dtype=torch.int32, device=self.device)
        ops.infer_topk(self.act[-1], self.w16(), self.view(f'layers.{Lo}.bias'), B, self.hidden[-1], self.E, Kl, v_, i_, self.ws, e_lo=self.e_lo)
        lv[:, :Kl] = v_; li[:, :Kl] = i_
        gv, gi = self.allgather(lv), self.allgather(li)
        ops.topk_merge(gv, gi, n, B, K, vals, idx)
        return vals, idx

    def select_topk(self, scores_buf, B, K, vals, idx):
        """rank order per team.  Expert-sharded: every rank ranks its own columns, the [B,K] lists (global expert ids) are
        all-gathered and merged by ntf_topk_merge (a collective)"""
        n = self.shard[1]
        if n == 1:
            ops.topk_select(scores_buf, B, self.E, K, 1.0, vals, idx)
            return vals, idx
        Kl = min(K, self.E)
        lv = torch.zeros(B, K, dtype=torch.float32, device=self.device)
        li = torch.full((B, K), -1, dtype=torch.int32, device=self.device)
        v_, i_ = torch.empty(B, Kl, dtype=torch.float32, device=self.device), torch.empty(B, Kl, dtype=torch.int32, device=self.device)
        ops.topk_select(scores_buf, B, self.E, Kl, 1.0, v_, i_)
        lv[:, :Kl] = v_; li[:, :Kl] = i_ + self.e_lo
        gv, gi = self.allgather(lv), self.allgather(li)  # [n, B, K]
        ops.topk_merge(gv, gi, n, B, K, vals, idx)
        return vals, idx

    def allgather(self, t):
        parts = [torch.empty_like(t) for _ in range(self.shard[1])]
        torch.distributed.all_gather(parts, t.contiguous())
        return torch.stack(parts).contiguous()
